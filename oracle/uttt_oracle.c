/*
 * uttt_oracle.c -- TEST INFRASTRUCTURE ONLY (see uttt_oracle.h).
 *
 * Plain-C restatement of the reference's rules (cpp/uttt_game.cpp) and search
 * driver (cpp/uttt_mcts.cpp).  Deliberately written on int arrays with the
 * reference's own control flow (not on bitboards) so it is an independent
 * check of the bitboard CUDA kernels.  Build: see oracle/Makefile.
 * Floating point: compile WITHOUT -ffast-math / -march=native so every float op
 * is a single IEEE-754 binary32 operation, like the reference build
 * (cpp/setup.py:10, "-O3" only).
 */
#include "uttt_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ rules */

/* cpp/uttt_game.cpp:9-19 */
void orc_init(orc_state *s) {
    memset(s, 0, sizeof(*s));
    s->active = -1;
}

/* cpp/uttt_game.cpp:35-61 -- a line counts when all three cells are non-zero */
static int line3(const int b[9], int x, int y, int dx, int dy) {
    for (int k = 0; k < 3; k++) {
        if (y < 0 || y > 2 || x < 0 || x > 2 || b[x + y * 3] == 0) return 0;
        x += dx;
        y += dy;
    }
    return 1;
}
int orc_check_win(const int b[9]) {
    if (line3(b, 0, 0, 1, 1) || line3(b, 0, 2, 1, -1)) return 1;
    for (int i = 0; i < 3; i++)
        if (line3(b, 0, i, 1, 0) || line3(b, i, 0, 0, 1)) return 1;
    return 0;
}

/* cpp/uttt_game.cpp:64-74 -- counts cells equal to 1 */
static int piece_count(const int p[9][9]) {
    int n = 0;
    for (int b = 0; b < 9; b++)
        for (int c = 0; c < 9; c++)
            if (p[b][c] == 1) n++;
    return n;
}

int orc_is_lose(const orc_state *s) { return orc_check_win(s->main_enemy); }      /* :77-79 */
int orc_is_draw(const orc_state *s) {                                              /* :82-84 */
    int tmp[81];
    return !orc_is_lose(s) && orc_legal_actions(s, tmp) == 0;
}
int orc_is_done(const orc_state *s) { return orc_is_lose(s) || orc_is_draw(s); }  /* :87-89 */
int orc_is_first_player(const orc_state *s) {                                      /* :92-94 */
    return piece_count(s->pieces) == piece_count(s->enemy);
}

/* cpp/uttt_game.cpp:97-145 */
void orc_next(const orc_state *s, int action, orc_state *out) {
    int b = action / 9, c = action % 9;
    orc_state n;
    memcpy(n.pieces, s->enemy, sizeof(n.pieces));          /* sides swap */
    memcpy(n.enemy, s->pieces, sizeof(n.enemy));
    memcpy(n.main_pieces, s->main_enemy, sizeof(n.main_pieces));
    memcpy(n.main_enemy, s->main_pieces, sizeof(n.main_enemy));
    n.enemy[b][c] = 1;                                     /* the stone just played */
    if (orc_check_win(n.enemy[b])) {
        n.main_enemy[b] = 1;
    } else {
        int full = 1;
        for (int j = 0; j < 9; j++)
            if (n.pieces[b][j] == 0 && n.enemy[b][j] == 0) { full = 0; break; }
        if (full) { n.main_pieces[b] = 1; n.main_enemy[b] = 1; }   /* drawn board marks both */
    }
    n.active = c;
    if (n.main_pieces[c] == 1 || n.main_enemy[c] == 1) n.active = -1;
    *out = n;
}

/* cpp/uttt_game.cpp:148-191 -- ascending action ids */
int orc_legal_actions(const orc_state *s, int out[81]) {
    int n = 0;
    if (orc_is_lose(s)) return 0;
    int cand[9], nc = 0;
    if (s->active == -1) {
        for (int i = 0; i < 9; i++)
            if (s->main_pieces[i] == 0 && s->main_enemy[i] == 0) cand[nc++] = i;
    } else if (s->main_pieces[s->active] == 0 && s->main_enemy[s->active] == 0) {
        cand[nc++] = s->active;
    } else {
        for (int i = 0; i < 9; i++)
            if (s->main_pieces[i] == 0 && s->main_enemy[i] == 0) cand[nc++] = i;
    }
    for (int k = 0; k < nc; k++) {
        int b = cand[k];
        for (int c = 0; c < 9; c++)
            if (s->pieces[b][c] == 0 && s->enemy[b][c] == 0) out[n++] = b * 9 + c;
    }
    return n;
}

/* cpp/uttt_game.cpp:244-280 -- HWC float[9*9*3]; ch0 mover, ch1 opponent, ch2 legal */
void orc_to_input_tensor(const orc_state *s, float out[243]) {
    int legal[81];
    int nl = orc_legal_actions(s, legal);
    for (int i = 0; i < 243; i++) out[i] = 0.0f;
    for (int b = 0; b < 9; b++)
        for (int c = 0; c < 9; c++) {
            int R = (b / 3) * 3 + (c / 3), C = (b % 3) * 3 + (c % 3);
            if (s->pieces[b][c] == 1) out[R * 27 + C * 3 + 0] = 1.0f;
            if (s->enemy[b][c] == 1) out[R * 27 + C * 3 + 1] = 1.0f;
        }
    for (int k = 0; k < nl; k++) {
        int b = legal[k] / 9, c = legal[k] % 9;
        int R = (b / 3) * 3 + (c / 3), C = (b % 3) * 3 + (c % 3);
        out[R * 27 + C * 3 + 2] = 1.0f;
    }
}

/* cpp/uttt_game.cpp:194-241 */
int orc_to_string(const orc_state *s, char *buf, int cap) {
    const char *ox = orc_is_first_player(s) ? "ox" : "xo";
    int n = 0;
#define PUT(...) do { n += snprintf(buf + n, (n < cap) ? (size_t)(cap - n) : 0, __VA_ARGS__); } while (0)
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) {
            for (int i = 0; i < 3; i++) {
                int b = r * 3 + i;
                for (int j = 0; j < 3; j++) {
                    int cell = c * 3 + j;
                    char p = '-';
                    if (s->pieces[b][cell] == 1) p = ox[0];
                    else if (s->enemy[b][cell] == 1) p = ox[1];
                    PUT("%c ", p);
                }
                if (i < 2) PUT("| ");
            }
            PUT("\n");
        }
        if (r < 2) PUT("---------------------\n");
    }
    PUT("\nMain Board Status:\n");
    for (int i = 0; i < 9; i++) {
        char m = '.';
        if (s->main_pieces[i] == 1 && s->main_enemy[i] == 1) m = 'D';
        else if (s->main_pieces[i] == 1) m = ox[0];
        else if (s->main_enemy[i] == 1) m = ox[1];
        PUT("%c", m);
        if (i % 3 == 2) PUT("\n");
    }
    PUT("Next Player: %c\n", ox[0]);
    if (s->active == -1) PUT("Active Board: Any\n");
    else PUT("Active Board: %d\n", s->active);
#undef PUT
    return n;
}

/* ------------------------------------------------- packed form (8 x u32) */
/* w0..w2: mover cells, 3 boards per word, 9 bits per board; w3..w5: opponent;
 * w6: main_pieces | main_enemy<<9 | (active+1)<<18; w7: 0 (reserved).         */
void orc_pack(const orc_state *s, uint32_t w[8]) {
    memset(w, 0, 8 * sizeof(uint32_t));
    for (int b = 0; b < 9; b++)
        for (int c = 0; c < 9; c++) {
            if (s->pieces[b][c]) w[b / 3] |= 1u << ((b % 3) * 9 + c);
            if (s->enemy[b][c]) w[3 + b / 3] |= 1u << ((b % 3) * 9 + c);
        }
    for (int b = 0; b < 9; b++) {
        if (s->main_pieces[b]) w[6] |= 1u << b;
        if (s->main_enemy[b]) w[6] |= 1u << (9 + b);
    }
    w[6] |= (uint32_t)(s->active + 1) << 18;
}
void orc_unpack(const uint32_t w[8], orc_state *s) {
    for (int b = 0; b < 9; b++)
        for (int c = 0; c < 9; c++) {
            s->pieces[b][c] = (w[b / 3] >> ((b % 3) * 9 + c)) & 1;
            s->enemy[b][c] = (w[3 + b / 3] >> ((b % 3) * 9 + c)) & 1;
        }
    for (int b = 0; b < 9; b++) {
        s->main_pieces[b] = (w[6] >> b) & 1;
        s->main_enemy[b] = (w[6] >> (9 + b)) & 1;
    }
    s->active = (int)((w[6] >> 18) & 15) - 1;
}

/* ------------------------------------------------------------- Philox RNG */
static inline void mulhilo(uint32_t a, uint32_t b, uint32_t *hi, uint32_t *lo) {
    uint64_t p = (uint64_t)a * b;
    *hi = (uint32_t)(p >> 32);
    *lo = (uint32_t)p;
}
/* Philox4x32-10 (Salmon et al., SC'11), the published algorithm. */
void orc_philox4x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2,
                    uint32_t c3, uint32_t out[4]) {
    for (int r = 0; r < 10; r++) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo(0xD2511F53u, c0, &hi0, &lo0);
        mulhilo(0xCD9E8D57u, c2, &hi1, &lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* ------------------------------------------- deterministic hash evaluator */
static inline uint32_t mix32(uint32_t x) {     /* murmur3 finaliser */
    x ^= x >> 16; x *= 0x85EBCA6Bu;
    x ^= x >> 13; x *= 0xC2B2AE35u;
    x ^= x >> 16;
    return x;
}
/* FNV-1a(32) over the 81 cells (0 empty / 1 mover / 2 opponent), the 9 main
 * flags (mover | opponent<<1) and active+1. */
uint32_t orc_state_hash(const orc_state *s) {
    uint32_t h = 0x811C9DC5u;
    for (int b = 0; b < 9; b++)
        for (int c = 0; c < 9; c++) {
            uint32_t v = s->pieces[b][c] ? 1u : (s->enemy[b][c] ? 2u : 0u);
            h = (h ^ v) * 0x01000193u;
        }
    for (int b = 0; b < 9; b++) {
        uint32_t v = (s->main_pieces[b] ? 1u : 0u) | (s->main_enemy[b] ? 2u : 0u);
        h = (h ^ v) * 0x01000193u;
    }
    h = (h ^ (uint32_t)(s->active + 1)) * 0x01000193u;
    return h;
}
/* policy[a] = ((mix(h+a)&0xFFFF)+1)/65536/81 (strictly positive, not a sum-to-one
 * vector and not exactly representable, so the serial renormalisation of
 * cpp/uttt_mcts.cpp:144-158 is exercised); value is an exact dyadic rational. */
void orc_hash_eval(const orc_state *s, float policy[81], float *value) {
    uint32_t h = orc_state_hash(s);
    for (int a = 0; a < 81; a++) {
        float num = (float)((mix32(h + (uint32_t)a) & 0xFFFFu) + 1u);
        policy[a] = (num / 65536.0f) / 81.0f;
    }
    *value = ((float)(int)(mix32(h ^ 0xABCDu) & 0xFFFFu) - 32768.0f) / 32768.0f;
}
static void hash_eval_cb(void *ctx, const orc_state *st, int n, float *pol, float *val) {
    (void)ctx;
    for (int i = 0; i < n; i++) orc_hash_eval(&st[i], pol + 81 * i, val + i);
}

/* ----------------------------------------------------------------- search */
/* cpp/uttt_mcts.h:22-50 */
typedef struct orc_node {
    orc_state state;
    float p, w;
    int n;
    struct orc_node **child;
    int n_child, cap_child;
} orc_node;

static int g_nodes_made;

static orc_node *node_new(const orc_state *s, float p) {     /* cpp/uttt_mcts.cpp:10-12 */
    orc_node *nd = (orc_node *)calloc(1, sizeof(orc_node));
    nd->state = *s;
    nd->p = p;
    g_nodes_made++;
    return nd;
}
static void node_free(orc_node *nd) {
    for (int i = 0; i < nd->n_child; i++) node_free(nd->child[i]);
    free(nd->child);
    free(nd);
}

/* cpp/uttt_mcts.cpp:57-81 -- PUCT with C=1, first maximum wins */
static orc_node *next_child_node(orc_node *nd) {
    const float C_PUCT = 1.0f;
    int total_n = 0;
    for (int i = 0; i < nd->n_child; i++) total_n += nd->child[i]->n;
    float sqrt_total = sqrtf((float)total_n);
    float max_pucb = -1e9f;
    orc_node *best = NULL;
    for (int i = 0; i < nd->n_child; i++) {
        orc_node *c = nd->child[i];
        float q = (c->n > 0) ? (-c->w / (float)c->n) : 0.0f;
        float u = C_PUCT * c->p * sqrt_total / (float)(1 + c->n);
        float pucb = q + u;
        if (pucb > max_pucb) { max_pucb = pucb; best = c; }
    }
    return best;
}

/* cpp/uttt_mcts.cpp:15-32 */
static orc_node *search_leaf(orc_node *nd, orc_node **path, int *plen, float *value) {
    for (;;) {
        path[(*plen)++] = nd;
        if (orc_is_done(&nd->state)) {
            float v = orc_is_lose(&nd->state) ? -1.0f : 0.0f;
            *value = -v;                       /* sign as written in the reference (Q-M2) */
            return nd;
        }
        if (nd->n_child == 0) { *value = 0.0f; return nd; }
        nd = next_child_node(nd);
    }
}

/* cpp/uttt_mcts.cpp:35-44 -- appends, never clears */
static void expand(orc_node *nd, const float *policies, int npol) {
    int legal[81];
    int nl = orc_legal_actions(&nd->state, legal);
    for (int i = 0; i < nl; i++) {
        float p = (i < npol) ? policies[i] : 0.0f;
        orc_state nx;
        orc_next(&nd->state, legal[i], &nx);
        if (nd->n_child == nd->cap_child) {
            nd->cap_child = nd->cap_child ? nd->cap_child * 2 : 16;
            nd->child = (orc_node **)realloc(nd->child, sizeof(orc_node *) * (size_t)nd->cap_child);
        }
        nd->child[nd->n_child++] = node_new(&nx, p);
    }
}

/* cpp/uttt_mcts.cpp:47-54 */
static void backpropagate(orc_node **path, int plen, float value) {
    for (int i = plen - 1; i >= 0; i--) {
        path[i]->w += value;
        path[i]->n += 1;
        value = -value;
    }
}

/* cpp/uttt_mcts.cpp:199-216 */
void orc_boltzman(const float *xs, int n, float temperature, float *out) {
    float sum = 0.0f;
    for (int i = 0; i < n; i++) {
        float v = powf(xs[i], 1.0f / temperature);
        out[i] = v;
        sum += v;
    }
    if (sum > 0)
        for (int i = 0; i < n; i++) out[i] /= sum;
}

#define ORC_MAX_PATH 128

/* cpp/uttt_mcts.cpp:84-196, loop structure kept literal (queue, flush rule). */
int orc_pv_mcts_scores(orc_eval_fn eval, void *ctx, const orc_state *root_state,
                       float temperature, int evaluate_count, int batch_size,
                       float *scores_out, int *counts_out, int *stats_out) {
    int root_legal[81];
    int nroot = orc_legal_actions(root_state, root_legal);
    if (stats_out) stats_out[0] = stats_out[1] = stats_out[2] = 0;
    if (nroot == 0) return 0;                                             /* :96-98 */

    g_nodes_made = 0;
    orc_node *root = node_new(root_state, 0.0f);
    float init_pol[81];
    float uniform = 1.0f / (float)nroot;                                  /* :101 */
    for (int i = 0; i < nroot; i++) init_pol[i] = uniform;
    expand(root, init_pol, nroot);

    int qcap = batch_size > 0 ? batch_size : 1;
    orc_node **q_leaf = (orc_node **)malloc(sizeof(orc_node *) * (size_t)qcap);
    orc_node **q_path = (orc_node **)malloc(sizeof(orc_node *) * (size_t)qcap * ORC_MAX_PATH);
    int *q_plen = (int *)malloc(sizeof(int) * (size_t)qcap);
    orc_state *bstates = (orc_state *)malloc(sizeof(orc_state) * (size_t)qcap);
    float *bpol = (float *)malloc(sizeof(float) * 81 * (size_t)qcap);
    float *bval = (float *)malloc(sizeof(float) * (size_t)qcap);
    int nq = 0, ncalls = 0, nstates = 0;

    for (int i = 0; i < evaluate_count; i++) {
        orc_node *path[ORC_MAX_PATH];
        int plen = 0;
        float value;
        orc_node *leaf = search_leaf(root, path, &plen, &value);          /* :112 */
        if (orc_is_done(&leaf->state)) {                                  /* :115-118 */
            backpropagate(path, plen, value);
            continue;
        }
        if (leaf->n == 0 && leaf->n_child == 0) {                         /* :121-124 */
            q_leaf[nq] = leaf;
            memcpy(q_path + (size_t)nq * ORC_MAX_PATH, path, sizeof(orc_node *) * (size_t)plen);
            q_plen[nq] = plen;
            nq++;
        }
        if (nq >= batch_size || i == evaluate_count - 1) {                /* :127 */
            if (nq > 0) {
                for (int j = 0; j < nq; j++) bstates[j] = q_leaf[j]->state;
                eval(ctx, bstates, nq, bpol, bval);                       /* :135 */
                ncalls++;
                nstates += nq;
                for (int j = 0; j < nq; j++) {                            /* :138-167 */
                    orc_node *lf = q_leaf[j];
                    int legal[81];
                    int nl = orc_legal_actions(&lf->state, legal);
                    float lp[81];
                    float sum = 0.0f;
                    for (int k = 0; k < nl; k++) {
                        float p = bpol[81 * j + legal[k]];
                        lp[k] = p;
                        sum += p;
                    }
                    if (sum > 0) {
                        for (int k = 0; k < nl; k++) lp[k] /= sum;
                    } else {
                        float u = nl == 0 ? 0.0f : 1.0f / (float)nl;
                        for (int k = 0; k < nl; k++) lp[k] = u;
                    }
                    expand(lf, lp, nl);
                    backpropagate(q_path + (size_t)j * ORC_MAX_PATH, q_plen[j], bval[j]);
                }
                nq = 0;
            }
        }
    }

    float scores[81 * 8];
    int nchild = root->n_child;      /* == nroot: the root is expanded exactly once */
    for (int i = 0; i < nchild; i++) {
        scores[i] = (float)root->child[i]->n;                             /* :177-180 */
        if (counts_out) counts_out[i] = root->child[i]->n;
    }
    if (temperature == 0.0f) {                                            /* :183-189 */
        int best = 0;
        for (int i = 1; i < nchild; i++)
            if (scores[i] > scores[best]) best = i;
        for (int i = 0; i < nchild; i++) scores_out[i] = 0.0f;
        if (nchild > 0) scores_out[best] = 1.0f;
    } else {
        orc_boltzman(scores, nchild, temperature, scores_out);
    }
    if (stats_out) { stats_out[0] = g_nodes_made; stats_out[1] = ncalls; stats_out[2] = nstates; }

    node_free(root);
    free(q_leaf); free(q_path); free(q_plen); free(bstates); free(bpol); free(bval);
    return nchild;
}

int orc_pv_mcts_scores_hash(const orc_state *root, float temperature, int evaluate_count,
                            int batch_size, float *scores_out, int *counts_out, int *stats_out) {
    return orc_pv_mcts_scores(hash_eval_cb, NULL, root, temperature, evaluate_count, batch_size,
                              scores_out, counts_out, stats_out);
}

/* ------------------------------------------------ the reference's pure-Python search (pv_mcts.py:74-180)
 * Used by its gating match (evaluate_network.py:73-75).  Same PUCT / backup arithmetic as the C++ search (float32
 * throughout under NumPy >= 2 scalar promotion: the network outputs are np.float32, Python ints / floats are weak), but
 *   - the root starts unexpanded and is evaluated like any leaf (pv_mcts.py:133,142),
 *   - expand() REPLACES the child list (pv_mcts.py:105-109): the k queued copies of a leaf leave ONE list,
 *   - priors: policy[legal] / np.sum(policy[legal]) with numpy's float32 pairwise summation (pv_mcts.py:45-49),
 *   - scores: visit counts, turned into float64 probabilities by the caller (pv_mcts.py:166-180).
 * The loop is kept literal (queue, flush rule pv_mcts.py:154-165). */
static float numpy_sum_f32(const float *a, int n) {
    if (n < 8) {
        float res = 0.0f;
        for (int i = 0; i < n; i++) res += a[i];
        return res;
    }
    float r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; j++) r[j] += a[i + j];
    float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
}

static void expand_replace(orc_node *nd, const float *policies, int npol) {      /* pv_mcts.py:105-109 */
    for (int i = 0; i < nd->n_child; i++) node_free(nd->child[i]);
    nd->n_child = 0;
    expand(nd, policies, npol);
}

int orc_py_mcts_counts(orc_eval_fn eval, void *ctx, const orc_state *root_state, int evaluate_count, int batch_size,
                       int *counts_out) {
    int root_legal[81];
    int nroot = orc_legal_actions(root_state, root_legal);
    if (nroot == 0 || orc_is_lose(root_state)) return 0;
    orc_node *root = node_new(root_state, 0.0f);                          /* pv_mcts.py:133 */
    int qcap = batch_size > 0 ? batch_size : 1;
    orc_node **q_leaf = (orc_node **)malloc(sizeof(orc_node *) * (size_t)qcap);
    orc_node **q_path = (orc_node **)malloc(sizeof(orc_node *) * (size_t)qcap * ORC_MAX_PATH);
    int *q_plen = (int *)malloc(sizeof(int) * (size_t)qcap);
    orc_state *bstates = (orc_state *)malloc(sizeof(orc_state) * (size_t)qcap);
    float *bpol = (float *)malloc(sizeof(float) * 81 * (size_t)qcap);
    float *bval = (float *)malloc(sizeof(float) * (size_t)qcap);
    int nq = 0;
    for (int i = 0; i < evaluate_count; i++) {                            /* pv_mcts.py:139 */
        orc_node *path[ORC_MAX_PATH];
        int plen = 0;
        float value;
        orc_node *leaf = search_leaf(root, path, &plen, &value);          /* :142 (same descent and terminal sign) */
        if (orc_is_done(&leaf->state)) {                                  /* :145-147 */
            backpropagate(path, plen, value);
            continue;
        }
        if (leaf->n == 0) {                                               /* :150-152 */
            q_leaf[nq] = leaf;
            memcpy(q_path + (size_t)nq * ORC_MAX_PATH, path, sizeof(orc_node *) * (size_t)plen);
            q_plen[nq] = plen;
            nq++;
        }
        if (nq >= batch_size || i == evaluate_count - 1) {                /* :157 */
            if (nq > 0) {
                for (int j = 0; j < nq; j++) bstates[j] = q_leaf[j]->state;
                eval(ctx, bstates, nq, bpol, bval);                       /* :160-161 predict_batch */
                for (int j = 0; j < nq; j++) {
                    orc_node *lf = q_leaf[j];
                    int legal[81];
                    int nl = orc_legal_actions(&lf->state, legal);
                    float lp[81];
                    for (int k = 0; k < nl; k++) lp[k] = bpol[81 * j + legal[k]];
                    float sum = numpy_sum_f32(lp, nl);                    /* :47 */
                    if (sum > 0) {
                        for (int k = 0; k < nl; k++) lp[k] /= sum;        /* :49 */
                    } else {
                        for (int k = 0; k < nl; k++) lp[k] = 1.0f / (float)nl;    /* :54 (float64 there; unreachable) */
                    }
                    expand_replace(lf, lp, nl);                           /* :164 */
                    backpropagate(q_path + (size_t)j * ORC_MAX_PATH, q_plen[j], bval[j]);   /* :165 */
                }
                nq = 0;
            }
        }
    }
    int nchild = root->n_child;
    for (int i = 0; i < nchild; i++) counts_out[i] = root->child[i]->n;
    node_free(root);
    free(q_leaf); free(q_path); free(q_plen); free(bstates); free(bpol); free(bval);
    return nchild;
}

int orc_py_mcts_counts_hash(const orc_state *root, int evaluate_count, int batch_size, int *counts_out) {
    return orc_py_mcts_counts(hash_eval_cb, NULL, root, evaluate_count, batch_size, counts_out);
}

/* Replay evaluator: the search is fed recorded (leaf state -> policy, value) rows, e.g. the rows a GPU engine's network
 * produced for exactly these leaves (tests/test_gpu_replay.py: "identical network outputs on both sides").  Entries are
 * consumed in order; the k queued copies of one leaf inside a batch (cpp/uttt_mcts.cpp:121-127) share one entry. */
typedef struct {
    int n;
    const uint32_t *states;   /* [n][8] packed */
    const float *policy;      /* [n][81] */
    const float *value;       /* [n] */
    uint8_t *used;
    int misses;
} table_ctx;

static void table_eval_cb(void *ctx, const orc_state *st, int n, float *pol, float *val) {
    table_ctx *t = (table_ctx *)ctx;
    int last = -1;
    uint32_t lastw[8];
    for (int i = 0; i < n; i++) {
        uint32_t w[8];
        orc_pack(&st[i], w);
        int e = -1;
        if (last >= 0 && memcmp(w, lastw, 28) == 0) e = last;
        else
            for (int j = 0; j < t->n; j++)
                if (!t->used[j] && memcmp(w, t->states + 8 * j, 28) == 0) { e = j; t->used[j] = 1; break; }
        if (e < 0) {
            t->misses++;
            for (int a = 0; a < 81; a++) pol[81 * i + a] = 1.0f / 81.0f;
            val[i] = 0.0f;
        } else {
            memcpy(pol + 81 * i, t->policy + 81 * e, 81 * sizeof(float));
            val[i] = t->value[e];
            last = e;
            memcpy(lastw, w, sizeof(w));
        }
    }
}

/* returns the number of root children; *misses_out = leaves that had no unused entry (+ entries left unused << 16) */
int orc_pv_mcts_scores_table(const orc_state *root, float temperature, int evaluate_count, int batch_size, int n_entries,
                             const uint32_t *states, const float *policy, const float *value, float *scores_out,
                             int *counts_out, int *misses_out) {
    table_ctx t = {n_entries, states, policy, value, (uint8_t *)calloc((size_t)(n_entries > 0 ? n_entries : 1), 1), 0};
    int n = orc_pv_mcts_scores(table_eval_cb, &t, root, temperature, evaluate_count, batch_size, scores_out, counts_out, NULL);
    int unused = 0;
    for (int j = 0; j < n_entries; j++) unused += !t.used[j];
    if (misses_out) *misses_out = t.misses + (unused << 16);
    free(t.used);
    return n;
}

/* the Python-semantics search fed recorded rows (the gating match's search with a network evaluator, replayed) */
int orc_py_mcts_counts_table(const orc_state *root, int evaluate_count, int batch_size, int n_entries, const uint32_t *states,
                             const float *policy, const float *value, int *counts_out, int *misses_out) {
    table_ctx t = {n_entries, states, policy, value, (uint8_t *)calloc((size_t)(n_entries > 0 ? n_entries : 1), 1), 0};
    int n = orc_py_mcts_counts(table_eval_cb, &t, root, evaluate_count, batch_size, counts_out);
    int unused = 0;
    for (int j = 0; j < n_entries; j++) unused += !t.used[j];
    if (misses_out) *misses_out = t.misses + (unused << 16);
    free(t.used);
    return n;
}

/* the hash-evaluator search, recording every distinct evaluated leaf in order (pins the replay machinery on the CPU:
 * replaying the record must reproduce the scores) */
typedef struct { int cap, n; uint32_t *states; float *policy; float *value; } record_ctx;
static void record_eval_cb(void *ctx, const orc_state *st, int n, float *pol, float *val) {
    record_ctx *r = (record_ctx *)ctx;
    hash_eval_cb(NULL, st, n, pol, val);
    for (int i = 0; i < n; i++) {
        uint32_t w[8];
        orc_pack(&st[i], w);
        if (i > 0) {
            uint32_t p[8];
            orc_pack(&st[i - 1], p);
            if (memcmp(w, p, 28) == 0) continue;
        }
        if (r->n < r->cap) {
            memcpy(r->states + 8 * r->n, w, 32);
            memcpy(r->policy + 81 * r->n, pol + 81 * i, 81 * sizeof(float));
            r->value[r->n] = val[i];
        }
        r->n++;
    }
}
int orc_pv_mcts_scores_hash_record(const orc_state *root, float temperature, int evaluate_count, int batch_size, int cap,
                                   uint32_t *states, float *policy, float *value, int *n_entries_out, float *scores_out) {
    record_ctx r = {cap, 0, states, policy, value};
    int n = orc_pv_mcts_scores(record_eval_cb, &r, root, temperature, evaluate_count, batch_size, scores_out, NULL, NULL);
    *n_entries_out = r.n;
    return n;
}

/* ------------------------------------------------------------ bulk drivers */
static inline uint64_t fnv64(uint64_t h, uint32_t x) {
    return (h ^ (uint64_t)x) * 0x100000001B3ull;
}

static void legal_mask_words(const int *legal, int nl, uint32_t lm[3]) {
    lm[0] = lm[1] = lm[2] = 0;
    for (int k = 0; k < nl; k++) lm[legal[k] / 27] |= 1u << (legal[k] % 27);
}

void orc_playout(uint32_t seed, uint64_t game, uint64_t *digest, int *plies, int *result,
                 uint8_t *actions_out) {
    orc_state s;
    orc_init(&s);
    uint64_t h = 0xCBF29CE484222325ull;
    int t = 0;
    for (;;) {
        int legal[81];
        int nl = orc_legal_actions(&s, legal);
        if (orc_is_lose(&s) || nl == 0) break;
        uint32_t r[4], lm[3], w[8];
        orc_philox4x32(seed, 0u, (uint32_t)game, (uint32_t)(game >> 32), (uint32_t)t, 0u, r);
        int a = legal[r[0] % (uint32_t)nl];
        legal_mask_words(legal, nl, lm);
        orc_pack(&s, w);
        h = fnv64(h, (uint32_t)a);
        h = fnv64(h, lm[0]); h = fnv64(h, lm[1]); h = fnv64(h, lm[2]);
        h = fnv64(h, w[6]);
        if (actions_out) actions_out[t] = (uint8_t)a;
        orc_state nx;
        orc_next(&s, a, &nx);
        s = nx;
        t++;
    }
    uint32_t w[8];
    orc_pack(&s, w);
    int lose = orc_is_lose(&s);
    for (int i = 0; i < 7; i++) h = fnv64(h, w[i]);
    h = fnv64(h, lose ? 1u : 2u);
    *digest = h;
    *plies = t;
    /* 0 draw, 1 first player won, 2 second player won */
    *result = lose ? (orc_is_first_player(&s) ? 2 : 1) : 0;
}

void orc_playouts(uint32_t seed, uint64_t game0, int n, uint64_t *digests, int *plies,
                  int *results) {
    for (int i = 0; i < n; i++)
        orc_playout(seed, game0 + (uint64_t)i, &digests[i], &plies[i], &results[i], NULL);
}

/* self_play_cpp.py:34-101 with np.random.choice replaced by Philox:
 * pick = floor(r * total / 2^32) over the visit counts in legal order. */
int orc_selfplay_hash(uint32_t seed, uint64_t game, int sims, int batch, uint32_t *states,
                      uint16_t *counts, uint8_t *actions, int8_t *z) {
    orc_state s;
    orc_init(&s);
    int t = 0;
    while (!orc_is_done(&s)) {
        int legal[81], cnt[81];
        float scores[81];
        int nl = orc_legal_actions(&s, legal);
        orc_pv_mcts_scores_hash(&s, 1.0f, sims, batch, scores, cnt, NULL);
        int total = 0;
        for (int i = 0; i < nl; i++) total += cnt[i];
        uint32_t r[4];
        orc_philox4x32(seed, 1u, (uint32_t)game, (uint32_t)(game >> 32), (uint32_t)t, 0u, r);
        uint32_t pick = (uint32_t)(((uint64_t)r[0] * (uint64_t)total) >> 32);
        int idx = 0, acc = 0;
        for (int i = 0; i < nl; i++) {
            acc += cnt[i];
            if ((uint32_t)acc > pick) { idx = i; break; }
        }
        orc_pack(&s, states + 8 * t);
        for (int a = 0; a < 81; a++) counts[81 * t + a] = 0;
        for (int i = 0; i < nl; i++) counts[81 * t + legal[i]] = (uint16_t)cnt[i];
        actions[t] = (uint8_t)legal[idx];
        orc_state nx;
        orc_next(&s, legal[idx], &nx);
        s = nx;
        t++;
    }
    int value = orc_is_lose(&s) ? -1 : 0;          /* self_play_cpp.py:95-99 (B3 kept) */
    for (int i = 0; i < t; i++) { z[i] = (int8_t)value; value = -value; }
    return t;
}

/* ------------------------------------------------ throughput-mode cross-check (see uttt_oracle.h) */
static void az_expand(orc_node *nd, const float *pol) {
    int legal[81];
    int nl = orc_legal_actions(&nd->state, legal);
    float lp[81], sum = 0.0f;
    for (int k = 0; k < nl; k++) { lp[k] = pol[legal[k]]; sum += lp[k]; }
    for (int k = 0; k < nl; k++) lp[k] = (sum > 0) ? lp[k] / sum : 1.0f / (float)nl;
    expand(nd, lp, nl);
}

int orc_az_search_hash(const orc_state *root_state, int sims, int *counts_out) {
    int tmp[81];
    if (orc_is_lose(root_state) || orc_legal_actions(root_state, tmp) == 0) return 0;
    float pol[81], val;
    orc_node *root = node_new(root_state, 0.0f);
    orc_hash_eval(root_state, pol, &val);
    az_expand(root, pol);                          /* evaluated root; its value is not backed up */
    for (int i = 0; i < sims; i++) {
        orc_node *path[ORC_MAX_PATH];
        int plen = 0;
        orc_node *nd = root;
        float v;
        for (;;) {
            path[plen++] = nd;
            if (orc_is_done(&nd->state)) { v = orc_is_lose(&nd->state) ? -1.0f : 0.0f; break; }
            if (nd->n_child == 0) {
                orc_hash_eval(&nd->state, pol, &v);
                az_expand(nd, pol);
                break;
            }
            nd = next_child_node(nd);              /* same PUCT arithmetic (C=1, first maximum) */
        }
        backpropagate(path, plen, v);
    }
    for (int k = 0; k < root->n_child; k++) counts_out[k] = root->child[k]->n;
    int n = root->n_child;
    node_free(root);
    return n;
}
