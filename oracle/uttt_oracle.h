/*
 * uttt_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, int arrays, no bitboards) of the reference's
 * Ultimate Tic-Tac-Toe rules and PUCT MCTS driver. It exists to check the
 * CUDA path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product never calls it.
 *
 * Parity status: PINNED. The reference ships no golden vectors of its own
 * (SURVEY.md section 4), so this oracle is pinned against outputs of the
 * reference itself compiled in the build container (oracle/_ref, built by
 * oracle/Makefile from the reference cpp sources) -- see tests/golden/ and
 * oracle/gen_golden.py.
 *
 * Each function cites the reference file:line it restates
 * (paths relative to the reference repository root).
 */
#ifndef UTTT_ORACLE_H
#define UTTT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cpp/uttt_game.h:48-54 -- the mover-relative position. */
typedef struct {
    int pieces[9][9];
    int enemy[9][9];
    int main_pieces[9];
    int main_enemy[9];
    int active;          /* -1 = any board, 0..8 = forced board */
} orc_state;

/* ---- rules (cpp/uttt_game.cpp) ---- */
void orc_init(orc_state *s);                                   /* :9-19   */
int  orc_check_win(const int b[9]);                            /* :35-61  */
int  orc_is_lose(const orc_state *s);                          /* :77-79  */
int  orc_is_draw(const orc_state *s);                          /* :82-84  */
int  orc_is_done(const orc_state *s);                          /* :87-89  */
int  orc_is_first_player(const orc_state *s);                  /* :92-94  */
void orc_next(const orc_state *s, int action, orc_state *out); /* :97-145 */
int  orc_legal_actions(const orc_state *s, int out[81]);       /* :148-191 */
void orc_to_input_tensor(const orc_state *s, float out[243]);  /* :244-280 */
int  orc_to_string(const orc_state *s, char *buf, int cap);    /* :194-241 */

/* ---- packed 32-byte form shared with the CUDA library (include/uttt_b200.h) ---- */
void orc_pack(const orc_state *s, uint32_t w[8]);
void orc_unpack(const uint32_t w[8], orc_state *s);

/* ---- counter-based RNG (Philox4x32-10) and the deterministic evaluators ---- */
void     orc_philox4x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                        uint32_t c2, uint32_t c3, uint32_t out[4]);
uint32_t orc_state_hash(const orc_state *s);
void     orc_hash_eval(const orc_state *s, float policy[81], float *value);

/* evaluator callback: fills policy[81], value for each of n states */
/* pv_mcts.py:74-180 (the reference's pure-Python search): root visit counts in legal order; returns their number */
int orc_py_mcts_counts_hash(const orc_state *root, int evaluate_count, int batch_size, int *counts_out);
int orc_py_mcts_counts_table(const orc_state *root, int evaluate_count, int batch_size, int n_entries, const uint32_t *states,
                             const float *policy, const float *value, int *counts_out, int *misses_out);
int orc_pv_mcts_scores_table(const orc_state *root, float temperature, int evaluate_count, int batch_size, int n_entries,
                             const uint32_t *states, const float *policy, const float *value, float *scores_out,
                             int *counts_out, int *misses_out);
int orc_pv_mcts_scores_hash_record(const orc_state *root, float temperature, int evaluate_count, int batch_size, int cap,
                                   uint32_t *states, float *policy, float *value, int *n_entries_out, float *scores_out);
typedef void (*orc_eval_fn)(void *ctx, const orc_state *states, int n,
                            float *policies /* n*81 */, float *values /* n */);

/* ---- search (cpp/uttt_mcts.cpp) ---- */
/* :84-196 restated literally (queue + flush); returns #scores (= #legal at root).
 * counts_out (may be NULL) receives the raw root-child visit counts. */
int  orc_pv_mcts_scores(orc_eval_fn eval, void *ctx, const orc_state *root,
                        float temperature, int evaluate_count, int batch_size,
                        float *scores_out /* 81 */, int *counts_out /* 81 */,
                        int *stats_out /* [0]=nodes,[1]=eval calls,[2]=eval states */);
int  orc_pv_mcts_scores_hash(const orc_state *root, float temperature,
                             int evaluate_count, int batch_size,
                             float *scores_out, int *counts_out, int *stats_out);
void orc_boltzman(const float *xs, int n, float temperature, float *out); /* :199-216 */

/* ---- bulk drivers used by tests / golden generation / CPU baseline ---- */
/* Random playout of game g: action at ply t = legal[ philox(seed; g, t).x % n_legal ].
 * digest = FNV-1a(64) over per-ply (action, legal mask words, active+1, main masks, status). */
void orc_playout(uint32_t seed, uint64_t game, uint64_t *digest, int *plies, int *result,
                 uint8_t *actions_out /* 81 or NULL */);
void orc_playouts(uint32_t seed, uint64_t game0, int n, uint64_t *digests, int *plies,
                  int *results);

/* One self-play game under the hash evaluator with Philox sampling
 * (self_play_cpp.py:34-101 with the sampler replaced by the counter RNG).
 * Per ply: packed state (8 words), visit counts by action id (81), action. */
int  orc_selfplay_hash(uint32_t seed, uint64_t game, int sims, int batch,
                       uint32_t *states /* 81*8 */, uint16_t *counts /* 81*81 */,
                       uint8_t *actions /* 81 */, int8_t *z /* 81 */);

/* ---- CPU cross-check of the CUDA library's THROUGHPUT search mode (csrc/tree_tp_kernels.cu) in its
 * sequential configuration: one leaf per round, no root noise.  NOT a restatement of the reference (the
 * reference has no such mode): plain PUCT with an evaluated root, single expansion per leaf,
 * correct terminal sign, same fp32 operation order as the kernels.  Returns #root children. */
int  orc_az_search_hash(const orc_state *root, int sims, int *counts_out /* 81 */);

#ifdef __cplusplus
}
#endif
#endif
