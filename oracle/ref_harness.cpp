// ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C-ABI harness around the UNMODIFIED reference sources (compiled where they lie
// under /root/reference/cpp by oracle/Makefile; output only into oracle/_ref/).
// It drives the reference's own UTTT::State / UTTT::pv_mcts_scores with the same
// deterministic inputs the oracle and the CUDA library use (Philox playouts, the
// integer-hash evaluator), so that
//   * oracle/uttt_oracle.c can be pinned against the real reference, and
//   * golden vectors can be generated (oracle/gen_golden.py -> tests/golden/).
// Nothing here is reference code; the reference is only #included and linked.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "uttt_game.h"   // /root/reference/cpp (via -I)
#include "uttt_mcts.h"
#include "uttt_oracle.h"

namespace {

UTTT::State from_orc(const orc_state &o) {
    std::array<std::array<int, 9>, 9> p{}, e{};
    std::array<int, 9> mp{}, me{};
    for (int b = 0; b < 9; b++) {
        for (int c = 0; c < 9; c++) { p[b][c] = o.pieces[b][c]; e[b][c] = o.enemy[b][c]; }
        mp[b] = o.main_pieces[b];
        me[b] = o.main_enemy[b];
    }
    return UTTT::State(p, e, mp, me, o.active);
}

orc_state to_orc(const UTTT::State &s) {
    orc_state o;
    for (int b = 0; b < 9; b++) {
        for (int c = 0; c < 9; c++) {
            o.pieces[b][c] = s.get_pieces()[b][c];
            o.enemy[b][c] = s.get_enemy_pieces()[b][c];
        }
        o.main_pieces[b] = s.get_main_board_pieces()[b];
        o.main_enemy[b] = s.get_main_board_enemy_pieces()[b];
    }
    o.active = s.get_active_board();
    return o;
}

UTTT::State from_packed(const uint32_t w[8]) {
    orc_state o;
    orc_unpack(w, &o);
    return from_orc(o);
}

void to_packed(const UTTT::State &s, uint32_t w[8]) {
    orc_state o = to_orc(s);
    orc_pack(&o, w);
}

inline uint64_t fnv64(uint64_t h, uint32_t x) { return (h ^ (uint64_t)x) * 0x100000001B3ull; }

std::vector<UTTT::InferenceResult> hash_model(const std::vector<UTTT::State> &states) {
    std::vector<UTTT::InferenceResult> out;
    for (const auto &s : states) {
        orc_state o = to_orc(s);
        UTTT::InferenceResult r;
        r.policy.resize(81);
        orc_hash_eval(&o, r.policy.data(), &r.value);
        out.push_back(r);
    }
    return out;
}

long g_eval_calls = 0, g_eval_states = 0;
std::vector<UTTT::InferenceResult> counting_hash_model(const std::vector<UTTT::State> &states) {
    g_eval_calls++;
    g_eval_states += (long)states.size();
    return hash_model(states);
}

}  // namespace

extern "C" {

// Same digest definition as orc_playout (oracle/uttt_oracle.c), computed with the reference's State.
void ref_playouts(uint32_t seed, uint64_t game0, int n, uint64_t *digests, int *plies, int *results) {
    for (int g = 0; g < n; g++) {
        uint64_t game = game0 + (uint64_t)g;
        UTTT::State s;
        uint64_t h = 0xCBF29CE484222325ull;
        int t = 0;
        for (;;) {
            std::vector<int> legal = s.legal_actions();
            if (s.is_lose() || legal.empty()) break;
            uint32_t r[4], lm[3] = {0, 0, 0}, w[8];
            orc_philox4x32(seed, 0u, (uint32_t)game, (uint32_t)(game >> 32), (uint32_t)t, 0u, r);
            int a = legal[r[0] % (uint32_t)legal.size()];
            for (int x : legal) lm[x / 27] |= 1u << (x % 27);
            to_packed(s, w);
            h = fnv64(h, (uint32_t)a);
            h = fnv64(h, lm[0]); h = fnv64(h, lm[1]); h = fnv64(h, lm[2]);
            h = fnv64(h, w[6]);
            s = s.next(a);
            t++;
        }
        uint32_t w[8];
        to_packed(s, w);
        bool lose = s.is_lose();
        for (int i = 0; i < 7; i++) h = fnv64(h, w[i]);
        h = fnv64(h, lose ? 1u : 2u);
        digests[g] = h;
        plies[g] = t;
        results[g] = lose ? (s.is_first_player() ? 2 : 1) : 0;
    }
}

// Per-state probe: everything the rules expose, for one packed state.
// flags: bit0 is_lose, bit1 is_draw, bit2 is_done, bit3 is_first_player
void ref_state_probe(const uint32_t w[8], int *flags, int *n_legal, int legal[81], float tensor[243]) {
    UTTT::State s = from_packed(w);
    *flags = (s.is_lose() ? 1 : 0) | (s.is_draw() ? 2 : 0) | (s.is_done() ? 4 : 0) |
             (s.is_first_player() ? 8 : 0);
    std::vector<int> l = s.legal_actions();
    *n_legal = (int)l.size();
    for (size_t i = 0; i < l.size(); i++) legal[i] = l[i];
    std::vector<float> t = s.to_input_tensor();
    std::memcpy(tensor, t.data(), sizeof(float) * 243);
}

void ref_state_next(const uint32_t w[8], int action, uint32_t out[8]) {
    UTTT::State s = from_packed(w).next(action);
    to_packed(s, out);
}

int ref_state_to_string(const uint32_t w[8], char *buf, int cap) {
    std::string str = from_packed(w).to_string();
    int n = (int)str.size();
    if (cap > 0) {
        int m = n < cap - 1 ? n : cap - 1;
        std::memcpy(buf, str.data(), (size_t)m);
        buf[m] = 0;
    }
    return n;
}

// UTTT::pv_mcts_scores under the hash evaluator. stats: [0]=eval calls, [1]=eval states
int ref_mcts_scores_hash(const uint32_t w[8], float temperature, int evaluate_count, int batch_size,
                         float *scores_out, int *stats_out) {
    g_eval_calls = g_eval_states = 0;
    std::vector<float> sc = UTTT::pv_mcts_scores(counting_hash_model, from_packed(w), temperature,
                                                 evaluate_count, batch_size);
    for (size_t i = 0; i < sc.size(); i++) scores_out[i] = sc[i];
    if (stats_out) { stats_out[0] = (int)g_eval_calls; stats_out[1] = (int)g_eval_states; }
    return (int)sc.size();
}

// UTTT::pv_mcts_scores fed recorded (leaf state -> policy, value) rows (see orc_pv_mcts_scores_table): entries are consumed
// in order, the k queued copies of a leaf inside one callback share one entry.  *misses_out = leaves without an unused
// entry + (entries left unused << 16).
int ref_mcts_scores_table(const uint32_t w[8], float temperature, int evaluate_count, int batch_size, int n_entries,
                          const uint32_t *states, const float *policy, const float *value, float *scores_out,
                          int *misses_out) {
    std::vector<uint8_t> used((size_t)n_entries, 0);
    int misses = 0;
    auto model = [&](const std::vector<UTTT::State> &batch) {
        std::vector<UTTT::InferenceResult> out;
        int last = -1;
        uint32_t lastw[8] = {0};
        for (const auto &s : batch) {
            uint32_t pw[8];
            to_packed(s, pw);
            int e = -1;
            if (last >= 0 && std::memcmp(pw, lastw, 28) == 0) e = last;
            else
                for (int j = 0; j < n_entries; j++)
                    if (!used[j] && std::memcmp(pw, states + 8 * j, 28) == 0) { e = j; used[j] = 1; break; }
            UTTT::InferenceResult r;
            if (e < 0) {
                misses++;
                r.policy.assign(81, 1.0f / 81.0f);
                r.value = 0.0f;
            } else {
                r.policy.assign(policy + 81 * e, policy + 81 * e + 81);
                r.value = value[e];
                last = e;
                std::memcpy(lastw, pw, sizeof(pw));
            }
            out.push_back(r);
        }
        return out;
    };
    std::vector<float> sc = UTTT::pv_mcts_scores(model, from_packed(w), temperature, evaluate_count, batch_size);
    for (size_t i = 0; i < sc.size(); i++) scores_out[i] = sc[i];
    int unused = 0;
    for (int j = 0; j < n_entries; j++) unused += !used[j];
    if (misses_out) *misses_out = misses + (unused << 16);
    return (int)sc.size();
}

void ref_boltzman(const float *xs, int n, float temperature, float *out) {
    std::vector<float> v(xs, xs + n);
    std::vector<float> r = UTTT::boltzman(v, temperature);
    for (int i = 0; i < n; i++) out[i] = r[i];
}

// One self-play game: reference search (T=1, hash evaluator), counts = round(score*sims),
// Philox sampling identical to orc_selfplay_hash.
int ref_selfplay_hash(uint32_t seed, uint64_t game, int sims, int batch, uint32_t *states,
                      uint16_t *counts, uint8_t *actions, int8_t *z) {
    UTTT::State s;
    int t = 0;
    while (!s.is_done()) {
        std::vector<int> legal = s.legal_actions();
        std::vector<float> sc = UTTT::pv_mcts_scores(hash_model, s, 1.0f, sims, batch);
        int nl = (int)legal.size();
        std::vector<int> cnt(nl);
        int total = 0;
        for (int i = 0; i < nl; i++) { cnt[i] = (int)lrintf(sc[i] * (float)sims); total += cnt[i]; }
        uint32_t r[4];
        orc_philox4x32(seed, 1u, (uint32_t)game, (uint32_t)(game >> 32), (uint32_t)t, 0u, r);
        uint32_t pick = (uint32_t)(((uint64_t)r[0] * (uint64_t)total) >> 32);
        int idx = 0, acc = 0;
        for (int i = 0; i < nl; i++) { acc += cnt[i]; if ((uint32_t)acc > pick) { idx = i; break; } }
        to_packed(s, states + 8 * t);
        for (int a = 0; a < 81; a++) counts[81 * t + a] = 0;
        for (int i = 0; i < nl; i++) counts[81 * t + legal[i]] = (uint16_t)cnt[i];
        actions[t] = (uint8_t)legal[idx];
        s = s.next(legal[idx]);
        t++;
    }
    int value = s.is_lose() ? -1 : 0;
    for (int i = 0; i < t; i++) { z[i] = (int8_t)value; value = -value; }
    return t;
}

// CPU-baseline helpers (timed by bench.py): search with a null evaluator (uniform pi, v=0.1)
static std::vector<UTTT::InferenceResult> null_model(const std::vector<UTTT::State> &states) {
    std::vector<UTTT::InferenceResult> out(states.size());
    for (auto &r : out) { r.policy.assign(81, 1.0f / 81.0f); r.value = 0.1f; }
    return out;
}
long ref_selfplay_null(uint32_t seed, int n_games, int sims, int batch) {
    long plies = 0;
    for (int g = 0; g < n_games; g++) {
        UTTT::State s;
        int t = 0;
        while (!s.is_done()) {
            std::vector<int> legal = s.legal_actions();
            std::vector<float> sc = UTTT::pv_mcts_scores(null_model, s, 1.0f, sims, batch);
            uint32_t r[4];
            orc_philox4x32(seed, 2u, (uint32_t)g, 0u, (uint32_t)t, 0u, r);
            float u = (float)(r[0] >> 8) * (1.0f / 16777216.0f), acc = 0.0f;
            int idx = (int)legal.size() - 1;
            for (size_t i = 0; i < sc.size(); i++) { acc += sc[i]; if (u < acc) { idx = (int)i; break; } }
            s = s.next(legal[idx]);
            t++;
        }
        plies += t;
    }
    return plies;
}

}  // extern "C"
