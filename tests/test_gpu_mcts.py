"""GPU: tree kernels (search + self-play loop) vs the reference goldens / the oracle under the
deterministic hash evaluator -- bit-exact float scores and visit counts."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import engine
    e = engine.Engine(n_slots=512, max_sims=800, max_batch=8, max_games=256)
    yield e
    e.close()


def test_search_scores_bit_exact_vs_reference_goldens(eng, golden_dir):
    import engine
    with np.load(os.path.join(golden_dir, "mcts.npz")) as z:
        states, cases, ref = z["states"], z["cases"], z["scores"]
    groups = {}
    for j, row in enumerate(cases):
        groups.setdefault((int(row[1]), int(row[2]), float(row[3])), []).append(j)
    checked = 0
    for (sims, batch, T), idx in groups.items():
        roots = states[[int(cases[j][0]) for j in idx]]
        for off in range(0, len(idx), eng.n_slots):
            sl = idx[off:off + eng.n_slots]
            scores, counts, ns = eng.mcts_search(roots[off:off + eng.n_slots], sims, batch, T, engine.EVAL_HASH)
            for r, j in enumerate(sl):
                n = int(cases[j][4])
                assert ns[r] == n, (sims, batch, T, j)
                if T in (0.0, 1.0):
                    assert (scores[r, :n].view(np.uint32) == ref[j, :n]).all(), (sims, batch, T, j)
                else:
                    np.testing.assert_allclose(scores[r, :n], ref[j, :n].view(np.float32), rtol=2e-6)
                if n:
                    assert counts[r, :n].sum() == sims
                checked += 1
    assert checked == len(cases)


def test_search_vs_oracle_random_states(eng):
    import engine
    sts = np.concatenate([O.playout_states(555, g)[0][:-1:3] for g in range(30)])[:500]
    for sims, batch in ((50, 8), (64, 3), (128, 8)):
        scores, counts, ns = eng.mcts_search(sts, sims, batch, 1.0, engine.EVAL_HASH)
        for i in range(0, len(sts), 7):
            sc, cn, _ = O.oracle_mcts(sts[i], 1.0, sims, batch)
            assert ns[i] == len(sc)
            assert (counts[i, :ns[i]] == cn).all()
            assert (scores[i, :ns[i]].view(np.uint32) == sc.view(np.uint32)).all()


def test_host_callback_evaluator_matches_reference_contract(golden_dir):
    """uttt_cpp.pv_mcts_scores(model=callable): same callback contract as python_bindings.cpp:11-47"""
    import uttt_cpp
    L = O.oracle()
    with np.load(os.path.join(golden_dir, "mcts.npz")) as z:
        states, cases, ref = z["states"], z["cases"], z["scores"]
    calls = []

    def model(batch):
        calls.append(len(batch))
        out = []
        for s in batch:
            pol = np.zeros(81, np.float32)
            val = C.c_float()
            L.orc_hash_eval(C.byref(O.state_from_packed(s.packed())), pol, C.byref(val))
            out.append((pol, val.value))
        return out
    done = 0
    for j, row in enumerate(cases):
        si, sims, batch, T, n = int(row[0]), int(row[1]), int(row[2]), float(row[3]), int(row[4])
        if sims > 50 or T == 0.5 or j % 23:
            continue
        calls.clear()
        sc = uttt_cpp.pv_mcts_scores(model, uttt_cpp.State._from_packed(states[si]), T, sims, batch)
        assert len(sc) == n
        assert (np.array(sc, np.float32).view(np.uint32) == ref[j, :n]).all()
        assert len(calls) == int(row[5]) and sum(calls) == int(row[6])     # same callback batching
        done += 1
    assert done > 20
    # initial position, 50 sims / batch 8: the reference's callback sizes are [8,8,8,8,8,8,2] (SURVEY Q-M3)
    calls.clear()
    uttt_cpp.pv_mcts_scores(model, uttt_cpp.State(), 1.0, 50, 8)
    assert calls == [8, 8, 8, 8, 8, 8, 2]


def test_selfplay_hash_matches_reference_goldens(eng, golden_dir):
    import engine
    with np.load(os.path.join(golden_dir, "selfplay.npz")) as z:
        g = {k: z[k] for k in z.files}
    off = 0
    for game, sims, batch, n in g["meta"]:
        h = eng.selfplay(1, sims=int(sims), batch=int(batch), seed=int(g["seed"]), evaluator=engine.EVAL_HASH,
                         game0=int(game))
        assert h.lens[0] == n
        assert (h.states[0, :n] == g["states"][off:off + n]).all()
        assert (h.counts[0, :n] == g["counts"][off:off + n]).all()
        assert (h.actions[0, :n] == g["actions"][off:off + n]).all()
        st, cn, z_ = h.samples()
        assert (z_ == g["z"][off:off + n]).all()
        assert h.stats[0] == n and h.stats[1] == n * sims
        off += n


@pytest.mark.parametrize("slots", [48, 160])      # 160 slots: two overlapped lanes on two streams
def test_selfplay_many_games_with_slot_recycling_vs_oracle(slots):
    import engine
    os.environ["UTTT_LANE_THRESHOLD"] = "100"
    e = engine.Engine(n_slots=slots, max_sims=50, max_batch=8, max_games=200)
    del os.environ["UTTT_LANE_THRESHOLD"]
    try:
        h = e.selfplay(200, sims=50, batch=8, seed=9, evaluator=engine.EVAL_HASH, game0=1000)
        assert (h.lens > 0).all() and h.stats[0] == h.lens.sum()
        L = O.oracle()
        for gi in (0, 47, 48, 131, 199):
            st = np.zeros((81, 8), np.uint32); cn = np.zeros((81, 81), np.uint16)
            ac = np.zeros(81, np.uint8); z = np.zeros(81, np.int8)
            n = L.orc_selfplay_hash(9, 1000 + gi, 50, 8, st, cn, ac, z)
            assert h.lens[gi] == n
            assert (h.states[gi, :n] == st[:n]).all() and (h.counts[gi, :n] == cn[:n]).all()
            assert (h.actions[gi, :n] == ac[:n]).all()
            assert h.final[gi] == (1 if z[0] == -1 else 0)
        # running it again gives the identical history (slot assignment does not matter)
        h2 = e.selfplay(200, sims=50, batch=8, seed=9, evaluator=engine.EVAL_HASH, game0=1000)
        assert (h2.lens == h.lens).all() and (h2.actions == h.actions).all() and (h2.counts == h.counts).all()
    finally:
        e.close()


def test_boltzman_device(golden_dir):
    import uttt_cpp
    with np.load(os.path.join(golden_dir, "boltzman.npz")) as z:
        g = {k: z[k] for k in z.files}
    for i in range(len(g["xs"])):
        out = np.array(uttt_cpp.boltzman(g["xs"][i].tolist(), 1.0), np.float32)
        assert (out.view(np.uint32) == g["T1"][i].view(np.uint32)).all()
        for T in (0.5, 2.0):
            out = np.array(uttt_cpp.boltzman(g["xs"][i].tolist(), T), np.float32)
            np.testing.assert_allclose(out, g["T%g" % T][i], rtol=2e-6, atol=1e-9)


# ------------------------------------------------------------------ throughput mode (not in the reference)
def test_throughput_search_sequential_matches_cpu_cross_check(eng):
    """m = 1 leaf per round, no root noise: plain PUCT with an evaluated root -- bit-exact visit counts vs
    oracle/uttt_oracle.c:orc_az_search_hash under the hash evaluator"""
    import engine
    eng.set_root_noise(0.3, 0.0)
    sts = np.concatenate([O.playout_states(808, g)[0][::4] for g in range(24)])[:300]
    for sims in (50, 200):
        scores, counts, ns = eng.mcts_search(sts, sims, 1, 1.0, engine.EVAL_HASH, flags=engine.SP_THROUGHPUT)
        for i in range(0, len(sts), 5):
            cn = O.oracle_az_search(sts[i], sims)
            assert ns[i] == len(cn)
            assert (counts[i, :ns[i]] == cn).all(), (sims, i)
    eng.set_root_noise(0.3, 0.25)


def test_throughput_search_invariants_with_virtual_loss_and_noise(eng):
    import engine
    sts = np.concatenate([O.playout_states(909, g)[0][:-1:5] for g in range(20)])[:200]
    for m in (4, 8):
        eng.set_root_noise(0.3, 0.25)
        s1, c1, n1 = eng.mcts_search(sts, 200, m, 1.0, engine.EVAL_HASH, flags=engine.SP_THROUGHPUT)
        s2, c2, n2 = eng.mcts_search(sts, 200, m, 1.0, engine.EVAL_HASH, flags=engine.SP_THROUGHPUT)
        assert (c1 == c2).all()                                   # reproducible: depends only on (seed, m)
        for i in range(len(sts)):
            _, legal, _ = O.oracle_probe(sts[i])
            assert n1[i] == len(legal)
            assert c1[i, :n1[i]].sum() == 200 and (c1[i, n1[i]:] == 0).all()      # every simulation lands on a root child
        eng.set_root_noise(0.3, 0.0)
        s3, c3, n3 = eng.mcts_search(sts, 200, m, 1.0, engine.EVAL_HASH, flags=engine.SP_THROUGHPUT)
        assert (c3 != c1).any()                                   # the noise does change the search
    eng.set_root_noise(0.3, 0.25)


def test_throughput_selfplay_runs_and_is_reproducible():
    import engine
    e = engine.Engine(n_slots=64, max_sims=200, max_batch=8, max_games=96)
    try:
        h1 = e.selfplay(96, sims=200, batch=8, seed=5, evaluator=engine.EVAL_HASH, flags=engine.SP_THROUGHPUT)
        st, cn, z = h1.samples()
        assert (h1.lens >= 17).all() and (cn.sum(1) == 200).all()
        h2 = e.selfplay(96, sims=200, batch=8, seed=5, evaluator=engine.EVAL_HASH, flags=engine.SP_THROUGHPUT)
        assert (h1.actions == h2.actions).all() and (h1.counts == h2.counts).all() and (h1.lens == h2.lens).all()
        h3 = e.selfplay(96, sims=200, batch=8, seed=6, evaluator=engine.EVAL_HASH, flags=engine.SP_THROUGHPUT)
        assert (h3.actions != h1.actions).any()
    finally:
        e.close()


# ------------------------------------------------------------------ edge cases
def test_empty_and_finished_inputs(eng, golden_dir):
    import engine
    # no roots
    s, c, n = eng.mcts_search(np.zeros((0, 8), np.uint32), 50, 8, 1.0, engine.EVAL_HASH)
    assert s.shape == (0, 81) and n.shape == (0,)
    # finished positions give an empty score vector (cpp/uttt_mcts.cpp:96-98), mixed with live ones
    finals = np.stack([O.playout_states(11, g)[0][-1] for g in range(6)])
    live = O.playout_states(11, 0)[0][3:5]
    roots = np.concatenate([finals[:3], live, finals[3:]])
    s, c, n = eng.mcts_search(roots, 50, 8, 1.0, engine.EVAL_HASH)
    assert (n[:3] == 0).all() and (n[5:] == 0).all() and (n[3:5] > 0).all()
    for i in (3, 4):
        sc, cn, _ = O.oracle_mcts(roots[i], 1.0, 50, 8)
        assert (s[i, :n[i]].view(np.uint32) == sc.view(np.uint32)).all()
    s2, c2, n2 = eng.mcts_search(roots, 50, 4, 1.0, engine.EVAL_HASH, flags=engine.SP_THROUGHPUT)
    assert (n2 == n).all()
    # zero games
    h = eng.selfplay(0, sims=50, batch=8, seed=1, evaluator=engine.EVAL_HASH)
    assert h.stats[0] == 0
    # argument validation surfaces as errors, not crashes
    with pytest.raises(RuntimeError):
        eng.mcts_search(live, 5000, 8, 1.0, engine.EVAL_HASH)          # > max_sims
    with pytest.raises(RuntimeError):
        eng.mcts_search(live, 50, 64, 1.0, engine.EVAL_HASH)           # > max_batch
    with pytest.raises(RuntimeError):
        eng.selfplay(10 ** 6, sims=50, batch=8, seed=1, evaluator=engine.EVAL_HASH, history=engine.History(1))


def test_largest_configuration_800_sims_matches_oracle(eng):
    """800 simulations / batch 8 (BASELINE config 4's search depth) on a handful of positions"""
    import engine
    sts = O.playout_states(2024, 3)[0][2:30:6]
    scores, counts, ns = eng.mcts_search(sts, 800, 8, 1.0, engine.EVAL_HASH)
    for i in range(len(sts)):
        sc, cn, st = O.oracle_mcts(sts[i], 1.0, 800, 8)
        assert (counts[i, :ns[i]] == cn).all() and (scores[i, :ns[i]].view(np.uint32) == sc.view(np.uint32)).all()


def test_full_size_cycle_500_games_every_game_bit_exact():
    """BASELINE config 3 at full size under the hash evaluator: all 500 games of a self-play cycle (states, visit
    counts, sampled actions, lengths, labels) are identical to the CPU oracle's"""
    import engine
    e = engine.Engine(n_slots=500, max_sims=50, max_batch=8, max_games=500)
    try:
        h = e.selfplay(500, sims=50, batch=8, seed=2025, evaluator=engine.EVAL_HASH, game0=0)
        L = O.oracle()
        st = np.zeros((81, 8), np.uint32); cn = np.zeros((81, 81), np.uint16)
        ac = np.zeros(81, np.uint8); z = np.zeros(81, np.int8)
        total = 0
        for g in range(500):
            n = L.orc_selfplay_hash(2025, g, 50, 8, st, cn, ac, z)
            assert h.lens[g] == n, g
            assert (h.states[g, :n] == st[:n]).all() and (h.counts[g, :n] == cn[:n]).all(), g
            assert (h.actions[g, :n] == ac[:n]).all() and h.final[g] == (1 if z[0] == -1 else 0), g
            total += n
        assert h.stats[0] == total and h.stats[1] == 50 * total
    finally:
        e.close()
