"""GPU: API completeness of the drop-in modules (VERDICT r1 "missing" 3 / 5): SP_TEMPERATURE != 1,
the per-game progress line, the uploaded-weights cache of uttt_cpp.pv_mcts_scores, the numerics option."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


def _fresh_scores(model, state, numerics):
    import engine
    e = engine.Engine(n_slots=4, max_sims=50, max_batch=8, max_games=1)
    try:
        e.upload_model(model)
        sc, _, ns = e.mcts_search(state.packed().reshape(1, 8), 50, 8, 1.0, engine.evaluator_of(numerics))
        return sc[0, :ns[0]].astype(np.float64)
    finally:
        e.close()


def test_weights_cache_follows_the_model_object_and_its_contents():
    """load A / del / load B re-uses A's address and version counters: B must not be searched with A's weights"""
    import gc
    import torch
    import uttt_cpp
    import pv_mcts_cpp
    from dual_network import DualNetwork
    s = uttt_cpp.State().next(40).next(37)
    seen = []
    for seed in range(6):
        torch.manual_seed(seed)
        model = DualNetwork().eval()
        got = pv_mcts_cpp.pv_mcts_scores_cpp(model, s, 1.0, 50, 8)
        assert (got == _fresh_scores(model, s, uttt_cpp.NUMERICS)).all(), seed
        seen.append(got)
        with torch.no_grad():                                  # in place, through .data: no version counter moves
            model.policy_fc.weight.data.mul_(-1.0)
            model.conv_input.weight.data.mul_(0.5)
        got2 = pv_mcts_cpp.pv_mcts_scores_cpp(model, s, 1.0, 50, 8)
        assert (got2 == _fresh_scores(model, s, uttt_cpp.NUMERICS)).all(), seed
        del model
        gc.collect()
    assert any((a != seen[0]).any() for a in seen[1:])          # the six networks do search differently
    assert uttt_cpp.NUMERICS == "bf16x3"                         # the conforming numerics are the default ...
    for numerics in ("bf16", "fp32"):                           # ... and the other two are one assignment away
        torch.manual_seed(1)
        model = DualNetwork().eval()
        old = uttt_cpp.NUMERICS
        try:
            uttt_cpp.NUMERICS = numerics
            got = pv_mcts_cpp.pv_mcts_scores_cpp(model, s, 1.0, 50, 8)
        finally:
            uttt_cpp.NUMERICS = old
        assert (got == _fresh_scores(model, s, numerics)).all()
    with pytest.raises(ValueError):
        import engine
        engine.evaluator_of("fp8")


def _warp_scan_pick(counts, inv_t, u):
    """numpy restatement of sample_move_temperature (csrc/tree_common.cuh): fp32 weights n^(1/T), Hillis-Steele inclusive
    scan per 32 children + carry, first prefix sum above u * total"""
    w = np.power(counts.astype(np.float32), np.float32(inv_t), dtype=np.float32)
    pref = np.zeros(len(w), np.float32)
    carry = np.float32(0)
    for base in range(0, len(w), 32):
        incl = np.zeros(32, np.float32)
        incl[:len(w[base:base + 32])] = w[base:base + 32]
        for off in (1, 2, 4, 8, 16):
            sh = np.concatenate([np.zeros(off, np.float32), incl[:-off]])
            incl = np.where(np.arange(32) >= off, incl + sh, incl).astype(np.float32)
        incl = (incl + carry).astype(np.float32)
        pref[base:base + 32] = incl[:len(w[base:base + 32])]
        carry = incl[31]
    thr = np.float32(u) * carry
    hit = np.nonzero(pref > thr)[0]
    return int(hit[0]) if len(hit) else int(np.nonzero(counts > 0)[0][-1])


def _same_games(a, b):
    """the history rows beyond a game's length are stale (never cleared between runs): compare the played plies only"""
    if not (a.lens == b.lens).all():
        return False
    mask = np.arange(81)[None, :] < a.lens[:, None]
    return bool((a.actions[mask] == b.actions[mask]).all() and (a.counts[mask] == b.counts[mask]).all())


def test_selfplay_temperature_on_device():
    """SP_TEMPERATURE passes through to the search's scores in the reference (self_play_cpp.py:27,62 ->
    cpp/uttt_mcts.cpp:183-216); here the device sampler takes it: T = 0 plays the first maximum of the visit counts,
    T = 0.5 / 2 draw from n^(1/T) with the documented Philox draw"""
    import engine
    e = engine.Engine(n_slots=64, max_sims=50, max_batch=8, max_games=64)
    try:
        base = e.selfplay(64, sims=50, batch=8, seed=9, evaluator=engine.EVAL_HASH)
        e.set_selfplay_temperature(0.0)
        h0 = e.selfplay(64, sims=50, batch=8, seed=9, evaluator=engine.EVAL_HASH)
        for g in range(64):
            for t in range(int(h0.lens[g])):
                legal = O.oracle_probe(h0.states[g, t])[1]
                cn = h0.counts[g, t][legal]
                assert h0.actions[g, t] == legal[int(np.argmax(cn))], (g, t)
        for T in (0.5, 2.0):
            e.set_selfplay_temperature(T)
            h = e.selfplay(64, sims=50, batch=8, seed=9, evaluator=engine.EVAL_HASH)
            h_again = e.selfplay(64, sims=50, batch=8, seed=9, evaluator=engine.EVAL_HASH)
            assert _same_games(h, h_again)
            agree = n = 0
            r = np.zeros(4, np.uint32)
            for g in range(64):
                for t in range(int(h.lens[g])):
                    legal = O.oracle_probe(h.states[g, t])[1]
                    cn = h.counts[g, t][legal]
                    O.oracle().orc_philox4x32(9, 1, g, 0, t, 0, r)
                    u = np.float32(int(r[0]) >> 8) * np.float32(1.0 / 16777216.0)
                    agree += int(h.actions[g, t] == legal[_warp_scan_pick(cn, 1.0 / T, u)])
                    n += 1
                    assert cn[list(legal).index(h.actions[g, t])] > 0
            assert n > 64 * 17 and agree >= 0.995 * n, (T, agree, n)      # powf may differ from numpy's in the last bit
            assert not _same_games(h, base)
        e.set_selfplay_temperature(1.0)
        h1 = e.selfplay(64, sims=50, batch=8, seed=9, evaluator=engine.EVAL_HASH)
        assert _same_games(h1, base)
        with pytest.raises(RuntimeError):
            e.set_selfplay_temperature(-1.0)
    finally:
        e.close()


def test_self_play_temperature_and_progress_lines(tmp_path, monkeypatch, capsys):
    """self_play(): one 'SelfPlay i/N (Backend: C++)' line per finished game as in self_play_cpp.py:121, and the policy
    targets are the search's scores at SP_TEMPERATURE (self_play_cpp.py:62-83)"""
    import pickle
    import torch
    import self_play_cpp
    from dual_network import dual_network
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(1)
    dual_network()
    monkeypatch.setattr(self_play_cpp, "SP_GAME_COUNT", 24)
    monkeypatch.setattr(self_play_cpp, "SP_TEMPERATURE", 0.5)
    np.random.seed(3)
    capsys.readouterr()
    path = self_play_cpp.self_play()
    out = capsys.readouterr().out
    lines = [x for x in out.replace("\n", "\r").split("\r") if x.startswith("SelfPlay")]
    assert lines == ["SelfPlay %d/24 (Backend: C++)" % (i + 1) for i in range(24)], lines[:3]
    with open(path, "rb") as f:
        history = pickle.load(f)
    pis = np.array([h[1] for h in history])
    assert np.allclose(pis.sum(1), 1.0) and len(history) == self_play_cpp.last_stats["plies"]
    # T = 0.5 squares the visit counts: sqrt(pi) * const must be integers summing to 50
    root = np.sqrt(pis)
    cn = root / root.sum(1, keepdims=True) * 50
    assert np.abs(cn - np.round(cn)).max() < 1e-3
    monkeypatch.setattr(self_play_cpp, "SP_TEMPERATURE", 0)
    np.random.seed(3)
    hist = self_play_cpp.play(torch.nn.Module.eval(_best()))
    assert all(sorted(h[1].tolist())[-2:] == [0.0, 1.0] for h in hist)


def _best():
    import torch
    from dual_network import DualNetwork
    m = DualNetwork()
    m.load_state_dict(torch.load("./model/best.pth", weights_only=True))
    return m


@pytest.mark.parametrize("alpha,L", [(0.3, 9), (0.3, 81), (1.0, 5), (0.1, 20), (2.5, 3)])
def test_dirichlet_root_noise_distribution(alpha, L):
    """the throughput mode's root noise (csrc/tree_tp_kernels.cu: Marsaglia-Tsang gammas from Philox, normalised) is
    Dirichlet(alpha): moments against the closed forms, every marginal's mean, and a Kolmogorov-Smirnov test of one
    marginal against Beta(alpha, (L-1) alpha) and against numpy's own sampler"""
    import engine
    from scipy import stats
    n = 200000
    x = engine.dirichlet_samples(17, 0, n, L, alpha).cpu().numpy().astype(np.float64)
    assert np.isfinite(x).all() and (x >= 0).all() and np.abs(x.sum(1) - 1).max() < 1e-5
    mean, var = 1.0 / L, (1.0 / L) * (1 - 1.0 / L) / (L * alpha + 1)
    se_mean = np.sqrt(var / n)
    assert np.abs(x.mean(0) - mean).max() < 6 * se_mean, (x.mean(0), mean)
    assert abs(x[:, 0].var() - var) < 0.03 * var
    cov01 = -(1.0 / L) ** 2 / (L * alpha + 1)                    # Cov(x_i, x_j) of a symmetric Dirichlet
    assert abs(np.cov(x[:, 0], x[:, 1])[0, 1] - cov01) < 0.05 * abs(cov01) + 6 * var / np.sqrt(n)
    for col in (0, L - 1):
        d, p = stats.kstest(x[:, col], stats.beta(alpha, (L - 1) * alpha).cdf)
        assert p > 1e-4 or d < 0.004, (col, d, p)
    ref = np.random.RandomState(1).dirichlet([alpha] * L, size=n)[:, 0]
    d2, p2 = stats.ks_2samp(x[:, 0], ref)
    assert p2 > 1e-4 or d2 < 0.005, (d2, p2)
    # a different seed / game range gives different, equally distributed draws; the same key the same draw
    y = engine.dirichlet_samples(17, 0, 64, L, alpha).cpu().numpy()
    z = engine.dirichlet_samples(18, 0, 64, L, alpha).cpu().numpy()
    assert (y == x[:64].astype(np.float32)).all() and (y != z).any()


def test_reference_test_script_flow_with_a_python_inference_closure():
    """The reference's own smoke test of its native module (test_cpp_mcts.py:22-101), step by step through the drop-in
    modules: State(), legal_actions(), to_input_tensor(), then uttt_cpp.pv_mcts_scores(model=<Python closure over a
    DualNetwork>, state, temperature=1.0, evaluate_count=10, batch_size=2).  Beyond the script's own checks (243 floats,
    one score per legal action, scores sum to 1) the closure's outputs are recorded and the reference search fed exactly
    those rows (cpp/python_bindings.cpp:11-47 contract) must return the same float scores bit for bit."""
    import torch
    import uttt_cpp
    from dual_network import DualNetwork
    torch.manual_seed(0)
    model = DualNetwork().eval()
    state = uttt_cpp.State()
    legal = state.legal_actions()
    assert len(legal) == 81 and legal[:5] == [0, 1, 2, 3, 4]                     # test_cpp_mcts.py:24-29
    tensor = state.to_input_tensor()
    assert len(tensor) == 243 and np.array(tensor, np.float32).reshape(9, 9, 3)[..., 2].sum() == 81     # :32-36
    seen_states, seen_pol, seen_val, calls = [], [], [], []

    def inference(states_list):                                                  # test_cpp_mcts.py:40-66
        calls.append(len(states_list))
        x = np.stack([np.array(s.to_input_tensor(), np.float32).reshape(9, 9, 3) for s in states_list]).transpose(0, 3, 1, 2)
        with torch.no_grad():
            policies, values = model(torch.from_numpy(np.ascontiguousarray(x)))
        out = []
        for i, s in enumerate(states_list):
            pol, val = policies[i].numpy().copy(), float(values[i][0])
            seen_states.append(s.packed().copy()); seen_pol.append(pol); seen_val.append(np.float32(val))
            out.append((pol, val))
        return out
    one = inference([state])                                                     # :69-73
    assert abs(float(one[0][0].sum()) - 1.0) < 1e-5 and -1.0 <= one[0][1] <= 1.0
    seen_states.clear(); seen_pol.clear(); seen_val.clear(); calls.clear()
    scores = uttt_cpp.pv_mcts_scores(model=inference, state=state, temperature=1.0, evaluate_count=10, batch_size=2)   # :78-83
    assert isinstance(scores, list) and len(scores) == len(legal)                # :96-99
    assert abs(sum(scores) - 1.0) < 1e-6 and calls == [2, 2, 2, 2, 2]
    # the reference queues the same leaf k times (Q-M3): one row per distinct evaluation is what the table replay consumes
    keep = np.cumsum([0] + calls[:-1])
    want, miss, unused = O.table_mcts(state.packed(), 1.0, 10, 2, np.array(seen_states)[keep], np.array(seen_pol)[keep],
                                      np.array(seen_val)[keep])
    assert miss == 0 and unused == 0
    assert np.array(scores, np.float32).tobytes() == want.tobytes()
