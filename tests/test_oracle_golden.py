"""CPU: pin oracle/uttt_oracle.c against the golden vectors minted from the compiled reference
(oracle/gen_golden.py), and -- where oracle/_ref exists -- against the reference directly."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O


@pytest.fixture(scope="module")
def rules(golden_dir):
    with np.load(os.path.join(golden_dir, "rules.npz")) as z:
        return {k: z[k] for k in z.files}        # decompress once


def test_playout_digests_match_reference(rules):
    L = O.oracle()
    n = len(rules["digests"])
    dg = np.zeros(n, np.uint64); pl = np.zeros(n, np.int32); rs = np.zeros(n, np.int32)
    L.orc_playouts(int(rules["seed"]), 0, n, dg, pl, rs)
    assert (dg == rules["digests"]).all()
    assert (pl == rules["plies"]).all()
    assert (rs == rules["results"]).all()
    assert pl.min() >= 17 and pl.max() <= 81


def test_state_probes_match_reference(rules):
    L = O.oracle()
    for i, w in enumerate(rules["states"]):
        flags, legal, tens = O.oracle_probe(w)
        assert flags == rules["flags"][i]
        n = rules["n_legal"][i]
        assert len(legal) == n and (legal == rules["legal"][i, :n]).all()
        assert (tens == rules["tensor"][i].astype(np.float32)).all()
        s = O.state_from_packed(w)
        for k in range(n):
            nx = O.OrcState()
            L.orc_next(C.byref(s), int(legal[k]), C.byref(nx))
            assert (O.packed_from_state(nx) == rules["next"][i, k]).all()
        buf = C.create_string_buffer(2048)
        L.orc_to_string(C.byref(s), buf, 2048)
        assert buf.value.decode() == str(rules["strings"][i])


def test_next_without_legality_check(rules):
    L = O.oracle()
    for w, a, ref in zip(rules["ill_states"], rules["ill_actions"], rules["ill_next"]):
        s = O.state_from_packed(w)
        nx = O.OrcState()
        L.orc_next(C.byref(s), int(a), C.byref(nx))
        assert (O.packed_from_state(nx) == ref).all()


def test_pack_roundtrip(rules):
    for w in rules["states"][::17]:
        assert (O.packed_from_state(O.state_from_packed(w)) == w).all()


def test_mcts_scores_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "mcts.npz"))
    states, cases, scores = g["states"], g["cases"], g["scores"]
    for row, ref_bits in zip(cases, scores):
        si, sims, batch, T, n = int(row[0]), int(row[1]), int(row[2]), float(row[3]), int(row[4])
        sc, cn, st = O.oracle_mcts(states[si], T, sims, batch)
        assert len(sc) == n
        if T in (0.0, 1.0):
            assert (sc.view(np.uint32) == ref_bits[:n]).all(), (si, sims, batch, T)
        else:
            np.testing.assert_allclose(sc, ref_bits[:n].view(np.float32), rtol=1e-6)
        if n:
            assert cn.sum() == sims
            assert st[1] == int(row[5]) and st[2] == int(row[6])


def test_selfplay_hash_matches_reference(golden_dir):
    with np.load(os.path.join(golden_dir, "selfplay.npz")) as z:
        g = {k: z[k] for k in z.files}
    L = O.oracle()
    off = 0
    for game, sims, batch, n in g["meta"]:
        st = np.zeros((81, 8), np.uint32); cn = np.zeros((81, 81), np.uint16)
        ac = np.zeros(81, np.uint8); z = np.zeros(81, np.int8)
        m = L.orc_selfplay_hash(int(g["seed"]), int(game), int(sims), int(batch), st, cn, ac, z)
        assert m == n
        assert (st[:n] == g["states"][off:off + n]).all()
        assert (cn[:n] == g["counts"][off:off + n]).all()
        assert (ac[:n] == g["actions"][off:off + n]).all()
        assert (z[:n] == g["z"][off:off + n]).all()
        off += n


def test_boltzman(golden_dir):
    g = np.load(os.path.join(golden_dir, "boltzman.npz"))
    L = O.oracle()
    xs = g["xs"]
    for T in (1.0, 0.5, 2.0):
        for i in range(len(xs)):
            o = np.zeros(xs.shape[1], np.float32)
            L.orc_boltzman(np.ascontiguousarray(xs[i]), xs.shape[1], T, o)
            assert (o.view(np.uint32) == g["T%g" % T][i].view(np.uint32)).all()


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    L = O.oracle()
    out = np.zeros(4, np.uint32)
    L.orc_philox4x32(0, 0, 0, 0, 0, 0, out)
    assert [hex(x) for x in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    L.orc_philox4x32(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, out)
    assert [hex(x) for x in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    L.orc_philox4x32(0xA4093822, 0x299F31D0, 0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, out)
    assert [hex(x) for x in out] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_vs_live_reference():
    L, R = O.oracle(), O.ref()
    n = 3000
    a = [np.zeros(n, np.uint64), np.zeros(n, np.int32), np.zeros(n, np.int32)]
    b = [np.zeros(n, np.uint64), np.zeros(n, np.int32), np.zeros(n, np.int32)]
    L.orc_playouts(123, 1 << 33, n, *a)
    R.ref_playouts(123, 1 << 33, n, *b)
    for x, y in zip(a, b):
        assert (x == y).all()
    for game in range(6):
        sts, _ = O.playout_states(321, game)
        for w in sts[1::5]:
            for sims, batch in ((50, 8), (64, 3)):
                for T in (1.0, 0.0):
                    sc, _, _ = O.oracle_mcts(w, T, sims, batch)
                    rc, _ = O.ref_mcts(w, T, sims, batch)
                    assert sc.shape == rc.shape and (sc.view(np.uint32) == rc.view(np.uint32)).all()


def test_replay_evaluator_reproduces_the_hash_search():
    """The table-fed search (what tests/test_gpu_replay.py uses to replay a GPU engine's network rows through the
    reference's UTTT::pv_mcts_scores, cpp/uttt_mcts.cpp:84-196): recording the leaves of a hash-evaluator search and
    replaying them gives the same float scores, in the C restatement and in the compiled reference."""
    detected = []
    for seed, game in ((77, 0), (77, 3)):
        sts = O.playout_states(seed, game)[0]
        for w in sts[::7]:
            for sims, batch, T in ((50, 8, 1.0), (50, 1, 1.0), (37, 5, 0.0), (200, 8, 1.0)):
                sc, st, pol, val = O.record_hash_mcts(w, T, sims, batch)
                want, _, _ = O.oracle_mcts(w, T, sims, batch)
                assert sc.tobytes() == want.tobytes()
                for use_ref in ([False, True] if O.ref_available() else [False]):
                    got, miss, unused = O.table_mcts(w, T, sims, batch, st, pol, val, use_ref=use_ref)
                    assert miss == 0 and unused == 0 and got.tobytes() == want.tobytes(), (sims, batch, T, use_ref)
                if len(val) > 3 and T == 1.0:        # rows given to the wrong leaves change the result
                    got, _, _ = O.table_mcts(w, T, sims, batch, st, np.roll(pol, 1, axis=0), np.roll(val, 1))
                    detected.append(got.tobytes() != want.tobytes())
    assert len(detected) > 20 and np.mean(detected) > 0.8, (len(detected), np.mean(detected))


def _py_scores(counts, T):
    """pv_mcts.py:166-180 on root visit counts (float64 python arithmetic)"""
    if T == 0:
        out = np.zeros(len(counts))
        out[int(np.argmax(counts))] = 1
        return out
    xs = [int(x) ** (1 / T) for x in counts]
    return np.array([x / sum(xs) for x in xs])


def test_python_search_oracle_matches_reference_goldens(golden_dir):
    """oracle restatement of the reference's pure-Python search (pv_mcts.py:74-180, the gating match's search) ==
    the scores of the unmodified reference module (tests/golden/pymcts.npz, minted by oracle/gen_golden.py)"""
    import os
    with np.load(os.path.join(golden_dir, "pymcts.npz")) as z:
        states, cases, scores = z["states"], z["cases"], z["scores"].view(np.float64)
    assert len(cases) > 2000
    differs_from_cpp = 0
    for (si, sims, batch, T, n), want in zip(cases, scores):
        si, sims, batch, n = int(si), int(sims), int(batch), int(n)
        cn = O.oracle_py_mcts(states[si], sims, batch)
        assert len(cn) == n and cn.sum() <= sims
        got = _py_scores(cn, T)
        assert got.tobytes() == want[:n].tobytes(), (si, sims, batch, T)
        if T == 1.0 and sims == 50 and batch == 8:
            differs_from_cpp += int((O.oracle_mcts(states[si], 1.0, sims, batch)[1] != cn).any())
    assert differs_from_cpp > 50          # the two searches of the reference really are different algorithms
