"""GPU: rules kernels (through the C ABI) vs the reference goldens and the CPU oracle, bit-exact."""
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rules(golden_dir):
    with np.load(os.path.join(golden_dir, "rules.npz")) as z:
        return {k: z[k] for k in z.files}


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_playout_digests_golden(rules):
    import engine
    n = len(rules["digests"])
    dg, pl, rs = engine.game_playout(int(rules["seed"]), 0, n)
    assert (dg.cpu().numpy().view(np.uint64) == rules["digests"]).all()
    assert (pl.cpu().numpy() == rules["plies"]).all()
    assert (rs.cpu().numpy() == rules["results"]).all()


def test_playout_full_size_checksums(rules):
    """config 2 of BASELINE.json: 2^20 concurrent playouts; checksum-of-digests vs the reference run"""
    import engine
    n = int(rules["full_n"])
    dg, pl, rs = engine.game_playout(int(rules["seed"]), 0, n)
    d = dg.cpu().numpy().view(np.uint64)
    assert np.bitwise_xor.reduce(d) == rules["full_xor"]
    assert np.uint64(int(d.astype(object).sum()) & 0xFFFFFFFFFFFFFFFF) == rules["full_sum"]
    assert int(pl.sum().item()) == int(rules["full_plies"])
    assert (np.bincount(rs.cpu().numpy(), minlength=3) == rules["full_hist"]).all()


def test_playout_other_seeds_vs_oracle():
    import engine
    L = O.oracle()
    for seed, g0, n in ((1, 0, 2000), (0xDEADBEEF, (1 << 40) + 17, 1500)):
        dg, pl, rs = engine.game_playout(seed, g0, n)
        d2 = np.zeros(n, np.uint64); p2 = np.zeros(n, np.int32); r2 = np.zeros(n, np.int32)
        L.orc_playouts(seed, g0, n, d2, p2, r2)
        assert (dg.cpu().numpy().view(np.uint64) == d2).all()
        assert (pl.cpu().numpy() == p2).all() and (rs.cpu().numpy() == r2).all()


def test_legal_status_encode_step_golden(rules):
    import engine
    st = _dev(rules["states"].view(np.int32))
    masks, status = engine.game_legal_mask(st)
    masks = masks.cpu().numpy().view(np.uint32)
    status = status.cpu().numpy()
    n = len(rules["states"])
    for i in range(n):
        legal = [a for a in range(81) if (masks[i, a // 27] >> (a % 27)) & 1]
        k = rules["n_legal"][i]
        assert legal == rules["legal"][i, :k].tolist()
        assert masks[i, 3] == k
        f = rules["flags"][i]
        assert status[i] == (1 if f & 1 else (2 if f & 2 else 0))
    planes = engine.game_encode(st).cpu().numpy().reshape(n, 243)
    assert (planes == rules["tensor"].astype(np.float32)).all()
    chw = engine.game_gather_planes(st).float().cpu().numpy()             # (n,3,9,9)
    assert (chw.transpose(0, 2, 3, 1).reshape(n, 243) == rules["tensor"].astype(np.float32)).all()
    # next() for every legal action of every golden state
    idx, acts, ref = [], [], []
    for i in range(n):
        for k in range(rules["n_legal"][i]):
            idx.append(i); acts.append(rules["legal"][i, k]); ref.append(rules["next"][i, k])
    import torch
    out = engine.game_step(st[torch.tensor(idx).cuda()].contiguous(), _dev(np.array(acts, np.int32)))
    assert (out.cpu().numpy().view(np.uint32) == np.stack(ref)).all()
    out = engine.game_step(_dev(rules["ill_states"].view(np.int32)), _dev(rules["ill_actions"].astype(np.int32)))
    assert (out.cpu().numpy().view(np.uint32) == rules["ill_next"]).all()


def test_rules_kernels_ragged_and_empty():
    import torch
    import engine
    empty = torch.empty((0, 8), dtype=torch.int32, device="cuda")
    assert engine.game_encode(empty).shape == (0, 9, 9, 3)
    assert engine.game_legal_mask(empty)[0].shape == (0, 4)
    assert engine.game_step(empty, torch.empty((0,), dtype=torch.int32, device="cuda")).shape == (0, 8)
    # odd sizes around the block size, checked against the oracle
    sts = np.concatenate([O.playout_states(77, g)[0] for g in range(12)])
    for n in (1, 31, 255, 257, len(sts)):
        m, s = engine.game_legal_mask(_dev(sts[:n].view(np.int32)))
        m = m.cpu().numpy().view(np.uint32)
        for i in (0, n // 2, n - 1):
            flags, legal, _ = O.oracle_probe(sts[i])
            assert [a for a in range(81) if (m[i, a // 27] >> (a % 27)) & 1] == legal.tolist()


def test_encode_gather_ragged_and_unaligned(rules):
    """plane writers: sizes around the 256-position block, and destinations off the 16-byte grid (other kernel)"""
    import torch
    import engine
    reps = 3
    sts = np.concatenate([rules["states"]] * reps)
    ref = np.concatenate([rules["tensor"]] * reps).astype(np.float32)             # (n,243) HWC
    ref_chw = ref.reshape(-1, 9, 9, 3).transpose(0, 3, 1, 2).reshape(-1, 243)
    st = _dev(sts.view(np.int32))
    for n in (1, 3, 255, 256, 257, 511, 1000, len(sts)):
        assert (engine.game_encode(st[:n]).cpu().numpy().reshape(n, 243) == ref[:n]).all()
        assert (engine.game_gather_planes(st[:n]).float().cpu().numpy().reshape(n, 243) == ref_chw[:n]).all()
        for off in (1, 2, 3):
            buf = torch.full((n * 243 + 8,), -7.0, dtype=torch.float32, device="cuda")
            engine.game_encode(st[:n], out=buf[off:off + n * 243])
            b = buf.cpu().numpy()
            assert (b[off:off + n * 243].reshape(n, 243) == ref[:n]).all()
            assert (b[:off] == -7.0).all() and (b[off + n * 243:] == -7.0).all()
            buf = torch.full((n * 243 + 8,), -7.0, dtype=torch.bfloat16, device="cuda")
            engine.game_gather_planes(st[:n], out=buf[off:off + n * 243])
            b = buf.float().cpu().numpy()
            assert (b[off:off + n * 243].reshape(n, 243) == ref_chw[:n]).all()
            assert (b[:off] == -7.0).all() and (b[off + n * 243:] == -7.0).all()
    # aligned destination: nothing written past the end either
    n = 257
    buf = torch.full((n * 243 + 16,), -7.0, dtype=torch.float32, device="cuda")
    engine.game_encode(st[:n], out=buf[:n * 243])
    assert (buf[n * 243:].cpu().numpy() == -7.0).all()
