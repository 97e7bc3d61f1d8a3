"""GPU: the packed-sample history (csrc/history_kernels.cu) and the multi-GPU cycle plumbing (parallel.SelfPlayCycle) on one GPU."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_pack_and_unpack_samples_on_device():
    """uttt_selfplay_pack == the host restatement (parallel.pack_history_host) byte for byte; uttt_samples_unpack gives the
    trainer's arrays: x = to_input_tensor planes in NCHW (train_network.py:47-49), policy = counts / sum, value = z"""
    import torch
    import engine
    import parallel
    e = engine.Engine(n_slots=96, max_sims=50, max_batch=8, max_games=200)
    try:
        e.selfplay_device(200, sims=50, batch=8, seed=4, evaluator=engine.EVAL_HASH)       # slots are recycled: 200 games on 96
        buf, n = e.selfplay_pack(200)
        h = e.selfplay_fetch(200)
        assert n == int(h.lens.sum()) and buf.numel() == n * engine.SAMPLE_BYTES
        want = parallel.pack_history_host(h, 200)
        assert torch.equal(buf.cpu(), want)
        st, cn, z, ply = engine.samples_to_numpy(buf.cpu().numpy())
        st2, cn2, z2 = h.samples()
        assert (st == st2).all() and (cn == cn2).all() and (z == z2).all() and ply.max() == h.lens.max() - 1
        x, p, v = engine.samples_unpack(buf)
        planes = engine.game_encode(torch.from_numpy(st.view(np.int32)).cuda()).permute(0, 3, 1, 2).contiguous()
        assert torch.equal(x, planes)
        pw = torch.from_numpy(cn.astype(np.float32) / cn.sum(1, keepdims=True).astype(np.float32)).cuda()
        assert torch.equal(p, pw) and torch.equal(v[:, 0].cpu(), torch.from_numpy(z.astype(np.float32)))
        # a second, shorter run re-uses the buffers: only the new games are packed
        e.selfplay_device(7, sims=50, batch=8, seed=5, evaluator=engine.EVAL_HASH)
        buf2, n2 = e.selfplay_pack(7, out=buf)
        assert n2 == int(e.selfplay_fetch(7).lens.sum()) and buf2.data_ptr() == buf.data_ptr()
        with pytest.raises(RuntimeError):
            e.selfplay_pack(7, out=torch.empty(engine.SAMPLE_BYTES, dtype=torch.uint8, device="cuda"))
    finally:
        e.close()


def test_selfplay_cycle_object_single_rank():
    """parallel.SelfPlayCycle with a world of one NCCL rank: engine and buffers persist across cycles, the result feeds the
    trainer drop-in directly"""
    import torch
    import torch.distributed as dist
    import engine
    import parallel
    import train_network as tn
    from dual_network import DualNetwork
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=0, world_size=1,
                            device_id=torch.device("cuda", 0))
    try:
        torch.manual_seed(0)
        model = DualNetwork().eval()
        cyc = parallel.SelfPlayCycle(48, sims=20, batch=8, numerics="bf16x3")
        try:
            h0 = cyc.engine.h
            r1 = cyc.run(model, seed=1, cycle=0)
            r2 = cyc.run(model, seed=1, cycle=0)
            r3 = cyc.run(model, seed=2, cycle=1)
            assert cyc.engine.h is h0 and r1["samples"].data_ptr() == r3["samples"].data_ptr()
            assert r1["n_samples"] == r2["n_samples"] >= 48 * 17 and r1["samples_per_rank"] == [r1["n_samples"]]
            assert torch.equal(r2["x"], r1["x"]) and torch.equal(r2["policy"], r1["policy"])      # same seed, same cycle
            assert r3["x"].shape == (r3["n_samples"], 3, 9, 9) and torch.allclose(r3["policy"].sum(1), torch.ones(r3["n_samples"], device="cuda"))
            assert set(cyc.timings) >= {"broadcast_ms", "selfplay_ms", "pack_ms", "gather_ms", "unpack_ms", "gather_bytes"}
            assert cyc.timings["gather_bytes"] == 0
            losses = tn.train_tensors(DualNetwork().cuda(), r3["x"], r3["policy"], r3["value"], epochs=2, batch_size=64,
                                      log=lambda s: None)
            assert len(losses) == 2 and np.isfinite(losses).all()
        finally:
            cyc.close()
    finally:
        dist.destroy_process_group()
