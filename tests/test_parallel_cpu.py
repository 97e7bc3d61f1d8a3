"""CPU, world_size 2, gloo: the multi-process host logic (sharding plan, weight broadcast, history gather)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_games, tmp):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "ultimate-tictactoe-alphazero_b200"))
    import engine
    import parallel
    from dual_network import DualNetwork
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                      # different weights on every rank ...
        model = DualNetwork()
        parallel.broadcast_state_dict(model.state_dict(), src=0)
        torch.manual_seed(100)
        ref = DualNetwork()
        for k, v in model.state_dict().items():            # ... identical to rank 0's afterwards
            assert torch.equal(v, ref.state_dict()[k]), k
        game0, count = parallel.shard_games(n_games, world, rank)
        h = engine.History(max(count, 1), pinned=False)
        for i in range(count):                              # synthetic history keyed by the global game id
            g = game0 + i
            h.lens[i] = 20 + g % 7
            h.final[i] = g % 2
            h.actions[i, :] = g % 81
            h.states[i, :, 0] = g
            h.counts[i, :, :] = g % 50
        out = parallel.gather_histories(h, count, dst=0)
        if rank == 0:
            lens = 20 + np.arange(n_games) % 7
            gid = np.repeat(np.arange(n_games), lens)                 # global game id of every sample, game-major
            ply = np.concatenate([np.arange(n) for n in lens])
            assert out["states"].shape == (lens.sum(), 8) and out["states"].dtype == np.uint32
            assert (out["states"][:, 0] == gid).all() and (out["ply"] == ply).all()
            assert out["counts"].dtype == np.uint16 and (out["counts"] == (gid % 50)[:, None]).all()
            z0 = np.where(gid % 2 != 0, -1, 0)                        # self_play_cpp.py:95-99
            assert (out["z"] == z0 * np.where(ply % 2 == 0, 1, -1)).all()
            sizes = [parallel.shard_games(n_games, world, r) for r in range(world)]
            assert out["samples_per_rank"].tolist() == [int(lens[g0:g0 + c].sum()) for g0, c in sizes]
            open(os.path.join(tmp, "ok"), "w").write("ok")
        else:
            assert out is None
        # exact-length transfer incl. an empty rank and a caller-provided destination buffer
        mine = torch.full(((rank * 3) * engine.SAMPLE_BYTES,), rank + 1, dtype=torch.uint8)
        dest = torch.zeros(10 * engine.SAMPLE_BYTES, dtype=torch.uint8) if rank == 1 else None
        buf, counts = parallel.gather_samples(mine, dst=1, out=dest)
        assert counts == [0, 3]
        if rank == 1:
            assert buf.data_ptr() == dest.data_ptr() and buf.numel() == 3 * engine.SAMPLE_BYTES and (buf == 2).all()
        else:
            assert buf is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_games", [7, 64])
def test_two_rank_broadcast_and_gather(tmp_path, n_games):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_games, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


def test_shard_plan_covers_every_game_once():
    import parallel
    for n in (0, 1, 5, 500, 32768):
        for w in (1, 2, 4, 8):
            seen = []
            for r in range(w):
                g0, c = parallel.shard_games(n, w, r)
                seen.extend(range(g0, g0 + c))
            assert seen == list(range(n))
            sizes = [parallel.shard_games(n, w, r)[1] for r in range(w)]
            assert max(sizes) - min(sizes) <= 1
