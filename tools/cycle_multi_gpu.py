"""Sharded self-play cycles over all ranks (launch with torchrun): NCCL weight broadcast, per-rank self-play, one packed
exact-length history transfer per rank to rank 0, samples expanded into the trainer's tensors there.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/cycle_multi_gpu.py --games 8192"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ultimate-tictactoe-alphazero_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import parallel  # noqa: E402
from dual_network import DualNetwork  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=8192)
ap.add_argument("--cycles", type=int, default=3)
ap.add_argument("--numerics", default="bf16")
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(rank)                       # ranks start with DIFFERENT weights; rank 0's must win
model = DualNetwork().eval()
cyc = parallel.SelfPlayCycle(a.games, numerics=a.numerics, device=local)
for c in range(a.cycles):
    res = cyc.run(model, seed=123 + c, cycle=c)
    t = torch.tensor([cyc.timings[k] for k in ("total_ms", "broadcast_ms", "selfplay_ms", "pack_ms", "gather_ms", "unpack_ms")],
                     dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        n = res["n_samples"]
        assert res["x"].shape == (n, 3, 9, 9) and sum(res["samples_per_rank"]) == n
        tot, bc, sp, pk, ga, up = t.tolist()
        print(json.dumps({"cycle": c, "ranks": world, "games": a.games, "samples": n, "moves_per_s_e2e": n / (tot / 1e3),
                          "total_ms": tot, "broadcast_ms": bc, "selfplay_ms": sp, "pack_ms": pk, "gather_ms": ga, "unpack_ms": up,
                          "gather_bytes": cyc.timings["gather_bytes"]}))
sums = [None] * world
dist.all_gather_object(sums, float(sum(v.double().sum() for v in model.state_dict().values())))
if rank == 0:
    assert all(abs(x - sums[0]) < 1e-9 for x in sums), "weights differ after broadcast"
cyc.close()
dist.destroy_process_group()
