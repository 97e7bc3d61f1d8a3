"""One sharded self-play cycle over all ranks (launch with torchrun): NCCL weight broadcast, per-rank self-play,
NCCL history gather, rank 0 converts to the reference's .history format.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/cycle_multi_gpu.py --games 1000"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ultimate-tictactoe-alphazero_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import engine  # noqa: E402
import parallel  # noqa: E402
from dual_network import DualNetwork  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=1000)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(rank)                       # ranks start with DIFFERENT weights; rank 0's must win
model = DualNetwork().eval()
ref_sum = float(sum(v.double().sum() for v in model.state_dict().values()))
t0 = time.perf_counter()
gathered, stats = parallel.sharded_self_play(model, a.games, sims=50, batch=8, seed=123)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
sums = [None] * world
dist.all_gather_object(sums, float(sum(v.double().sum() for v in model.state_dict().values())))
if rank == 0:
    assert all(abs(x - sums[0]) < 1e-9 for x in sums), "weights differ after broadcast"
    assert gathered["lens"].shape == (a.games,) and (gathered["lens"] >= 17).all()
    n_samples = int(gathered["lens"].sum())
    cnt = gathered["counts"]
    assert all((cnt[g, :gathered["lens"][g]].sum(1) == 50).all() for g in range(0, a.games, 37))
    print("ranks=%d games=%d samples=%d wall=%.2fs (incl. engine creation) -> %.0f moves/s; weights identical on all ranks; "
          "history gathered on rank 0: %.1f MB packed" % (world, a.games, n_samples, dt, n_samples / dt,
                                                            sum(v.nbytes for v in gathered.values()) / 1e6))
dist.destroy_process_group()
