"""Phase timeline of the two-groups-in-flight trunk (layers 0..15 of groups A and B): python tools/pp_timeline.py [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ultimate-tictactoe-alphazero_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, engine, oracle_lib as O
from dual_network import DualNetwork
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
torch.manual_seed(0)
e = engine.Engine(n_slots=max(n, 8), max_sims=50, max_batch=8, max_games=8)
e.upload_model(DualNetwork().eval())
sts = np.concatenate([O.playout_states(1, g)[0][:-1] for g in range(n // 40 + 2)])[:n]
d = torch.from_numpy(sts.view(np.int32)).cuda()
for _ in range(3):
    e.net_forward(d, engine.EVAL_NET_BF16)
torch.cuda.synchronize()
tl = e.trunk_timeline()
A, B = tl[:16], tl[16:]
base = A[4, 0]
print("n=%d; per group and layer: [MMA start, MMA issued, accumulators ready, epilogue done] (cycles rel. MMA start of A, layer 4)" % n)
for l in range(4, 8):
    print(" layer %d  A %s   B %s" % (l, (A[l] - base).tolist(), (B[l] - base).tolist()))
print("period %.0f" % np.diff(A[2:15, 0]).mean())
