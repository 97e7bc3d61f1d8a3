"""Phase timeline of one trunk CTA: python tools/trunk_timeline.py [n_positions] [bf16|bf16x3]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ultimate-tictactoe-alphazero_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, engine, oracle_lib as O
from dual_network import DualNetwork
n = int(sys.argv[1]) if len(sys.argv) > 1 else 345
MODE = engine.evaluator_of(sys.argv[2] if len(sys.argv) > 2 else "bf16")
torch.manual_seed(0)
e = engine.Engine(n_slots=max(n, 8), max_sims=50, max_batch=8, max_games=8)
e.upload_model(DualNetwork().eval())
sts = np.concatenate([O.playout_states(1, g)[0][:-1] for g in range(n // 40 + 2)])[:n]
d = torch.from_numpy(sts.view(np.int32)).cuda()
for _ in range(3):
    e.net_forward(d, MODE)
torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(); e.net_forward(d, MODE); t1.record(); torch.cuda.synchronize()
tl = e.trunk_timeline()
mma = tl[:, 1] - tl[:, 0]; wait_acc = tl[:, 2] - tl[:, 1]; epi = tl[:, 3] - tl[:, 2]
layer = np.diff(tl[:, 0])
print("n=%d forward %.3f ms" % (n, t0.elapsed_time(t1)))
print("per layer (cycles): MMA issue span mean %.0f | issue->accum ready %.0f | epilogue %.0f | layer period %.0f" % (
    mma.mean(), wait_acc.mean(), epi[:-1].mean(), layer.mean()))
print("even layers epi %.0f odd layers epi %.0f" % (epi[0:-1:2].mean(), epi[1:-1:2].mean()))
print("first 4 layers:", (tl[:4] - tl[0, 0]).tolist())
print("last 2 layers:", (tl[30:] - tl[30, 0]).tolist(), "last-layer epilogue", int(epi[-1]))
