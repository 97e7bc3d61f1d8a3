"""temporary: chunk-level timeline of tc2<2> (needs UTTT_TC2_DETAIL build)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ultimate-tictactoe-alphazero_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, engine, oracle_lib as O
from dual_network import DualNetwork
n = int(sys.argv[1]) if len(sys.argv) > 1 else 345
torch.manual_seed(0)
e = engine.Engine(n_slots=max(n, 8), max_sims=50, max_batch=8, max_games=8)
e.upload_model(DualNetwork().eval())
sts = np.concatenate([O.playout_states(1, g)[0][:-1] for g in range(n // 40 + 2)])[:n]
d = torch.from_numpy(sts.view(np.int32)).cuda()
for _ in range(3):
    e.net_forward(d, engine.EVAL_NET_BF16)
torch.cuda.synchronize()
tl = e.trunk_timeline().reshape(-1)
base = tl[4 * 4 + 0]   # MMA start of layer 4
print("layers 3..6 [mma start, issued, accum ready, epi done] rel:", (tl[12:28].reshape(4, 4) - base).tolist())
print("epilogue(4) accum-wait done per warp:", (tl[96:112] - base).tolist())
print("epilogue(4) chunk 0 published per warp:", (tl[64:80] - base).tolist())
print("epilogue(4) chunk 3 published per warp:", (tl[80:96] - base).tolist())
for lt in range(2):
    I = tl[112 + 8 * lt:120 + 8 * lt] - base
    print("issuer tile %d (layer 5) [before wait, after wait] per quarter:" % lt, I.reshape(4, 2).tolist())
