"""Small end-to-end case for compute-sanitizer: rules kernels, both search modes, all three trunk variants, self-play."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ultimate-tictactoe-alphazero_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, engine, oracle_lib as O
from dual_network import DualNetwork
torch.manual_seed(0)
e = engine.Engine(n_slots=64, max_sims=50, max_batch=8, max_games=64)
e.upload_model(DualNetwork().eval())
engine.game_playout(1, 0, 4096)
sts = np.concatenate([O.playout_states(1, g)[0][:-1] for g in range(14)])
for n in (3, 40, 400, 560):                       # tc2<2> (P=1), tc2<2>, tc2<3>, tc
    d = torch.from_numpy(sts[:n].view(np.int32)).cuda()
    p, v = e.net_forward(d, engine.EVAL_NET_BF16)
p, v = e.net_forward(torch.from_numpy(sts[:8].view(np.int32)).cuda(), engine.EVAL_NET_FP32)
e.mcts_search(sts[:32], 50, 8, 1.0, engine.EVAL_HASH)
e.mcts_search(sts[:32], 50, 4, 1.0, engine.EVAL_HASH, flags=engine.SP_THROUGHPUT)
h = e.selfplay(16, sims=20, batch=4, seed=1, evaluator=engine.EVAL_NET_BF16)
h = e.selfplay(16, sims=20, batch=4, seed=1, evaluator=engine.EVAL_HASH, flags=engine.SP_THROUGHPUT)
torch.cuda.synchronize()
print("sanitize case done", int(h.stats[0]))
