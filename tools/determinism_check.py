"""Is a self-play cycle with the network evaluator reproducible run to run?  python tools/determinism_check.py [games]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ultimate-tictactoe-alphazero_b200"))
import numpy as np, torch, engine
from dual_network import DualNetwork
games = int(sys.argv[1]) if len(sys.argv) > 1 else 500
torch.manual_seed(0)
e = engine.Engine(n_slots=min(games, 4096), max_sims=50, max_batch=8, max_games=games)
e.upload_model(DualNetwork().eval())
runs = []
for r in range(3):
    h = e.selfplay(games, sims=50, batch=8, seed=77, evaluator=engine.EVAL_NET_BF16)
    runs.append((h.lens.copy(), h.actions.copy(), h.counts.copy()))
for r in (1, 2):
    same_games = sum(int(runs[0][0][g] == runs[r][0][g] and (runs[0][1][g] == runs[r][1][g]).all() and
                         (runs[0][2][g] == runs[r][2][g]).all()) for g in range(games))
    print("games=%d run %d vs run 0: %d / %d games identical (lengths, moves and visit counts)" % (games, r, same_games, games))
