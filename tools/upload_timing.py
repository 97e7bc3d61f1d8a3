"""time Engine.upload_state_dict (pinned host state_dict) and the history fetch: python tools/upload_timing.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ultimate-tictactoe-alphazero_b200"))
import numpy as np, torch, engine
from dual_network import DualNetwork
torch.manual_seed(0)
e = engine.Engine(n_slots=500, max_sims=50, max_batch=8, max_games=500)
sd = {k: v.pin_memory() for k, v in DualNetwork().state_dict().items()}
e.upload_state_dict(sd)
for name, fn in (("upload_state_dict (pinned)", lambda: e.upload_state_dict(sd)),
                 ("  of which pack_small + pointer table (python)", lambda: (engine.pack_small(sd), engine.scattered_residual_tensors(sd)))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        fn()
    torch.cuda.synchronize(); print("%s: %.3f ms" % (name, (time.perf_counter() - t0) / 20 * 1e3))
e.selfplay_device(500, sims=50, batch=8, seed=1, evaluator=engine.EVAL_NET_BF16)
hist = engine.History(500)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    e.selfplay_fetch(500, history=hist)
print("selfplay_fetch (7.9 MB, pinned): %.3f ms" % ((time.perf_counter() - t0) / 20 * 1e3))
