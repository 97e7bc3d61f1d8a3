"""Measured data points for the BASELINE.json configs that are not the bench line:
C2 (2^20 random playouts, rules only) and C4 (4096 games x 800 sims/move, throughput mode, Dirichlet noise)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ultimate-tictactoe-alphazero_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import engine  # noqa: E402
import oracle_lib as O  # noqa: E402
from dual_network import DualNetwork  # noqa: E402

out = {}
# ---- C2: rules only
n = 1 << 20
engine.game_playout(1, 0, n)                     # warm-up at full size (first launch + output allocation)
torch.cuda.synchronize()
ms = 1e9
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dg, pl, rs = engine.game_playout(0x5EED, 0, n); e1.record(); torch.cuda.synchronize()
    ms = min(ms, e0.elapsed_time(e1))
plies = int(pl.sum().item())
m = 20000
d2 = np.zeros(m, np.uint64); p2 = np.zeros(m, np.int32); r2 = np.zeros(m, np.int32)
t0 = time.perf_counter(); O.oracle().orc_playouts(0x5EED, 0, m, d2, p2, r2); dt = time.perf_counter() - t0
out["C2"] = {"playouts": n, "transitions": plies, "gpu_ms": ms, "gpu_transitions_per_s": plies / (ms / 1e3),
             "cpu_oracle_transitions_per_s_1core": int(p2.sum()) / dt,
             "bit_exact_vs_oracle_first_20000": bool((dg[:m].cpu().numpy().view(np.uint64) == d2).all())}
print(json.dumps(out["C2"]), flush=True)
# ---- C4: 4096 concurrent games x 800 sims, throughput mode, Dirichlet eps 0.25 alpha 0.3, 8 leaves / tree / round
games = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.manual_seed(0)
e = engine.Engine(n_slots=games, max_sims=800, max_batch=8, max_games=games)
e.upload_model(DualNetwork().eval())
e.set_root_noise(0.3, 0.25)
torch.cuda.synchronize(); t0 = time.perf_counter()
st = e.selfplay_device(games, sims=800, batch=8, seed=1, evaluator=engine.EVAL_NET_BF16, flags=engine.SP_THROUGHPUT)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
out["C4"] = {"games": games, "sims_per_move": 800, "leaves_per_tree_per_round": 8, "plies": int(st[0]), "sims": int(st[1]),
             "nn_evals": int(st[2]), "rounds": int(st[3]), "wall_s": dt, "sims_per_s": st[1] / dt, "evals_per_s": st[2] / dt,
             "moves_per_s": st[0] / dt, "trunk_tflops_equiv": st[2] * 764411904 / dt / 1e12}
print(json.dumps(out["C4"]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "config_points.json"), "w"), indent=1)
