"""One self-play cycle (for ncu / timing experiments): python tools/prof_selfplay.py --games 500"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ultimate-tictactoe-alphazero_b200"))
import torch  # noqa: E402
import engine  # noqa: E402
from dual_network import DualNetwork  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=500)
ap.add_argument("--slots", type=int, default=4096)
ap.add_argument("--sims", type=int, default=50)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--numerics", default="bf16")
a = ap.parse_args()
torch.manual_seed(0)
e = engine.Engine(n_slots=min(a.games, a.slots), max_sims=a.sims, max_batch=a.batch, max_games=a.games)
e.upload_model(DualNetwork().eval())
e.set_profile_level(int(os.environ.get("UTTT_PROFILE", "2")))
ev = engine.evaluator_of(a.numerics)
for r in range(a.reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = e.selfplay_device(a.games, sims=a.sims, batch=a.batch, seed=1, evaluator=ev, game0=r * a.games)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    prof = e.last_run_profile()
    t_ms, t_n, t_ev = prof["trunk_timed"]          # the trunk launches bracketed by events (all of them at profile level 2)
    flop = t_ev * 32 * 2 * 81 * 128 * 1152
    print("games=%d plies=%d evals=%d rounds=%d wall=%.3fs moves/s=%.0f | tree %.1f ms trunk %.1f ms (%.1f TFLOP/s, %.3f ms/launch, %d launches timed) heads %.1f ms"
          % (a.games, st[0], st[2], st[3], dt, st[0] / dt, prof["tree"][0], t_ms,
             flop / (t_ms / 1e3) / 1e12 if t_ms else 0, t_ms / max(t_n, 1), t_n, prof["heads"][0]))
    h = e.batch_histogram()
    print("  evaluator batch sizes (positions: launches): " +
          " ".join("%d-%d:%d" % (16 * i, 16 * i + 15, n) for i, n in enumerate(h) if n))
