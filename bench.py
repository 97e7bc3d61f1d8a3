#!/usr/bin/env python
"""bench.py -- self-play throughput of the B200 engine on BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                    (the reference's CPU path on the host cores)

One "step" = one self-play cycle of --games games per GPU (default 500 = SP_GAME_COUNT of the
reference, BASELINE.json configs[2]) at 50 simulations/move, MCTS_BATCH_SIZE 8, random-init
DualNetwork 128x16, reference-exact ("compat") search semantics.  Games are independent, so
ranks shard them with no collective on the search path (weak scaling: fixed games per GPU).

value : plies of all ranks / device time (CUDA events on the launching stream, max over ranks),
        weights and buffers resident in HBM
e2e   : the same cycle through the public host API -- Engine.upload_state_dict(host weights) +
        Engine.selfplay(...) -> host History -- with the H2D / D2H copies inside the timed region
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "ultimate-tictactoe-alphazero_b200")
for p in (PKG, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

TRUNK_FLOP_PER_POSITION = 32 * 2 * 81 * 128 * 1152          # 764,411,904 (dense 3x3 convs)
NET_FLOP_PER_POSITION = 765102212


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "sm_max_mhz": 1965.0}, \
        "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args, rank):
    """reference arm: the reference's own CPU implementation (compiled C++ MCTS + rules from
    oracle/_ref, DualNetwork fp32 through PyTorch on all host threads); each step is a bounded
    sample of the same workload (args.ref_moves plies of 50-simulation self-play)."""
    if rank != 0:
        return
    import torch
    import cpu_selfplay
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    from dual_network import DualNetwork
    import numpy as np
    try:
        ref = cpu_selfplay.load_reference_module()
        kind = "reference"
    except FileNotFoundError as e:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built: %s" % e}))
        return
    model = DualNetwork().eval()
    rng = np.random.RandomState(0)
    for _ in range(args.warmup):
        cpu_selfplay.play_moves(ref, model, 2, args.sims, args.batch, rng=rng)
    stats = {"forward_s": 0.0, "forwards": 0, "positions": 0}
    t0 = time.perf_counter()
    plies = 0
    for _ in range(args.steps):
        plies += cpu_selfplay.play_moves(ref, model, args.ref_moves, args.sims, args.batch, rng=rng, stats=stats)
    dt = time.perf_counter() - t0
    v = plies / dt
    sample = "%d plies/step of %d-sim batch-%d self-play (reference C++ MCTS + PyTorch fp32 forward on %d host threads)" % (
        args.ref_moves, args.sims, args.batch, cores)
    print(json.dumps({
        "impl": "reference", "metric": "self-play moves/sec (50 sims/move)", "value": v, "unit": "moves/s",
        "sims_per_s": v * args.sims, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (random-init weights, games from the initial position)",
        "config": workload_config(args), "forward_frac": stats["forward_s"] / dt,
        "cpu_baseline": {"value": v, "unit": "moves/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(args):
    return {"workload": "C3 (BASELINE.json configs[2]): %d-game self-play cycle per GPU, %d sims/move, "
                        "MCTS_BATCH_SIZE=%d, DualNetwork 128fx16 random-init" % (args.games, args.sims, args.batch),
            "games_per_gpu": args.games, "sims_per_move": args.sims, "mcts_batch_size": args.batch,
            "search": "compat (reference-exact queue/flush semantics)", "numerics": args.numerics,
            "l2": "flushed between timed steps (256 MiB device write)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--games", type=int, default=500, help="self-play games per GPU per step (SP_GAME_COUNT)")
    ap.add_argument("--sims", type=int, default=50)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--numerics", default="bf16", choices=["bf16", "bf16x3", "fp32"])
    ap.add_argument("--ref-moves", type=int, default=24, help="plies per step of the reference arm")
    ap.add_argument("--cpu-baseline-moves", type=int, default=96)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--saturated-games", type=int, default=4096,
                    help="also time one cycle of this many concurrent games per GPU (0 = skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import engine
    from dual_network import DualNetwork

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev_kind = {"fp32": engine.EVAL_NET_FP32, "bf16": engine.EVAL_NET_BF16, "bf16x3": engine.EVAL_NET_BF16X3}[args.numerics]
    eng = engine.Engine(n_slots=min(args.games, 4096), max_sims=args.sims, max_batch=args.batch,
                        max_games=args.games, device=local_rank)
    torch.manual_seed(0)
    model = DualNetwork().eval()
    sd_host = {k: v.pin_memory() for k, v in model.state_dict().items()}
    w_bytes = sum(v.numel() * v.element_size() for v in sd_host.values())
    eng.upload_state_dict(sd_host)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step(i):
        return eng.selfplay_device(args.games, sims=args.sims, batch=args.batch, seed=0x5EED, evaluator=ev_kind,
                                   game0=(i * world + rank) * args.games, stream=stream)

    for i in range(args.warmup):
        step(1000 + i)
    # ---------------- device-resident timed region
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    dev_ms, plies, sims, evals, rounds = 0.0, 0, 0, 0, 0
    trunk_ms, trunk_launches, all_launches = 0.0, 0, 0
    for i in range(args.steps):
        flush.fill_(i & 0xFF)                      # evict L2 between timed steps (not timed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        st = step(i)
        e1.record(stream)
        e1.synchronize()
        dev_ms += e0.elapsed_time(e1)
        plies += int(st[0]); sims += int(st[1]); evals += int(st[2]); rounds += int(st[3])
        prof = eng.last_run_profile()
        trunk_ms += prof["trunk"][0]; trunk_launches += prof["trunk"][1]; all_launches += prof["all"][1]
    barrier()
    wall_s = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    # diagnostic, outside the timed region: one more cycle with every kernel bracketed by events (level 2 costs ~1.5 %
    # of a cycle in GPU idle time at the extra event boundaries, which is why the timed steps only bracket the trunk)
    eng.set_profile_level(2)
    st = step(2000)
    torch.cuda.synchronize()
    prof = eng.last_run_profile()
    split = {"tree_kernels": prof["tree"][0], "trunk": prof["trunk"][0], "heads": prof["heads"][0],
             "device_total": prof["all"][0], "rounds": int(st[3])}
    eng.set_profile_level(1)

    # ---------------- end-to-end through the public host API (H2D weights, D2H history inside the timed region)
    hist = engine.History(args.games)
    eng.upload_state_dict(sd_host); eng.selfplay(args.games, sims=args.sims, batch=args.batch, seed=1, evaluator=ev_kind,
                                                 game0=10 ** 6, history=hist)       # warm-up
    barrier()
    t0 = time.perf_counter()
    e2e_plies = 0
    for i in range(args.steps):
        eng.upload_state_dict(sd_host)
        h = eng.selfplay(args.games, sims=args.sims, batch=args.batch, seed=0x5EED, evaluator=ev_kind,
                         game0=(i * world + rank) * args.games, history=hist)
        e2e_plies += int(h.stats[0])
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---------------- secondary: the machine-filling workload (C5's per-GPU share: 4096 concurrent games)
    sat = None
    if args.saturated_games > 0:
        eng.close()
        eng = engine.Engine(n_slots=min(args.saturated_games, 4096), max_sims=args.sims, max_batch=args.batch,
                            max_games=args.saturated_games, device=local_rank)
        eng.upload_state_dict(sd_host)
        eng.selfplay_device(args.saturated_games, sims=args.sims, batch=args.batch, seed=7, evaluator=ev_kind,
                            game0=5 * 10 ** 6, stream=stream)                       # warm-up
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        st = eng.selfplay_device(args.saturated_games, sims=args.sims, batch=args.batch, seed=8, evaluator=ev_kind,
                                 game0=6 * 10 ** 6 + rank * args.saturated_games, stream=stream)
        e1.record(stream)
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        sat = {"games_per_gpu": args.saturated_games, "moves_per_s_rank0": int(st[0]) / (ms / 1e3),
               "nn_evals_per_s_rank0": int(st[2]) / (ms / 1e3),
               "trunk_tflops_equiv_rank0": int(st[2]) * TRUNK_FLOP_PER_POSITION / (ms / 1e3) / 1e12,
               "note": "whole-pipeline rate (tree + trunk + heads, two overlapped lanes); trunk_tflops_equiv = evals x "
                       "764.4 MFLOP / wall, i.e. a lower bound on the trunk kernels' own rate"}

    red = torch.tensor([dev_ms, e2e_s, wall_s], dtype=torch.float64, device=dev)
    tot = torch.tensor([plies, sims, evals, e2e_plies, all_launches], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_s_max, wall_max = [float(x) for x in red.tolist()]
    plies_all, sims_all, evals_all, e2e_plies_all, launches_all = [float(x) for x in tot.tolist()]

    if rank == 0:
        peaks, peak_src = measured_peaks()
        value = plies_all / (dev_ms_max / 1e3)
        peak_tf = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
        # dominant kernel = the residual trunk (tcgen05 implicit GEMM); rank 0's launches
        achieved_tf = (evals * TRUNK_FLOP_PER_POSITION / (trunk_ms / 1e3)) / 1e12 if trunk_ms > 0 else 0.0
        out = {
            "metric": "self-play moves/sec (50 sims/move)", "value": value, "unit": "moves/s",
            "sims_per_s": sims_all / (dev_ms_max / 1e3), "nn_evals_per_s": evals_all / (dev_ms_max / 1e3),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.numerics == "bf16" else "f32",
            "data": "synthetic (random-init DualNetwork weights, seed 0; games from the initial position)",
            "config": workload_config(args),
            "clocks": clocks,
            "e2e": {"value": e2e_plies_all / e2e_s_max, "unit": "moves/s", "h2d_bytes_per_step": int(w_bytes),
                    "d2h_bytes_per_step": int(hist.nbytes)},
            "gpu_launches": int(launches_all),
            "roofline": {"bound": "tensor",
                         "kernel": "trunk_auto_kernel (one launch per round; on the device: trunk_tc2_body<2> up to 370 positions, "
                                   "trunk_pp_body<1> with cta_group::2 MMAs above)"
                         if args.numerics == "bf16" else "conv3x3_fp32_kernel",
                         "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf if peak_tf else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of trunk_auto_kernel at a
                         # batch of 500 positions (profiles/r1_trunkpp_full.md): weights 9.4 MB (+ the per-CTA-half copy),
                         # planes, head features, policy / value rows; activations and the skip connection never reach DRAM
                         "traffic": 15.18e6 if args.numerics == "bf16" else None,
                         "traffic_unit": "bytes per launch (tensor-bound kernel: informational)",
                         "peak_source": peak_src + ", sustained bf16 (kernel timed inside a long step)",
                         "flop_per_launch": evals * TRUNK_FLOP_PER_POSITION / max(trunk_launches, 1),
                         "avg_launch_ms": trunk_ms / max(trunk_launches, 1), "launches": trunk_launches},
            "breakdown_ms_rank0": {"trunk": trunk_ms, "tree_heads_and_gaps": dev_ms - trunk_ms, "device_total": dev_ms,
                                   "rounds": rounds, "evals": evals, "plies": plies,
                                   "one_extra_cycle_with_all_kernels_timed": split},
            "wall_s": wall_max,
            "saturated": sat,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                import cpu_selfplay
                r = cpu_selfplay.time_cpu_selfplay(args.cpu_baseline_moves, args.sims, args.batch)
                out["cpu_baseline"] = {
                    "value": r["moves_per_s"], "unit": "moves/s", "cores": r["cores"], "kind": "reference",
                    "sample": "%d plies of %d-sim batch-%d self-play: reference C++ MCTS (oracle/_ref) + PyTorch fp32 "
                              "DualNetwork on %d host threads, %.1f s, %.0f%% in forward" % (
                                  r["moves"], args.sims, args.batch, r["cores"], r["seconds"], 100 * r["forward_frac"])}
            except Exception as e:  # noqa: BLE001
                out["cpu_baseline"] = {"value": None, "unit": "moves/s", "cores": os.cpu_count(), "kind": "reference",
                                       "sample": "unavailable: %s" % e}
        print(json.dumps(out))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
