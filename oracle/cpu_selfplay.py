"""cpu_selfplay.py -- TEST / BASELINE INFRASTRUCTURE ONLY (never imported by the product).

The reference's CPU self-play path, timed as the CPU baseline by bench.py:
  * search + rules: the UNMODIFIED reference C++ (`uttt_cpp` pybind11 module compiled from the
    reference sources into oracle/_ref/ by oracle/Makefile), cpp/uttt_mcts.cpp:84-196
  * the Python glue is a restatement of pv_mcts_cpp.py:17-89 (inference closure) and
    self_play_cpp.py:34-101 (play loop) -- the reference's .py files cannot travel to the GPU box
  * the network: DualNetwork fp32 on the host cores through PyTorch, like the reference on a
    machine without CUDA (dual_network.py:18).
"""
import importlib.util
import os
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference_module():
    """import oracle/_ref/uttt_cpp*.so WITHOUT registering it as `uttt_cpp` (that name is the product shim)"""
    import sysconfig
    path = os.path.join(HERE, "_ref", "uttt_cpp" + sysconfig.get_config_var("EXT_SUFFIX"))
    if not os.path.exists(path):
        raise FileNotFoundError(path + " (build with `make -C oracle ref` where /root/reference exists)")
    spec = importlib.util.spec_from_file_location("uttt_cpp", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_inference(model, stats):
    """pv_mcts_cpp.py:37-78"""
    def inference_func(states_list):
        t0 = time.perf_counter()
        x = np.stack([np.array(s.to_input_tensor(), dtype=np.float32).reshape(9, 9, 3) for s in states_list], axis=0)
        x = torch.from_numpy(np.ascontiguousarray(np.transpose(x, (0, 3, 1, 2))))
        with torch.no_grad():
            policies, values = model(x)
        policies = policies.numpy()
        values = values.numpy()
        stats["forward_s"] += time.perf_counter() - t0
        stats["forwards"] += 1
        stats["positions"] += len(states_list)
        return [(policies[i], float(values[i][0])) for i in range(len(states_list))]
    return inference_func


def play_moves(ref, model, max_moves, sims=50, batch=8, temperature=1.0, rng=None, stats=None):
    """self_play_cpp.py:34-101, stopping after max_moves plies (a bounded sample of the workload);
    starts a new game whenever one finishes.  Returns the number of plies played."""
    rng = rng or np.random.RandomState(0)
    stats = stats if stats is not None else {"forward_s": 0.0, "forwards": 0, "positions": 0}
    infer = make_inference(model, stats)
    model.eval()
    plies = 0
    state = ref.State()
    while plies < max_moves:
        if state.is_done():
            state = ref.State()
        scores = np.array(ref.pv_mcts_scores(infer, state, temperature, sims, batch), dtype=np.float64)
        legal = state.legal_actions()
        scores = np.ones(len(scores)) / len(scores) if scores.sum() == 0 else scores / scores.sum()
        state = state.next(int(rng.choice(legal, p=scores)))
        plies += 1
    return plies


def time_cpu_selfplay(n_moves, sims=50, batch=8, threads=None, seed=0):
    """-> dict(moves_per_s, sims_per_s, seconds, cores, forward_frac, positions)"""
    from dual_network import DualNetwork
    ref = load_reference_module()
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    torch.manual_seed(seed)
    model = DualNetwork().eval()
    stats = {"forward_s": 0.0, "forwards": 0, "positions": 0}
    play_moves(ref, model, 2, sims, batch, stats={"forward_s": 0.0, "forwards": 0, "positions": 0})   # warm-up
    t0 = time.perf_counter()
    plies = play_moves(ref, model, n_moves, sims, batch, rng=np.random.RandomState(seed), stats=stats)
    dt = time.perf_counter() - t0
    return {"moves_per_s": plies / dt, "sims_per_s": plies * sims / dt, "seconds": dt, "cores": threads,
            "forward_frac": stats["forward_s"] / dt, "positions": stats["positions"], "moves": plies}
