"""Multi-GPU plumbing around the self-play path (one process per GPU, torch.distributed).

Games are independent, so the search path has NO collective: each rank plays a contiguous block of
game ids on its own GPU.  Collectives are used off the path only, once per cycle:
  * broadcast_state_dict: the new ./model/best.pth weights, one flattened buffer from `src`
  * gather_histories:     the packed self-play histories to `dst`, which writes the .history file
Backend "nccl" (GPU tensors over NVLink) in production, "gloo" (CPU tensors) in the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_games(n_games, world, rank):
    """contiguous block of game ids for `rank`: (game0, count); blocks differ by at most one game"""
    base, rem = divmod(n_games, world)
    count = base + (1 if rank < rem else 0)
    game0 = rank * base + min(rank, rem)
    return game0, count


def _comm_device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def broadcast_state_dict(sd, src=0):
    """In-place broadcast of a DualNetwork state_dict (216 tensors, 19.1 MB fp32) as ONE flat buffer."""
    dev = _comm_device()
    keys = list(sd.keys())
    flat = torch.cat([sd[k].detach().reshape(-1).to(torch.float32) for k in keys]).to(dev)
    dist.broadcast(flat, src=src)
    off = 0
    for k in keys:
        n = sd[k].numel()
        sd[k].copy_(flat[off:off + n].reshape(sd[k].shape).to(sd[k].dtype))
        off += n
    return sd


def gather_histories(hist, n_local, dst=0):
    """hist: engine.History of this rank (first n_local games valid).  Returns on `dst` a dict of numpy arrays
    (states, counts, actions, lens, final) concatenated in rank order (= global game-id order for
    shard_games blocks); None elsewhere."""
    dev = _comm_device()
    world, rank = dist.get_world_size(), dist.get_rank()
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    counts[rank] = n_local
    dist.all_reduce(counts)
    cap = int(counts.max().item())
    out = {}
    # every array travels as raw bytes (gloo has no int16/uint16 collectives)
    fields = (("states", hist.states, (81, 8), np.uint32), ("counts", hist.counts, (81, 81), np.uint16),
              ("actions", hist.actions, (81,), np.uint8), ("lens", hist.lens, (), np.int32),
              ("final", hist.final, (), np.int8))
    for name, arr, tail, npdt in fields:
        row_bytes = int(np.prod(tail, dtype=np.int64)) * np.dtype(npdt).itemsize
        buf = torch.zeros((cap, row_bytes), dtype=torch.uint8, device=dev)
        if n_local:
            raw = np.ascontiguousarray(arr[:n_local]).view(np.uint8).reshape(n_local, row_bytes)
            buf[:n_local] = torch.from_numpy(raw).to(dev)
        if dist.get_backend() == "nccl":
            parts = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(parts, buf)
        else:
            parts = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
            dist.gather(buf, parts, dst=dst)
        if rank == dst:
            raw = np.concatenate([parts[r][:int(counts[r])].cpu().numpy() for r in range(world)])
            out[name] = np.ascontiguousarray(raw).view(npdt).reshape((raw.shape[0],) + tail)
    if rank != dst:
        return None
    return out


def sharded_self_play(model, n_games, sims=50, batch=8, seed=0, numerics="bf16", engine_obj=None):
    """One self-play cycle over all ranks: broadcast weights from rank 0, play this rank's block of games on
    its GPU, gather the packed histories on rank 0.  Returns (gathered dict or None, local stats)."""
    import engine as _eng
    world, rank = dist.get_world_size(), dist.get_rank()
    broadcast_state_dict(model.state_dict(), src=0)
    game0, count = shard_games(n_games, world, rank)
    e = engine_obj or _eng.Engine(n_slots=min(max(count, 1), 4096), max_sims=sims, max_batch=batch,
                                  max_games=max(count, 1))
    e.upload_model(model)
    ev = _eng.EVAL_NET_FP32 if numerics == "fp32" else _eng.EVAL_NET_BF16
    hist = e.selfplay(count, sims=sims, batch=batch, seed=seed, evaluator=ev, game0=game0)
    gathered = gather_histories(hist, count, dst=0)
    return gathered, hist.stats.copy()
