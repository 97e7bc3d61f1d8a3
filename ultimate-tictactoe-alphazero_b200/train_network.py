"""train_network.py -- drop-in for the reference's parameter update step (train_network.py:21-121 there), SURVEY 8(f)-2.

Same data source (`./data/*.history`, newest file), model file protocol (`./model/best.pth` in, `./model/latest.pth`
out), loss (policy cross entropy on the softmax output + value MSE, `:69-76`), optimizer (Adam, lr 1e-3), schedule
(x0.5 from epoch 50, x0.25 from epoch 80, `:79-87`), epochs / batch size and progress lines.  What changes is the feed:

  * the whole data set lives on the training device as three tensors ((N,3,9,9), (N,81), (N,1) float32) instead of going
    through a `Dataset` / `DataLoader` that collates 128 numpy rows per step on the host and copies them over;
    `train_tensors()` takes such tensors directly (e.g. from `self_play_cpp.history_tensors`, no pickle round trip);
  * the batch order is the `DataLoader(shuffle=True)` order of the reference, bit for bit: per epoch the loader draws
    one int64 from the global torch RNG for its base seed and the `RandomSampler` a second one that seeds the
    `torch.randperm` (torch/utils/data/dataloader.py `_BaseDataLoaderIter.__init__`, sampler.py `RandomSampler.__iter__`).
    On the same device type and seed the trained weights are therefore identical to the reference trainer's
    (tests/test_train_cpu.py checks that against the reference module itself, on the CPU).

Opt-in, not reference numerics: `UTTT_TRAIN_BF16=1` runs the forward / backward under bf16 autocast in channels_last.
The trainer is PyTorch code (the reference's is too); no self-play kernel is involved.
"""
import os
import pickle
from pathlib import Path

import numpy as np
import torch
from torch import nn, optim

from dual_network import DualNetwork, device

RN_EPOCHS = 100          # train_network.py:16-18
BATCH_SIZE = 128
NUM_WORKERS = 0


def load_data():
    """train_network.py:21-24"""
    history_path = sorted(Path("./data").glob("*.history"))[-1]
    with history_path.open(mode="rb") as f:
        return pickle.load(f)


def history_to_tensors(history, dev=None):
    """the reference's `zip(*history)` + `np.array` + `HistoryDataset` conversion (train_network.py:44-49, 27-33)
    -> xs (N,3,9,9), policies (N,81), values (N,1), float32 on `dev`"""
    dev = device if dev is None else dev
    xs, y_policies, y_values = zip(*history)
    xs = np.transpose(np.array(xs), (0, 3, 1, 2)).astype(np.float32)
    y_policies = np.array(y_policies).astype(np.float32)
    y_values = np.array(y_values).astype(np.float32).reshape(-1, 1)
    return (torch.from_numpy(np.ascontiguousarray(xs)).to(dev), torch.from_numpy(y_policies).to(dev),
            torch.from_numpy(y_values).to(dev))


def _loader_epoch_order(n):
    """the index order one `for batch in DataLoader(dataset, shuffle=True)` pass visits, consuming the global torch
    RNG exactly as the loader does"""
    torch.empty((), dtype=torch.int64).random_()                              # _BaseDataLoaderIter._base_seed
    seed = int(torch.empty((), dtype=torch.int64).random_().item())           # RandomSampler.__iter__
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g)


def policy_loss_fn(pred, target):
    """train_network.py:69-72: cross entropy between the target distribution and the (softmax) policy output"""
    return -torch.sum(target * torch.log(pred + 1e-8)) / pred.size(0)


def lr_lambda(epoch):
    """train_network.py:79-85"""
    if epoch >= 80:
        return 0.25
    elif epoch >= 50:
        return 0.5
    return 1.0


def _eager_epochs(model, xs, y_policies, y_values, epochs, batch_size, bf16, log):
    """train_network.py:89-113, literally, over device-resident tensors"""
    dev = xs.device
    criterion_value = nn.MSELoss()
    optimizer = optim.Adam(model.parameters(), lr=0.001)
    scheduler = optim.lr_scheduler.LambdaLR(optimizer, lr_lambda=lr_lambda)
    n = xs.shape[0]
    n_batches = (n + batch_size - 1) // batch_size
    losses = []
    for epoch in range(epochs):
        order = _loader_epoch_order(n).to(dev)
        total = torch.zeros((), dtype=torch.float64, device=dev)              # one host read per epoch, not per step
        for b in range(n_batches):
            idx = order[b * batch_size:(b + 1) * batch_size]
            inputs, target_policies, target_values = xs[idx], y_policies[idx], y_values[idx]
            optimizer.zero_grad()
            with torch.autocast(device_type=dev.type, dtype=torch.bfloat16, enabled=bf16):
                pred_policies, pred_values = model(inputs)
            loss = policy_loss_fn(pred_policies.float(), target_policies) + criterion_value(pred_values.float(), target_values)
            loss.backward()
            optimizer.step()
            total += loss.detach().double()
        scheduler.step()
        avg_loss = float(total) / n_batches
        losses.append(avg_loss)
        log(f"Epoch {epoch + 1}/{epochs}, Loss: {avg_loss:.4f}, LR: {scheduler.get_last_lr()[0]:.6f}")
    return losses


def _graphed_epochs(model, xs, y_policies, y_values, epochs, batch_size, bf16, log):
    """The same loop with the full-size step (gather the batch, zero_grad, forward, loss, backward, Adam) captured
    ONCE in a CUDA graph and replayed: the eager step of this network is launch-bound (~7.6 ms for ~1 ms of GPU work on a
    B200).  Same batch order, loss, optimizer and schedule; Adam runs in its `capturable` form (step count and learning
    rate live on the device), the ragged last batch of an epoch runs eagerly through the same optimizer."""
    dev = xs.device
    criterion_value = nn.MSELoss()
    lr = torch.tensor(0.001, device=dev)
    optimizer = optim.Adam(model.parameters(), lr=lr, capturable=True)
    scheduler = optim.lr_scheduler.LambdaLR(optimizer, lr_lambda=lr_lambda)
    n = xs.shape[0]
    n_full, n_batches = n // batch_size, (n + batch_size - 1) // batch_size
    s_idx = torch.zeros(batch_size, dtype=torch.int64, device=dev)
    total = torch.zeros((), dtype=torch.float64, device=dev)

    def step(idx):
        inputs, target_policies, target_values = xs[idx], y_policies[idx], y_values[idx]
        optimizer.zero_grad(set_to_none=False)
        with torch.autocast(device_type=dev.type, dtype=torch.bfloat16, enabled=bf16):
            pred_policies, pred_values = model(inputs)
        loss = policy_loss_fn(pred_policies.float(), target_policies) + criterion_value(pred_values.float(), target_values)
        loss.backward()
        optimizer.step()
        total.add_(loss.detach().double())

    graph = None
    if n_full > 0:
        # warm-up + capture must not count as training: snapshot, then restore weights, BN statistics and Adam state
        snap = {k: v.clone() for k, v in model.state_dict().items()}
        s_idx.copy_(torch.arange(batch_size, device=dev) % n)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                step(s_idx)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step(s_idx)
        with torch.no_grad():
            model.load_state_dict(snap)
            for st in optimizer.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()
            total.zero_()
    losses = []
    for epoch in range(epochs):
        order = _loader_epoch_order(n).to(dev)
        total.zero_()
        for b in range(n_full):
            s_idx.copy_(order[b * batch_size:(b + 1) * batch_size])
            graph.replay()
        if n_batches > n_full:
            step(order[n_full * batch_size:])
        scheduler.step()
        avg_loss = float(total) / n_batches
        losses.append(avg_loss)
        log(f"Epoch {epoch + 1}/{epochs}, Loss: {avg_loss:.4f}, LR: {float(scheduler.get_last_lr()[0]):.6f}")
    return losses


def train_tensors(model, xs, y_policies, y_values, epochs=None, batch_size=None, bf16=None, graph=None, log=print):
    """The training loop of train_network.py:89-113 over device-resident tensors.  Returns the per-epoch mean losses.
    graph (opt-in: `UTTT_TRAIN_GRAPH=1` or graph=True; CUDA tensors only): replay the step as a CUDA graph with
    Adam(capturable=True).  The default is the eager loop, whose arithmetic is the reference trainer's (bit-identical
    weights on the CPU, tests/test_train_cpu.py); the graph path orders Adam's arithmetic differently (same losses to
    ~4 digits, tools/train_equiv.py) and needs torch >= 2.1 (tensor learning rate)."""
    epochs = RN_EPOCHS if epochs is None else epochs
    batch_size = BATCH_SIZE if batch_size is None else batch_size
    bf16 = (os.environ.get("UTTT_TRAIN_BF16", "0") == "1") if bf16 is None else bf16
    graph = (os.environ.get("UTTT_TRAIN_GRAPH", "0") == "1") if graph is None else graph
    if bf16:
        model = model.to(memory_format=torch.channels_last)
        xs = xs.contiguous(memory_format=torch.channels_last)
    model.train()
    loop = _graphed_epochs if (graph and xs.device.type == "cuda") else _eager_epochs
    return loop(model, xs, y_policies, y_values, epochs, batch_size, bf16, log)


def load_tensors():
    """the newest cycle's samples as device tensors: from the packed sidecar `<timestamp>.packed.npz` of the newest
    `.history` file if self-play wrote one (self_play_cpp.SP_WRITE_PACKED; 357 B per sample, planes re-encoded on the
    GPU), else from the reference's pickle itself.  The sidecar's policy targets are fp32 counts / sum; the pickle's are
    that value re-normalised in float64 and cast back (self_play_cpp.py:74-78 -> train_network.py:51): equal to ~1e-7,
    not bit for bit -- delete the sidecar (or leave SP_WRITE_PACKED off, the default) for the reference's exact targets."""
    history_path = sorted(Path("./data").glob("*.history"))[-1]
    sidecar = Path(str(history_path).replace(".history", ".packed.npz"))
    if sidecar.exists() and device.type == "cuda":
        from self_play_cpp import load_packed_history
        xs, pis, zs = load_packed_history(str(sidecar))
        return (torch.from_numpy(xs).to(device), torch.from_numpy(pis).to(device),
                torch.from_numpy(zs.astype(np.float32)).reshape(-1, 1).to(device))
    return history_to_tensors(load_data())


def train_network():
    """train_network.py:41-121"""
    xs, y_policies, y_values = load_tensors()
    model = DualNetwork().to(device)
    model.load_state_dict(torch.load("./model/best.pth", map_location=device, weights_only=True))
    train_tensors(model, xs, y_policies, y_values)
    torch.save({k: v.contiguous() for k, v in model.state_dict().items()}, "./model/latest.pth")
    print("Model saved to ./model/latest.pth")
    del model


if __name__ == "__main__":
    train_network()
