"""Trainer feed measurement: seconds per epoch of the reference's loop (Dataset + DataLoader, one host
collate + H2D copy + loss.item() per step; restated from train_network.py:27-113) vs train_network.train_tensors
(device-resident tensors) in fp32 and under bf16 autocast, on one cycle's worth of samples (29,000 x (3,9,9))."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ultimate-tictactoe-alphazero_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch import nn, optim  # noqa: E402
from torch.utils.data import DataLoader, Dataset  # noqa: E402
import train_network as tn  # noqa: E402
from dual_network import DualNetwork  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 29000
dev = torch.device("cuda")
rng = np.random.RandomState(0)
xs = (rng.rand(n, 9, 9, 3) < 0.3).astype(np.float32)
ps = rng.rand(n, 81); ps /= ps.sum(1, keepdims=True)
zs = rng.randint(-1, 2, size=n)


class HistoryDataset(Dataset):                      # train_network.py:27-38
    def __init__(self, xs, y_policies, y_values):
        self.xs = np.transpose(xs, (0, 3, 1, 2)).astype(np.float32)
        self.y_policies = y_policies.astype(np.float32)
        self.y_values = y_values.astype(np.float32).reshape(-1, 1)

    def __len__(self):
        return len(self.xs)

    def __getitem__(self, idx):
        return self.xs[idx], self.y_policies[idx], self.y_values[idx]


def reference_epoch(model, loader, optimizer, crit):
    total = 0.0
    for inputs, tp, tv in loader:
        inputs, tp, tv = inputs.to(dev), tp.to(dev), tv.to(dev)
        optimizer.zero_grad()
        pp, pv = model(inputs)
        loss = tn.policy_loss_fn(pp, tp) + crit(pv, tv)
        loss.backward()
        optimizer.step()
        total += loss.item()
    return total / len(loader)


out = {"samples": n, "batch_size": tn.BATCH_SIZE, "steps_per_epoch": (n + 127) // 128}
torch.manual_seed(0)
base = DualNetwork().to(dev)
# reference-style feed
model = DualNetwork().to(dev); model.load_state_dict(base.state_dict()); model.train()
loader = DataLoader(HistoryDataset(xs, ps, zs), batch_size=128, shuffle=True, num_workers=0, pin_memory=True)
opt = optim.Adam(model.parameters(), lr=0.001); crit = nn.MSELoss()
reference_epoch(model, loader, opt, crit)           # warm-up epoch (cudnn autotune)
torch.cuda.synchronize(); t0 = time.perf_counter()
loss = reference_epoch(model, loader, opt, crit)
torch.cuda.synchronize(); out["reference_loop_s_per_epoch"] = time.perf_counter() - t0
out["reference_loop_loss"] = loss
# device-resident feed
x, p, z = tn.history_to_tensors([[xs[i], ps[i], int(zs[i])] for i in range(n)], dev)
for name, bf16, graph in (("train_tensors_fp32_eager", False, False), ("train_tensors_fp32_graph", False, True),
                          ("train_tensors_bf16_eager", True, False), ("train_tensors_bf16_graph", True, True)):
    model = DualNetwork().to(dev); model.load_state_dict(base.state_dict())
    torch.manual_seed(5)
    stamps = []

    def log(_):
        torch.cuda.synchronize(); stamps.append(time.perf_counter())
    torch.cuda.synchronize(); t0 = time.perf_counter()
    losses = tn.train_tensors(model, x, p, z, epochs=3, bf16=bf16, graph=graph, log=log)
    out[name + "_s_per_epoch"] = stamps[2] - stamps[1]              # third epoch: no warm-up, no graph capture
    out[name + "_first_epoch_s"] = stamps[0] - t0
    out[name + "_losses"] = losses
print(json.dumps(out))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "train_points.json"), "w"), indent=1)
