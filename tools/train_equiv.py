"""per-step loss of the eager loop vs the CUDA-graph loop of train_network.train_tensors on the same data / seed"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ultimate-tictactoe-alphazero_b200"))
import numpy as np, torch
import train_network as tn
from dual_network import DualNetwork
dev = torch.device("cuda")
n = 128 * 12 + 50
rng = np.random.RandomState(0)
xs = torch.from_numpy((rng.rand(n, 3, 9, 9) < 0.3).astype(np.float32)).to(dev)
ps = rng.rand(n, 81); ps /= ps.sum(1, keepdims=True)
ps = torch.from_numpy(ps.astype(np.float32)).to(dev)
zs = torch.from_numpy(rng.randint(-1, 2, size=(n, 1)).astype(np.float32)).to(dev)
torch.manual_seed(0)
base = DualNetwork().to(dev).state_dict()
res = {}
for graph in (False, True):
    m = DualNetwork().to(dev); m.load_state_dict(base)
    torch.manual_seed(5)
    res[graph] = tn.train_tensors(m, xs, ps, zs, epochs=4, bf16=False, graph=graph, log=lambda s: None)
    res[graph, "w"] = m.state_dict()["conv_input.weight"].clone()
print("eager", res[False]); print("graph", res[True])
print("max |dW| conv_input:", float((res[False, "w"] - res[True, "w"]).abs().max()), "scale", float(res[False, "w"].abs().max()))
