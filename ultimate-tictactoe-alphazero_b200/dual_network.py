"""Policy/value ResNet used by the self-play engine -- weight owner, trainer model and fp32 numerics reference.

The parameter names, shapes and forward semantics are those of the reference's `DualNetwork`
(dual_network.py:28-121 there), so `./model/best.pth` / `./model/latest.pth` files are interchangeable:
216 state_dict entries, 4,765,338 parameters.  Self-play inference does NOT run through this module: the
engine folds the BatchNorms, repacks the weights (`Engine.upload_model`) and evaluates positions with its own
kernels (csrc/net_tc.cu, csrc/net_tc2.cu, csrc/net_fp32.cu).
"""
import os

import torch
from torch import nn
from torch.nn import functional as F

# architecture constants, same names as the reference module (dual_network.py:12-15)
DN_FILTERS = 128
DN_RESIDUAL_NUM = 16
DN_INPUT_SHAPE = (9, 9, 3)      # H, W, C of the board planes
DN_OUTPUT_SIZE = 81

device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")

BEST_PATH = "./model/best.pth"


def _conv(cin, cout, k):
    return nn.Conv2d(cin, cout, kernel_size=k, padding=k // 2, bias=False)


def _reset(module):
    """Kaiming-normal (fan_out, relu) for conv / linear weights, zero biases, identity BatchNorm
    (the scheme of dual_network.py:77-87)."""
    if isinstance(module, (nn.Conv2d, nn.Linear)):
        nn.init.kaiming_normal_(module.weight, mode="fan_out", nonlinearity="relu")
        if module.bias is not None:
            nn.init.zeros_(module.bias)
    elif isinstance(module, nn.BatchNorm2d):
        nn.init.ones_(module.weight)
        nn.init.zeros_(module.bias)


class ResidualBlock(nn.Module):
    """x -> relu(bn2(conv2(relu(bn1(conv1(x))))) + x)"""

    def __init__(self, filters):
        super().__init__()
        self.conv1, self.bn1 = _conv(filters, filters, 3), nn.BatchNorm2d(filters)
        self.conv2, self.bn2 = _conv(filters, filters, 3), nn.BatchNorm2d(filters)

    def forward(self, x):
        h = torch.relu(self.bn1(self.conv1(x)))
        return torch.relu(self.bn2(self.conv2(h)) + x)


class DualNetwork(nn.Module):
    def __init__(self, input_shape=DN_INPUT_SHAPE, filters=DN_FILTERS, residual_num=DN_RESIDUAL_NUM,
                 output_size=DN_OUTPUT_SIZE):
        super().__init__()
        height, width, planes = input_shape
        cells = height * width
        # stem
        self.conv_input, self.bn_input = _conv(planes, filters, 3), nn.BatchNorm2d(filters)
        # tower
        self.residual_blocks = nn.ModuleList([ResidualBlock(filters) for _ in range(residual_num)])
        # policy head: 1x1 conv to 2 planes, FC to the action logits
        self.policy_conv, self.policy_bn = _conv(filters, 2, 1), nn.BatchNorm2d(2)
        self.policy_fc = nn.Linear(2 * cells, output_size)
        # value head: 1x1 conv to 1 plane, FC 256, FC 1
        self.value_conv, self.value_bn = _conv(filters, 1, 1), nn.BatchNorm2d(1)
        self.value_fc1 = nn.Linear(cells, 256)
        self.value_fc2 = nn.Linear(256, 1)
        self.apply(_reset)

    def trunk(self, x):
        x = torch.relu(self.bn_input(self.conv_input(x)))
        for block in self.residual_blocks:
            x = block(x)
        return x

    def forward(self, x):
        """(N,3,9,9) planes -> (softmax policy (N,81), tanh value (N,1))"""
        feat = self.trunk(x)
        logits = self.policy_fc(torch.relu(self.policy_bn(self.policy_conv(feat))).flatten(1))
        hidden = torch.relu(self.value_fc1(torch.relu(self.value_bn(self.value_conv(feat))).flatten(1)))
        return F.softmax(logits, dim=1), torch.tanh(self.value_fc2(hidden))


def dual_network():
    """Write a freshly initialised network to ./model/best.pth unless that file already exists."""
    if os.path.exists(BEST_PATH):
        return
    os.makedirs(os.path.dirname(BEST_PATH), exist_ok=True)
    torch.save(DualNetwork().state_dict(), BEST_PATH)
    print("Model saved to '%s'" % BEST_PATH)


if __name__ == "__main__":
    dual_network()
