"""DualNetwork: the 128-filter x 16-block policy/value ResNet of the reference, re-declared so that
`state_dict()` keys, shapes and forward semantics match dual_network.py:28-121 of the reference
(./model/best.pth and ./model/latest.pth stay interchangeable).

In this repository the module is the owner of the weights (training stays PyTorch) and the fp32
numerics reference for the CUDA forward; self-play inference itself runs in the engine's kernels
(csrc/net_tc.cu, csrc/net_fp32.cu) after `Engine.upload_model(model)`.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

DN_FILTERS = 128          # dual_network.py:12
DN_RESIDUAL_NUM = 16      # dual_network.py:13
DN_INPUT_SHAPE = (9, 9, 3)  # (H, W, C), dual_network.py:14
DN_OUTPUT_SIZE = 81       # dual_network.py:15

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")


class ResidualBlock(nn.Module):
    """conv3x3-BN-ReLU-conv3x3-BN, skip, ReLU (dual_network.py:28-45)."""

    def __init__(self, filters):
        super().__init__()
        self.conv1 = nn.Conv2d(filters, filters, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(filters)
        self.conv2 = nn.Conv2d(filters, filters, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(filters)

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return F.relu(y + x)


class DualNetwork(nn.Module):
    def __init__(self, input_shape=DN_INPUT_SHAPE, filters=DN_FILTERS, residual_num=DN_RESIDUAL_NUM,
                 output_size=DN_OUTPUT_SIZE):
        super().__init__()
        h, w, c = input_shape
        self.conv_input = nn.Conv2d(c, filters, 3, padding=1, bias=False)
        self.bn_input = nn.BatchNorm2d(filters)
        self.residual_blocks = nn.ModuleList(ResidualBlock(filters) for _ in range(residual_num))
        self.policy_conv = nn.Conv2d(filters, 2, 1, bias=False)
        self.policy_bn = nn.BatchNorm2d(2)
        self.policy_fc = nn.Linear(2 * h * w, output_size)
        self.value_conv = nn.Conv2d(filters, 1, 1, bias=False)
        self.value_bn = nn.BatchNorm2d(1)
        self.value_fc1 = nn.Linear(h * w, 256)
        self.value_fc2 = nn.Linear(256, 1)
        self._initialize_weights()

    def _initialize_weights(self):
        # same scheme as dual_network.py:77-87: Kaiming-normal(fan_out) on convs and linears, identity BN
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if getattr(m, "bias", None) is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        """x: (N,3,9,9) -> (policy (N,81) softmax, value (N,1) tanh)   dual_network.py:89-121"""
        x = F.relu(self.bn_input(self.conv_input(x)))
        for block in self.residual_blocks:
            x = block(x)
        p = F.relu(self.policy_bn(self.policy_conv(x)))
        p = F.softmax(self.policy_fc(torch.flatten(p, 1)), dim=1)
        v = F.relu(self.value_bn(self.value_conv(x)))
        v = F.relu(self.value_fc1(torch.flatten(v, 1)))
        v = torch.tanh(self.value_fc2(v))
        return p, v


def dual_network():
    """Create ./model/best.pth with a fresh network unless it already exists (dual_network.py:124-135)."""
    if os.path.exists("./model/best.pth"):
        return
    model = DualNetwork()
    os.makedirs("./model/", exist_ok=True)
    torch.save(model.state_dict(), "./model/best.pth")
    print("Model saved to './model/best.pth'")


if __name__ == "__main__":
    dual_network()
