"""Drop-in for the reference's dual_quaternion/dual_quaternion_layers.py (star-imported by model.py:7)."""
import os
import sys

sys.path.append(os.path.join(os.path.dirname(__file__), '..', 'dual_quaternion'))   # as dual_quaternion_layers.py:9-11
from dual_quaternion_ops import *  # noqa: E402,F401,F403
from dual_quaternion_ops import _pkg  # noqa: E402

DualQuaternionConv = _pkg.DualQuaternionConv
DualQuaternionLinear = _pkg.DualQuaternionLinear


# The real-valued helper blocks of dual_quaternion_layers.py:19-47 (nothing in the reference instantiates them):
# depthwise convolution -> 1 x 1 convolution -> BatchNorm -> ReLU, plain torch modules with the reference's sub-module
# names (depthwise, pointwise, bn, relu), so their state dicts interchange.
import torch.nn as _nn  # noqa: E402


class _DepthwiseSeparable(_nn.Module):
    _conv, _bn = None, None

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0):
        super(_DepthwiseSeparable, self).__init__()
        self.depthwise = self._conv(in_channels, in_channels, kernel_size, stride, padding, groups=in_channels)
        self.pointwise = self._conv(in_channels, out_channels, kernel_size=1)
        self.bn = self._bn(out_channels)
        self.relu = _nn.ReLU()

    def forward(self, x):
        return self.relu(self.bn(self.pointwise(self.depthwise(x))))


class DepthwiseSeparableConv2D(_DepthwiseSeparable):
    _conv, _bn = _nn.Conv2d, _nn.BatchNorm2d


class DepthwiseSeparableConv1D(_DepthwiseSeparable):
    _conv, _bn = _nn.Conv1d, _nn.BatchNorm1d
