"""ORACLE -- TEST / BASELINE INFRASTRUCTURE ONLY (imported by tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package).

CPU "port" of the reference's arithmetic for whole-model runs: the layer classes below own the
same parameters as the product layers but their forward does what the reference does every call
-- build the dense expanded weight with torch.cat (quaternion_ops.py:131-135,
dual_quaternion_ops.py:122-140, :170-188) and hand it to stock F.conv1d / F.conv2d / torch.mm --
so that timing it on the host cores times the reference's algorithm, and so that the product's
model assembly can be run on the CPU in float64 as a second parity reference.

The reference tree itself cannot travel to the GPU box (it is not pip-installable and is absent
there), hence kind = "port" in bench.py; tests/test_cpu_port.py pins this port against the
golden fixtures minted from the real reference.
"""
import importlib
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
pkg = importlib.import_module("sound-event-localization-and-detection_b200")

from oracle.algebra import block_table  # noqa: E402


def expand_weight_torch(weights, algebra, linear=False):
    """The torch.cat expansion the reference performs on every forward call."""
    widx, sign = block_table(algebra)
    nc = widx.shape[0]
    zero = None
    rows = []
    for a in range(nc):
        blocks = []
        for b in range(nc):
            e = int(widx[a, b])
            if e < 0:
                if zero is None:
                    zero = torch.zeros_like(weights[0])
                blocks.append(zero)
            else:
                blocks.append(weights[e] if sign[a, b] > 0 else -weights[e])
        # conv: out components stack on dim 0, in components on dim 1; linear (in,out): transposed roles
        rows.append(torch.cat(blocks, dim=0 if linear else 1))
    return torch.cat(rows, dim=1 if linear else 0)


def _conv(x, weights, bias, stride, padding, dilation, algebra):
    W = expand_weight_torch(weights, algebra)
    fn = F.conv1d if x.dim() == 3 else F.conv2d
    return fn(x, W, bias, stride, padding, dilation, 1)


def _linear(x, weights, bias, algebra):
    W = expand_weight_torch(weights, algebra, linear=True)
    y = torch.matmul(x, W)
    return y if bias is None else y + bias


class QuaternionConv(pkg.layers.QuaternionConv):
    def forward(self, x):
        return _conv(x, self._weights(), self.bias, self.stride, self.padding, self.dilatation, "Q")


class DualQuaternionConv(pkg.layers.DualQuaternionConv):
    def forward(self, x):
        return _conv(x, self._weights(), self.bias, self.stride, self.padding, self.dilatation, "DQ")


class QuaternionLinear(pkg.layers.QuaternionLinear):
    def forward(self, x):
        return _linear(x, self._weights(), self.bias, "Q")


class DualQuaternionLinear(pkg.layers.DualQuaternionLinear):
    def forward(self, x):
        return _linear(x, self._weights(), self.bias, "DQ_LINEAR")


class _Lib(object):
    pass


LAYER_LIB = _Lib()
LAYER_LIB.QuaternionConv = QuaternionConv
LAYER_LIB.DualQuaternionConv = DualQuaternionConv
LAYER_LIB.QuaternionLinear = QuaternionLinear
LAYER_LIB.DualQuaternionLinear = DualQuaternionLinear


def build_model(**kwargs):
    """The product's model assembly with the CPU reference arithmetic in every Q/DQ layer."""
    return pkg.SELD_Model(layer_lib=LAYER_LIB, **kwargs)


def seld_loss(sed, doa, target, n_sed=42, sed_weight=1.0, doa_weight=5.0):
    """train.py:186-204 (BCE on SED + 5 * MSE on DOA)."""
    t_sed = torch.flatten(target[:, :, :n_sed], start_dim=1)
    t_doa = torch.flatten(target[:, :, n_sed:], start_dim=1)
    return (F.binary_cross_entropy(torch.flatten(sed, 1), t_sed) * sed_weight
            + F.mse_loss(torch.flatten(doa, 1), t_doa) * doa_weight)


def train_step(model, optimizer, x, target):
    """train.py:546-561."""
    optimizer.zero_grad()
    sed, doa = model(x)
    loss = seld_loss(sed, doa, target)
    loss.backward()
    optimizer.step()
    return loss
