"""ORACLE -- TEST INFRASTRUCTURE ONLY.

What an ideal "bf16 operands, wide accumulation" implementation of the Q/DQ convolutions computes:
the CPU port of the model (oracle/cpu_model.py) in float64, except that every convolution rounds
its operands to bfloat16 first -- x and the (expanded) weight in the forward pass, gy and the
weight in dgrad, gy and x in wgrad -- exactly the roundings the tcgen05 path performs (the linear
layers and everything between the convolutions stay exact, as they do on the GPU in fp32).

Why it exists: whole-network gradients of this model are badly conditioned (train-mode BatchNorm
over short sequences, gated tanh/sigmoid, 10 residual blocks), so bf16 rounding noise that is
2.7e-3 per convolution grows to tens of percent on some parameter gradients.  Comparing the GPU's
bf16 gradients with the float64 reference therefore measures the model's conditioning, not the
kernels.  Comparing them with THIS emulation isolates the kernels: a correct bf16 implementation
must agree with it to within accumulation-order noise.
"""
import torch
import torch.nn.functional as F

from oracle import cpu_model

# True: 2-d convolutions also round their OUTPUT to IEEE fp16 (identity gradient), which is what the fused
# CNN-block path does when it stores the conv output once in 16 bits for BatchNorm / ReLU / max-pool (fused.py;
# fp16 rather than bf16 because the tensor is only stored, never a tensor-core operand)
STORE_CONV2D_F16 = False


def _bf(t):
    return t.to(torch.bfloat16).to(t.dtype)


class Bf16OperandConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, stride, pad, dil):
        ctx.save_for_backward(x, W)
        ctx.args = (stride, pad, dil)
        fn = F.conv1d if x.dim() == 3 else F.conv2d
        y = fn(_bf(x), _bf(W), None, stride, pad, dil)
        return y.to(torch.float16).to(y.dtype) if (x.dim() == 4 and STORE_CONV2D_F16) else y

    @staticmethod
    def backward(ctx, gy):
        x, W = ctx.saved_tensors
        stride, pad, dil = ctx.args
        g = _bf(gy)
        if x.dim() == 3:
            gx = torch.nn.grad.conv1d_input(x.shape, _bf(W), g, stride, pad, dil)
            gW = torch.nn.grad.conv1d_weight(_bf(x), W.shape, g, stride, pad, dil)
        else:
            gx = torch.nn.grad.conv2d_input(x.shape, _bf(W), g, stride, pad, dil)
            gW = torch.nn.grad.conv2d_weight(_bf(x), W.shape, g, stride, pad, dil)
        return gx, gW, None, None, None


class _Bf16Convs(object):
    """Context manager: the CPU port's convolutions round their operands to bf16."""

    def __init__(self, store_conv2d_f16=False):
        self._store = store_conv2d_f16

    def __enter__(self):
        global STORE_CONV2D_F16
        self._saved = cpu_model._conv
        self._saved_flag, STORE_CONV2D_F16 = STORE_CONV2D_F16, self._store

        def conv(x, weights, bias, stride, padding, dilation, algebra):
            y = Bf16OperandConv.apply(x, cpu_model.expand_weight_torch(weights, algebra), stride, padding, dilation)
            if bias is not None:
                y = y + bias.view(1, -1, *([1] * (y.dim() - 2)))
            return y
        cpu_model._conv = conv
        return self

    def __exit__(self, *exc):
        global STORE_CONV2D_F16
        cpu_model._conv = self._saved
        STORE_CONV2D_F16 = self._saved_flag


def bf16_operand_convs(store_conv2d_f16=False):
    return _Bf16Convs(store_conv2d_f16)
