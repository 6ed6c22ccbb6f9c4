"""nn.Module owners of the compact quaternion / dual-quaternion parameters.

Mirror of quaternion/quaternion_layers.py (QuaternionConv :100-172, QuaternionLinear :227-286,
QuaternionLinearAutograd :174-225, QuaternionTransposeConv :19-98) and
dual_quaternion/dual_quaternion_layers.py (DualQuaternionConv :49-135, DualQuaternionLinear :138-206):
same constructor signatures, attribute names, parameter names / shapes / registration order
(state_dict compatibility, train.py:32-45), same initialisation stream, same __repr__.
forward() calls the sm_100a kernels through functional.block_conv / block_linear.
"""
import numpy as np
import torch
from numpy.random import RandomState
from torch.nn import Module
from torch.nn.parameter import Parameter

from . import functional as F
from . import init as I
from ._lib import ALG_DQ, ALG_Q



def _dropin_hook():
    """install_dropin(fuse_model=True): the first layer the reference's model.py constructs wires the fused glue
    kernels into model.TC_Block / ConvTC_Block (package __init__._maybe_patch_reference_model)."""
    import sys
    hook = getattr(sys.modules.get(__package__), "_maybe_patch_reference_model", None)
    if hook is not None:
        hook()


_Q_NAMES = ("r_weight", "i_weight", "j_weight", "k_weight")
_DQ_NAMES = _Q_NAMES + tuple(n + "_2" for n in _Q_NAMES)


def _dq_unitary(*a, **k):
    return I.unitary_init(*a, dual=True, **k)


def _dq_random(*a, **k):
    return I.random_init(*a, dual=True, **k)


_Q_INITS = {"quaternion": I.quaternion_init, "unitary": I.unitary_init, "random": I.random_init}
_DQ_INITS = {"quaternion": I.dual_quaternion_init, "unitary": _dq_unitary, "random": _dq_random}


class _BlockConvBase(Module):
    _ncomp = 4
    _names = _Q_NAMES
    _algebra = ALG_Q
    _inits = _Q_INITS

    def _setup(self, in_channels, out_channels, kernel_size, stride, dilatation, padding, groups, bias,
               init_criterion, weight_init, seed, operation, rotation, quaternion_format):
        self.in_channels = in_channels // self._ncomp
        self.out_channels = out_channels // self._ncomp
        self.stride = stride
        self.padding = padding
        self.groups = groups
        self.dilatation = dilatation
        self.init_criterion = init_criterion
        self.weight_init = weight_init
        self.seed = seed if seed is not None else np.random.randint(0, 1234)
        self.rng = RandomState(self.seed)
        self.operation = operation
        self.rotation = rotation
        self.quaternion_format = quaternion_format
        self.winit = self._inits[self.weight_init]
        self.kernel_size, self.w_shape = I.get_kernel_and_weight_shape(
            self.operation, self.in_channels, self.out_channels, kernel_size)
        for name in self._names:
            setattr(self, name, Parameter(torch.Tensor(*self.w_shape)))

    def _finish(self, bias, out_channels):
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()
        _dropin_hook()

    def _weights(self):
        return tuple(getattr(self, n) for n in self._names)

    def forward(self, input):
        if self.rotation:        # quaternion_layers.py:151-155
            return F.quaternion_conv_rotation(input, self._weights(), self.bias, self.stride, self.padding, self.groups,
                                              self.dilatation, self.quaternion_format)
        if self.groups != 1:
            raise NotImplementedError("seldq: groups != 1 is not implemented")
        return F.block_conv(input, self._weights(), self.bias, self.stride, self.padding, self.dilatation,
                            self._algebra)


class QuaternionConv(_BlockConvBase):
    """Quaternion convolution layer (quaternion_layers.py:100-172)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride,
                 dilatation=1, padding=0, groups=1, bias=True, init_criterion='glorot',
                 weight_init='quaternion', seed=None, operation='convolution2d', rotation=False,
                 quaternion_format=False):
        super(QuaternionConv, self).__init__()
        self._setup(in_channels, out_channels, kernel_size, stride, dilatation, padding, groups, bias,
                    init_criterion, weight_init, seed, operation, rotation, quaternion_format)
        self._finish(bias, out_channels)

    def reset_parameters(self):
        I.affect_init_conv(self._weights(), self.kernel_size, self.winit, self.rng, self.init_criterion)
        if self.bias is not None:
            self.bias.data.zero_()

    def __repr__(self):
        return (self.__class__.__name__ + '(in_channels=' + str(self.in_channels)
                + ', out_channels=' + str(self.out_channels) + ', bias=' + str(self.bias is not None)
                + ', kernel_size=' + str(self.kernel_size) + ', stride=' + str(self.stride)
                + ', padding=' + str(self.padding) + ', dilatation=' + str(self.dilatation)
                + ', init_criterion=' + str(self.init_criterion) + ', weight_init=' + str(self.weight_init)
                + ', seed=' + str(self.seed) + ', operation=' + str(self.operation) + ')')


class DualQuaternionConv(_BlockConvBase):
    """Dual-quaternion convolution layer (dual_quaternion_layers.py:49-135)."""
    _ncomp = 8
    _names = _DQ_NAMES
    _algebra = ALG_DQ
    _inits = _DQ_INITS

    def __init__(self, in_channels, out_channels, kernel_size, stride,
                 dilatation=1, padding=0, groups=1, bias=True, init_criterion='glorot',
                 weight_init='quaternion', seed=None, operation='convolution2d', rotation=False,
                 quaternion_format=True, scale=False):
        super(DualQuaternionConv, self).__init__()
        self.scale = scale
        self._setup(in_channels, out_channels, kernel_size, stride, dilatation, padding, groups, bias,
                    init_criterion, weight_init, seed, operation, rotation, quaternion_format)
        # registration order of the reference: 8 weights, scale_param, zero_kernel, bias
        if self.scale:
            self.scale_param = Parameter(torch.Tensor(self.r_weight.shape))
        else:
            self.scale_param = None
        if self.rotation:
            self.zero_kernel = Parameter(torch.zeros(self.r_weight.shape), requires_grad=False)
        self._finish(bias, out_channels)

    def reset_parameters(self):
        ws = self._weights()
        I.affect_init_conv(ws[:4], self.kernel_size, self.winit, self.rng, self.init_criterion, ws2=ws[4:])
        if self.scale_param is not None:
            torch.nn.init.xavier_uniform_(self.scale_param.data)
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, input):
        # the reference forward ignores `rotation` / `scale_param` (dual_quaternion_layers.py:115-119)
        if self.groups != 1:
            raise NotImplementedError("seldq: groups != 1 is not implemented")
        return F.block_conv(input, self._weights(), self.bias, self.stride, self.padding, self.dilatation,
                            self._algebra)

    def __repr__(self):
        return (self.__class__.__name__ + '(in_channels=' + str(self.in_channels)
                + ', out_channels=' + str(self.out_channels) + ', bias=' + str(self.bias is not None)
                + ', kernel_size=' + str(self.kernel_size) + ', stride=' + str(self.stride)
                + ', padding=' + str(self.padding) + ', init_criterion=' + str(self.init_criterion)
                + ', weight_init=' + str(self.weight_init) + ', seed=' + str(self.seed)
                + ', rotation=' + str(self.rotation) + ', q_format=' + str(self.quaternion_format)
                + ', operation=' + str(self.operation) + ')')


class _BlockLinearBase(Module):
    _ncomp = 4
    _names = _Q_NAMES
    _algebra = ALG_Q
    _inits = _Q_INITS

    def _setup(self, in_features, out_features, bias, init_criterion, weight_init, seed):
        self.in_features = in_features // self._ncomp
        self.out_features = out_features // self._ncomp
        for name in self._names:
            setattr(self, name, Parameter(torch.Tensor(self.in_features, self.out_features)))
        if bias:
            self.bias = Parameter(torch.Tensor(self.out_features * self._ncomp))
        else:
            self.register_parameter("bias", None)
        self.init_criterion = init_criterion
        self.weight_init = weight_init
        self.seed = seed if seed is not None else np.random.randint(0, 1234)
        self.rng = RandomState(self.seed)
        self.reset_parameters()

    def _weights(self):
        return tuple(getattr(self, n) for n in self._names)

    def reset_parameters(self):
        winit = self._inits[self.weight_init]
        if self.bias is not None:
            self.bias.data.fill_(0)
        ws = self._weights()
        I.affect_init(ws[:4], winit, self.rng, self.init_criterion, ws2=ws[4:] if len(ws) == 8 else None)

    def forward(self, input):
        # 3-d (T, N, C) inputs are flattened and restored (quaternion_layers.py:265-270,
        # dual_quaternion_layers.py:183-189); anything else is rejected like the reference does
        if input.dim() == 3:
            T, N, C = input.size()
            out = F.block_linear(input.reshape(T * N, C), self._weights(), self.bias, self._algebra)
            return out.reshape(T, N, out.size(1))
        if input.dim() == 2:
            return F.block_linear(input, self._weights(), self.bias, self._algebra)
        raise NotImplementedError

    def __repr__(self):
        return (self.__class__.__name__ + '(in_features=' + str(self.in_features)
                + ', out_features=' + str(self.out_features) + ', bias=' + str(self.bias is not None)
                + ', init_criterion=' + str(self.init_criterion) + ', weight_init=' + str(self.weight_init)
                + ', seed=' + str(self.seed) + ')')


class QuaternionLinear(_BlockLinearBase):
    """Quaternion linear layer (quaternion_layers.py:227-286; forward = QuaternionLinearFunction)."""

    def __init__(self, in_features, out_features, bias=True,
                 init_criterion='glorot', weight_init='quaternion', seed=None):
        super(QuaternionLinear, self).__init__()
        self._setup(in_features, out_features, bias, init_criterion, weight_init, seed)


class QuaternionLinearAutograd(_BlockLinearBase):
    """quaternion_layers.py:174-225.  Same math as QuaternionLinear (quaternion_linear instead of the
    custom Function); rotation=True runs quaternion_linear_rotation (functional.py)."""

    def __init__(self, in_features, out_features, bias=True,
                 init_criterion='glorot', weight_init='quaternion',
                 seed=None, rotation=False, quaternion_format=False):
        super(QuaternionLinearAutograd, self).__init__()
        self.rotation = rotation
        self.quaternion_format = quaternion_format
        self._setup(in_features, out_features, bias, init_criterion, weight_init, seed)

    def forward(self, input):
        if self.rotation:        # quaternion_layers.py:212-214
            return F.quaternion_linear_rotation(input, self._weights(), self.bias, self.quaternion_format)
        return F.block_linear(input, self._weights(), self.bias, self._algebra)


class DualQuaternionLinear(_BlockLinearBase):
    """Dual-quaternion linear layer (dual_quaternion_layers.py:138-206); note the 'he' default."""
    _ncomp = 8
    _names = _DQ_NAMES
    _algebra = ALG_DQ
    _inits = {"quaternion": I.dual_quaternion_init, "unitary": _dq_unitary}

    def __init__(self, in_features, out_features, bias=True,
                 init_criterion='he', weight_init='quaternion', seed=None):
        super(DualQuaternionLinear, self).__init__()
        self._setup(in_features, out_features, bias, init_criterion, weight_init, seed)


class QuaternionTransposeConv(Module):
    """quaternion_layers.py:19-98 (never instantiated by model.py): same constructor, parameters
    (in / 4, out / 4, k...) and initialisation; forward = quaternion_transpose_conv on the convolution kernels
    (functional.block_conv_transpose), or quaternion_transpose_conv_rotation with rotation=True."""

    def __init__(self, in_channels, out_channels, kernel_size, stride, dilatation=1, padding=0, output_padding=0, groups=1,
                 bias=True, init_criterion='glorot', weight_init='quaternion', seed=None, operation='convolution2d',
                 rotation=False, quaternion_format=False):
        super(QuaternionTransposeConv, self).__init__()
        self.in_channels = in_channels // 4
        self.out_channels = out_channels // 4
        self.stride, self.padding, self.output_padding = stride, padding, output_padding
        self.groups, self.dilatation = groups, dilatation
        self.init_criterion, self.weight_init = init_criterion, weight_init
        self.seed = seed if seed is not None else np.random.randint(0, 1234)
        self.rng = RandomState(self.seed)
        self.operation, self.rotation, self.quaternion_format = operation, rotation, quaternion_format
        self.winit = _Q_INITS[self.weight_init]
        # (out, in) swapped as in quaternion_layers.py:49-50: the weights are (in / 4, out / 4, k...)
        self.kernel_size, self.w_shape = I.get_kernel_and_weight_shape(self.operation, self.out_channels, self.in_channels,
                                                                       kernel_size)
        for name in _Q_NAMES:
            setattr(self, name, Parameter(torch.Tensor(*self.w_shape)))
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        I.affect_init_conv(tuple(getattr(self, n) for n in _Q_NAMES), self.kernel_size, self.winit, self.rng,
                           self.init_criterion)
        if self.bias is not None:
            self.bias.data.zero_()

    def forward(self, input):
        if self.rotation:        # quaternion_layers.py:75-80
            return F.quaternion_transpose_conv_rotation(input, tuple(getattr(self, n) for n in _Q_NAMES), self.bias,
                                                        self.stride, self.padding, self.output_padding, self.groups,
                                                        self.dilatation, self.quaternion_format)
        return F.block_conv_transpose(input, tuple(getattr(self, n) for n in _Q_NAMES), self.bias, self.stride,
                                      self.padding, self.output_padding, self.groups, self.dilatation, ALG_Q)
