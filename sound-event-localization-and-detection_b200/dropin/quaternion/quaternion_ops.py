"""Drop-in for the reference's quaternion/quaternion_ops.py: same public names and positional
signatures, computed by libseldq.so (sm_100a) -- including the functions model.py never reaches
(transpose / rotation variants, hamilton_product)."""
import os as _os
import sys as _sys

_sys.path.append(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from _seldq_pkg import pkg as _pkg  # noqa: E402

import torch  # noqa: E402

_F = _pkg.functional
_I = _pkg.init
_ALG_Q = _pkg._lib.ALG_Q


def check_input(input):
    # quaternion_ops.py:53-65
    if input.dim() not in {2, 3}:
        raise RuntimeError("quaternion linear accepts only input of dimension 2 or 3."
                           " input.dim = " + str(input.dim()))
    nb_hidden = input.size()[-1]
    if nb_hidden % 4 != 0:
        raise RuntimeError("Quaternion Tensors must be divisible by 4."
                           " input.size()[1] = " + str(nb_hidden))


def _component(input, idx):
    check_input(input)
    n = input.size()[-1] // 4
    return input.narrow(input.dim() - 1, idx * n, n)


def get_r(input):
    return _component(input, 0)


def get_i(input):
    return _component(input, 1)


def get_j(input):
    return _component(input, 2)


def get_k(input):
    return _component(input, 3)


def get_modulus(input, vector_form=False):
    r, i, j, k = get_r(input), get_i(input), get_j(input), get_k(input)
    sq = r * r + i * i + j * j + k * k
    return torch.sqrt(sq) if vector_form else torch.sqrt(sq.sum(dim=0))


def get_normalized(input, eps=0.0001):
    check_input(input)
    m = get_modulus(input)
    rep = m.repeat(1, 4) if input.dim() == 2 else m.repeat(1, 1, 4)
    return input / (rep.expand_as(input) + eps)


def quaternion_conv(input, r_weight, i_weight, j_weight, k_weight, bias, stride,
                    padding, groups, dilatation):
    """quaternion_ops.py:125-147 -- the Hamilton expansion is fused into the kernels."""
    if groups != 1:
        raise NotImplementedError("seldq: groups != 1 is not implemented")
    return _F.block_conv(input, (r_weight, i_weight, j_weight, k_weight), bias, stride, padding, dilatation,
                         _ALG_Q)


def quaternion_linear(input, r_weight, i_weight, j_weight, k_weight, bias=True):
    """quaternion_ops.py:299-327 (`bias` is a tensor or None, as the layers call it)."""
    if bias is True:
        bias = None
    return _F.block_linear(input, (r_weight, i_weight, j_weight, k_weight), bias, _ALG_Q)


class QuaternionLinearFunction(object):
    """quaternion_ops.py:392-464: the reference's hand-derived backward equals the autograd of
    quaternion_linear; both map onto the same kernels here."""

    @staticmethod
    def apply(input, r_weight, i_weight, j_weight, k_weight, bias=None):
        check_input(input)
        return _F.block_linear(input, (r_weight, i_weight, j_weight, k_weight), bias, _ALG_Q)


def quaternion_transpose_conv(input, r_weight, i_weight, j_weight, k_weight, bias, stride, padding, output_padding, groups,
                              dilatation):
    """quaternion_ops.py:149-172 on the convolution kernels (groups 1; functional.block_conv_transpose)."""
    return _F.block_conv_transpose(input, (r_weight, i_weight, j_weight, k_weight), bias, stride, padding, output_padding,
                                   groups, dilatation, _ALG_Q)


def quaternion_conv_rotation(input, r_weight, i_weight, j_weight, k_weight, bias, stride,
                             padding, groups, dilatation, quaternion_format):
    """quaternion_ops.py:174-232: the rotation weight (csrc/rotation.cu) feeds the real-algebra convolution kernels."""
    return _F.quaternion_conv_rotation(input, (r_weight, i_weight, j_weight, k_weight), bias, stride, padding, groups,
                                       dilatation, quaternion_format)


def quaternion_transpose_conv_rotation(input, r_weight, i_weight, j_weight, k_weight, bias, stride,
                                       padding, output_padding, groups, dilatation, quaternion_format):
    """quaternion_ops.py:235-295 (groups 1)."""
    return _F.quaternion_transpose_conv_rotation(input, (r_weight, i_weight, j_weight, k_weight), bias, stride, padding,
                                                 output_padding, groups, dilatation, quaternion_format)


def quaternion_linear_rotation(input, r_weight, i_weight, j_weight, k_weight, bias=None, quaternion_format=False):
    """quaternion_ops.py:330-388."""
    return _F.quaternion_linear_rotation(input, (r_weight, i_weight, j_weight, k_weight), bias, quaternion_format)


def hamilton_product(q0, q1):
    """quaternion_ops.py:467-507: (batch_size, 4 n) operands, as its docstring states (the 3-d inputs check_input would
    let through slice the last dimension but concatenate along dimension 1 there)."""
    check_input(q0)
    check_input(q1)
    if q0.dim() != 2:
        raise NotImplementedError("seldq: hamilton_product takes (batch_size, quaternion_number) operands")
    return _F.hamilton_product(q0, q1)


unitary_init = _I.unitary_init
random_init = _I.random_init
quaternion_init = _I.quaternion_init
get_kernel_and_weight_shape = _I.get_kernel_and_weight_shape
create_dropout_mask = _I.create_dropout_mask


def affect_init(r_weight, i_weight, j_weight, k_weight, init_func, rng, init_criterion):
    _I.affect_init((r_weight, i_weight, j_weight, k_weight), init_func, rng, init_criterion)


def affect_init_conv(r_weight, i_weight, j_weight, k_weight, kernel_size, init_func, rng, init_criterion):
    _I.affect_init_conv((r_weight, i_weight, j_weight, k_weight), kernel_size, init_func, rng, init_criterion)
