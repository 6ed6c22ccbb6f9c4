"""Drop-in for the reference's quaternion/quaternion_ops.py: same public names and positional
signatures, computed by libseldq.so (sm_100a).  Functions that model.py never reaches
(transpose / rotation variants, hamilton_product) keep their names and raise NotImplementedError
(SURVEY.md section 2, row 1)."""
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


def _out_of_scope(name):
    def fn(*args, **kwargs):
        raise NotImplementedError("seldq: %s is never called by the SELD models and is not implemented "
                                  "(SURVEY.md section 2 / 8f N4)" % name)
    fn.__name__ = name
    return fn


def quaternion_transpose_conv(input, r_weight, i_weight, j_weight, k_weight, bias, stride, padding, output_padding, groups,
                              dilatation):
    """quaternion_ops.py:149-172 on the convolution kernels (stride 1, groups 1; functional.block_conv_transpose)."""
    return _F.block_conv_transpose(input, (r_weight, i_weight, j_weight, k_weight), bias, stride, padding, output_padding,
                                   groups, dilatation, _ALG_Q)


quaternion_conv_rotation = _out_of_scope("quaternion_conv_rotation")
quaternion_transpose_conv_rotation = _out_of_scope("quaternion_transpose_conv_rotation")
quaternion_linear_rotation = _out_of_scope("quaternion_linear_rotation")
hamilton_product = _out_of_scope("hamilton_product")

unitary_init = _I.unitary_init
random_init = _I.random_init
quaternion_init = _I.quaternion_init
get_kernel_and_weight_shape = _I.get_kernel_and_weight_shape


def affect_init(r_weight, i_weight, j_weight, k_weight, init_func, rng, init_criterion):
    _I.affect_init((r_weight, i_weight, j_weight, k_weight), init_func, rng, init_criterion)


def affect_init_conv(r_weight, i_weight, j_weight, k_weight, kernel_size, init_func, rng, init_criterion):
    _I.affect_init_conv((r_weight, i_weight, j_weight, k_weight), kernel_size, init_func, rng, init_criterion)
