"""Drop-in for the reference's dual_quaternion/dual_quaternion_ops.py: same public names and
positional signatures, computed by libseldq.so (sm_100a)."""
import os as _os
import sys as _sys

_sys.path.append(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from _seldq_pkg import pkg as _pkg  # noqa: E402

import torch  # noqa: E402

_F = _pkg.functional
_I = _pkg.init
_ALG_DQ = _pkg._lib.ALG_DQ


def check_input(input):
    # dual_quaternion_ops.py:14-31
    if input.dim() not in {2, 3, 4, 5}:
        raise RuntimeError("Quaternion linear accepts only input of dimension 2 or 3. Quaternion conv accepts "
                           "up to 5 dim  input.dim = " + str(input.dim()))
    nb_hidden = input.size()[-1] if input.dim() < 4 else input.size()[1]
    if nb_hidden % 4 != 0:
        raise RuntimeError("Quaternion Tensors must be divisible by 4."
                           " input.size()[1] = " + str(nb_hidden))


def _component(input, idx):
    check_input(input)
    axis = input.dim() - 1 if input.dim() < 4 else 1
    n = input.size()[axis] // 4
    return input.narrow(axis, idx * n, n)


def get_r(input):
    return _component(input, 0)


def get_i(input):
    return _component(input, 1)


def get_j(input):
    return _component(input, 2)


def get_k(input):
    return _component(input, 3)


def get_modulus(input, vector_form=False):
    # dual_quaternion_ops.py:88-97
    r, i, j, k = get_r(input), get_i(input), get_j(input), get_k(input)
    sq = r * r + i * i + j * j + k * k
    return torch.sqrt(sq) if vector_form else torch.sqrt(sq.sum(dim=0))


def get_normalized(input, eps=0.0001):
    # dual_quaternion_ops.py:100-107 (2-d and 3-d inputs, as there)
    check_input(input)
    m = get_modulus(input)
    rep = m.repeat(1, 4) if input.dim() == 2 else m.repeat(1, 1, 4)
    return input / (rep.expand_as(input) + eps)


def dual_quaternion_conv(input, r_weight, i_weight, j_weight, k_weight,
                         r_weight_2, i_weight_2, j_weight_2, k_weight_2, bias, stride,
                         padding, groups, dilatation):
    """dual_quaternion_ops.py:111-153 -- | q 0 ; q_e q | block structure fused into the kernels;
    the zero block is never multiplied."""
    if groups != 1:
        raise NotImplementedError("seldq: groups != 1 is not implemented")
    ws = (r_weight, i_weight, j_weight, k_weight, r_weight_2, i_weight_2, j_weight_2, k_weight_2)
    return _F.block_conv(input, ws, bias, stride, padding, dilatation, _ALG_DQ)


def dual_quaternion_linear(input, r_weight, i_weight, j_weight, k_weight,
                           r_weight_2, i_weight_2, j_weight_2, k_weight_2, bias=True):
    """dual_quaternion_ops.py:156-203 (transposed block table)."""
    if bias is True:
        bias = None
    ws = (r_weight, i_weight, j_weight, k_weight, r_weight_2, i_weight_2, j_weight_2, k_weight_2)
    return _F.block_linear(input, ws, bias, _ALG_DQ)


class DualQuaternionLinearFunction(object):
    """dual_quaternion_ops.py:248-372.  Its forward builds the weight matrix of dual_quaternion_linear; its backward takes
    two output gradients for a single output and cannot run in the reference (nothing calls it).  Here: the same forward
    on the same kernels, with dual_quaternion_linear's gradients."""

    @staticmethod
    def apply(input, r_weight, i_weight, j_weight, k_weight, r_weight_2, i_weight_2, j_weight_2, k_weight_2, bias=None):
        check_input(input)
        ws = (r_weight, i_weight, j_weight, k_weight, r_weight_2, i_weight_2, j_weight_2, k_weight_2)
        return _F.block_linear(input, ws, bias, _ALG_DQ)


def _dim1_quaternions(input, what):
    # get_r / get_i / get_j / get_k slice dimension 1 of 2-d and >= 4-d inputs (dual_quaternion_ops.py:34-85); for 3-d
    # inputs they slice the LAST dimension while the results are concatenated along dimension 1
    check_input(input)
    if input.dim() == 3:
        raise NotImplementedError("seldq: %s takes 2-d or >= 4-d inputs (components along dimension 1)" % what)


def q_normalize(input, channel=1):
    """dual_quaternion_ops.py:206-223."""
    _dim1_quaternions(input, "q_normalize")
    if channel != 1:
        raise NotImplementedError("seldq: q_normalize concatenates along dimension 1")
    return _F.q_normalize(input)


def quaternion_exp(input):
    """dual_quaternion_ops.py:227-246."""
    _dim1_quaternions(input, "quaternion_exp")
    return _F.quaternion_exp(input)


def hamilton_product(q0, q1):
    """dual_quaternion_ops.py:374-414."""
    _dim1_quaternions(q0, "hamilton_product")
    _dim1_quaternions(q1, "hamilton_product")
    return _F.hamilton_product(q0, q1)


quaternion_init = _I.dual_quaternion_init
get_kernel_and_weight_shape = _I.get_kernel_and_weight_shape
get_kernel_and_weight_shape_dual = _I.get_kernel_and_weight_shape      # dual_quaternion_ops.py:671-703: the same shapes
create_dropout_mask = _I.create_dropout_mask


def unitary_init(in_features, out_features, rng, kernel_size=None, criterion='he'):
    return _I.unitary_init(in_features, out_features, rng, kernel_size, criterion, dual=True)


def random_init(in_features, out_features, rng, kernel_size=None, criterion='glorot'):
    return _I.random_init(in_features, out_features, rng, kernel_size, criterion, dual=True)


def affect_init(r_weight, i_weight, j_weight, k_weight, r_weight_2, i_weight_2, j_weight_2, k_weight_2,
                init_func, rng, init_criterion):
    _I.affect_init((r_weight, i_weight, j_weight, k_weight), init_func, rng, init_criterion,
                   ws2=(r_weight_2, i_weight_2, j_weight_2, k_weight_2))


def affect_init_conv(r_weight, i_weight, j_weight, k_weight, kernel_size, init_func, rng, init_criterion,
                     r_weight_2=None, i_weight_2=None, j_weight_2=None, k_weight_2=None):
    ws2 = None if r_weight_2 is None else (r_weight_2, i_weight_2, j_weight_2, k_weight_2)
    _I.affect_init_conv((r_weight, i_weight, j_weight, k_weight), kernel_size, init_func, rng, init_criterion,
                        ws2=ws2)
