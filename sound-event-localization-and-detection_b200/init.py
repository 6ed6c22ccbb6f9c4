"""Weight initialisation of the quaternion / dual-quaternion layers.

Host-side numpy, float64 -> parameter dtype, exactly as the reference does it, because identical
seeds must give identical weights (checkpoints and golden fixtures depend on
it).  What has to match is the order and kind of the random draws:

  quaternion flavour   (quaternion_ops.py:596-645): imaginary axis from the GLOBAL numpy RNG
      (three normal(0, s) draws), normalised by |v| + 1e-4; modulus U(-s, s) and phase U(-pi, pi)
      from a private RandomState(123).
  dual-quaternion flavour (dual_quaternion_ops.py:501-552): a private RandomState seeded by
      np.random.randint(1, 1234) FIRST, then modulus chi(4, scale=s) through scipy (global RNG),
      then the axis from three uniform(-1, 1) draws normalised by sqrt(|v|^2 + 1e-4); phase
      U(-pi, pi) from the private RandomState.

The per-weight Python loops of the reference are replaced by the equivalent vectorised
expressions (bit-identical: same IEEE operations per element; checked by tests/test_init_parity.py).
"""
import numpy as np
import torch
from numpy.random import RandomState


def _fans(in_features, out_features, kernel_size):
    rf = 1 if kernel_size is None else int(np.prod(kernel_size))
    return in_features * rf, out_features * rf


def _scale(in_features, out_features, kernel_size, criterion):
    fan_in, fan_out = _fans(in_features, out_features, kernel_size)
    if criterion == "glorot":
        return 1.0 / np.sqrt(2 * (fan_in + fan_out))
    if criterion == "he":
        return 1.0 / np.sqrt(2 * fan_in)
    raise ValueError("Invalid criterion: " + str(criterion))


def _kernel_shape(in_features, out_features, kernel_size):
    if kernel_size is None:
        return (in_features, out_features)
    if type(kernel_size) is int:
        return (out_features, in_features, kernel_size)
    return (out_features, in_features) + tuple(kernel_size)


def _polar(modulus, axis, phase):
    vi, vj, vk = axis
    sp = np.sin(phase)
    return modulus * np.cos(phase), modulus * vi * sp, modulus * vj * sp, modulus * vk * sp


def quaternion_init(in_features, out_features, rng, kernel_size=None, criterion="glorot"):
    """Quaternion-layer flavour (quaternion_ops.py:596-645).  `rng` is accepted and ignored, as in
    the reference (it builds RandomState(123) itself)."""
    s = _scale(in_features, out_features, kernel_size, criterion)
    private = RandomState(123)
    shape = _kernel_shape(in_features, out_features, kernel_size)
    n = int(np.prod(shape))
    vi = np.random.normal(0.0, s, n)
    vj = np.random.normal(0.0, s, n)
    vk = np.random.normal(0.0, s, n)
    norm = np.sqrt(vi ** 2 + vj ** 2 + vk ** 2) + 0.0001
    axis = [(v / norm).reshape(shape) for v in (vi, vj, vk)]
    modulus = private.uniform(low=-s, high=s, size=shape)
    phase = private.uniform(low=-np.pi, high=np.pi, size=shape)
    return _polar(modulus, axis, phase)


def dual_quaternion_init(in_features, out_features, rng, kernel_size=None, criterion="glorot"):
    """Dual-quaternion-layer flavour (dual_quaternion_ops.py:501-552)."""
    from scipy.stats import chi
    s = _scale(in_features, out_features, kernel_size, criterion)
    private = RandomState(np.random.randint(1, 1234))
    shape = _kernel_shape(in_features, out_features, kernel_size)
    modulus = chi.rvs(4, loc=0, scale=s, size=shape)
    n = int(np.prod(shape))
    vi = np.random.uniform(-1.0, 1.0, n)
    vj = np.random.uniform(-1.0, 1.0, n)
    vk = np.random.uniform(-1.0, 1.0, n)
    norm = np.sqrt(vi ** 2 + vj ** 2 + vk ** 2 + 0.0001)
    axis = [(v / norm).reshape(shape) for v in (vi, vj, vk)]
    phase = private.uniform(low=-np.pi, high=np.pi, size=shape)
    return _polar(modulus, axis, phase)


def unitary_init(in_features, out_features, rng, kernel_size=None, criterion="he", dual=False):
    """Unit quaternions from the GLOBAL numpy RNG: four normal(0, s) draws in the quaternion file
    (quaternion_ops.py:509-550), four uniform(-1, 1) draws in the dual file
    (dual_quaternion_ops.py:417-452); both divide by |q| + 1e-4."""
    shape = _kernel_shape(in_features, out_features, kernel_size)
    n = int(np.prod(shape))
    if dual:
        vr, vi, vj, vk = (np.random.uniform(-1.0, 1.0, n) for _ in range(4))
    else:
        s = _scale(in_features, out_features, kernel_size, criterion)
        vr, vi, vj, vk = (np.random.normal(0.0, s, n) for _ in range(4))
    norm = np.sqrt(vr ** 2 + vi ** 2 + vj ** 2 + vk ** 2) + 0.0001
    return tuple((v / norm).reshape(shape) for v in (vr, vi, vj, vk))


def random_init(in_features, out_features, rng, kernel_size=None, criterion="glorot", dual=False):
    """Four uniform draws from the global RNG: U(0,1)*s in the quaternion file
    (quaternion_ops.py:553-593), plain U(-1,1) in the dual file (dual_quaternion_ops.py:455-498)."""
    s = _scale(in_features, out_features, kernel_size, criterion)
    shape = _kernel_shape(in_features, out_features, kernel_size)
    n = int(np.prod(shape))
    if dual:
        return tuple(np.random.uniform(-1.0, 1.0, n).reshape(shape) for _ in range(4))
    return tuple(np.random.uniform(0.0, 1.0, n).reshape(shape) * s for _ in range(4))


def _assign(params, values):
    for p, v in zip(params, values):
        p.data = torch.from_numpy(np.asarray(v)).type_as(p.data)


def _check_same_size(ws):
    if any(w.size() != ws[0].size() for w in ws[1:]):
        raise ValueError("The real and imaginary weights should have the same size. Found: "
                         + " ".join("%s:%s" % (n, tuple(w.size())) for n, w in zip("rijk", ws)))


def affect_init(ws, init_func, rng, criterion, ws2=None):
    """Linear layers: weights are (in, out) matrices (quaternion_ops.py:656-674,
    dual_quaternion_ops.py:564-592)."""
    _check_same_size(ws)
    if ws[0].dim() != 2:
        raise Exception("affect_init accepts only matrices. Found dimension = " + str(ws[0].dim()))
    _assign(ws, init_func(ws[0].size(0), ws[0].size(1), rng, None, criterion))
    if ws2 is not None:
        _assign(ws2, init_func(ws2[0].size(0), ws2[0].size(1), rng, None, criterion))


def affect_init_conv(ws, kernel_size, init_func, rng, criterion, ws2=None):
    """Convolutions: weights are (out, in, *k) (quaternion_ops.py:677-703, dual_quaternion_ops.py:595-636)."""
    _check_same_size(ws)
    if ws[0].dim() <= 2:
        raise Exception("affect_conv_init accepts only tensors that have more than 2 dimensions. "
                        "Found dimension = " + str(ws[0].dim()))
    _assign(ws, init_func(ws[0].size(1), ws[0].size(0), rng=rng, kernel_size=kernel_size, criterion=criterion))
    if ws2 is not None:
        _assign(ws2, init_func(ws2[0].size(1), ws2[0].size(0), rng=rng, kernel_size=kernel_size,
                               criterion=criterion))


def get_kernel_and_weight_shape(operation, in_channels, out_channels, kernel_size):
    """(kernel_size, weight_shape) for a compact tensor (quaternion_ops.py:706-735)."""
    dims = {"convolution1d": 1, "convolution2d": 2, "convolution3d": 3}.get(operation, 2)
    if dims == 1:
        if type(kernel_size) is not int:
            raise ValueError("An invalid kernel_size was supplied for a 1d convolution. The kernel size "
                             "must be integer in the case. Found kernel_size = " + str(kernel_size))
        return kernel_size, (out_channels, in_channels, kernel_size)
    if type(kernel_size) is int:
        ks = (kernel_size,) * dims
    else:
        if len(kernel_size) != dims:
            raise ValueError("An invalid kernel_size was supplied for a %dd convolution. The kernel size must be "
                             "either an integer or a tuple of %d. Found kernel_size = %s"
                             % (dims, dims, str(kernel_size)))
        ks = kernel_size
    return ks, (out_channels, in_channels) + tuple(ks)


def create_dropout_mask(dropout_p, size, rng, as_type, operation='linear'):
    """quaternion_ops.py:648-653 / dual_quaternion_ops.py:555-561: a Bernoulli(1 - p) keep mask drawn from the numpy
    RandomState `rng`, as a tensor of type `as_type`."""
    if operation != 'linear':
        raise Exception("create_dropout_mask accepts only 'linear'. Found operation = " + str(operation))
    import torch
    return torch.from_numpy(rng.binomial(n=1, p=1 - dropout_p, size=size)).type(as_type)
