"""ORACLE -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy, float64 by default) of the reference's quaternion / dual-quaternion
hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import this module; the product package never does.

Parity pin: the reference ships no tests or golden vectors for this path, so the
pin is (1) live calls into the imported reference in the build container
(tests/test_oracle_vs_reference.py) and (2) the committed fixtures under tests/golden/ minted
from the reference by oracle/make_golden.py.

What is restated, and where it lives in the reference:
  * Hamilton sign / component tables ............ quaternion/quaternion_ops.py:131-135
  * dual-quaternion block matrix [[Q,0],[Q2,Q]] .. dual_quaternion/dual_quaternion_ops.py:122-140
  * quaternion_conv / dual_quaternion_conv ...... quaternion_ops.py:125-147, dual_quaternion_ops.py:111-153
  * quaternion_linear ........................... quaternion_ops.py:299-327 (and :392-464)
  * dual_quaternion_linear (transposed table) ... dual_quaternion_ops.py:156-203
  * spectrum_fast (scipy.signal.stft defaults) .. utility_functions.py:129-155
  * rotation variants (conv / transposed / linear) quaternion_ops.py:174-295, :330-388
  * hamilton_product, q_normalize, quaternion_exp  quaternion_ops.py:467-507, dual_quaternion_ops.py:206-246, :374-414
The third-party arithmetic underneath (torch F.conv*, torch.mm, scipy.signal.stft 1.18) is
restated from its published definition: cross-correlation with zero padding, and
Z = rfft(frame * w) / sum(w) with boundary='zeros', padded=True, periodic Hamming.
"""
import numpy as np

# sign[a][b], comp[a][b] = a ^ b : expanded Q weight Wq[a*O+o, b*I+i] = sign[a][b] * W_{a^b}[o, i]
HAMILTON_SIGN = np.array([[+1, -1, -1, -1],
                          [+1, +1, -1, +1],
                          [+1, +1, +1, -1],
                          [+1, -1, +1, +1]], dtype=np.int8)


def block_table(algebra):
    """(widx, sign) with shape (ncomp_out, ncomp_in): which compact weight (index into
    [r,i,j,k,r2,i2,j2,k2]) feeds output component a from input component b, and its sign;
    widx = -1 marks a structural zero block.

    'R'        : 1x1 identity
    'Q'        : quaternion conv and quaternion linear (quaternion_ops.py:131-135, :310-314)
    'DQ'       : dual-quaternion conv (dual_quaternion_ops.py:122-140)
    'DQ_LINEAR': dual_quaternion_linear builds the same block matrix but multiplies from the
                 right on (in,out) weights (dual_quaternion_ops.py:170-197), i.e. the transpose.
    """
    if algebra == "R":
        return np.zeros((1, 1), np.int8), np.ones((1, 1), np.int8)
    if algebra == "Q":
        a, b = np.meshgrid(np.arange(4), np.arange(4), indexing="ij")
        return (a ^ b).astype(np.int8), HAMILTON_SIGN.copy()
    if algebra in ("DQ", "DQ_LINEAR"):
        widx = -np.ones((8, 8), np.int8)
        sign = np.zeros((8, 8), np.int8)
        for a in range(8):
            for b in range(8):
                ha, ca, hb, cb = a // 4, a % 4, b // 4, b % 4
                if algebra == "DQ":
                    if ha == hb:
                        widx[a, b], sign[a, b] = ca ^ cb, HAMILTON_SIGN[ca, cb]
                    elif ha == 1 and hb == 0:
                        widx[a, b], sign[a, b] = 4 + (ca ^ cb), HAMILTON_SIGN[ca, cb]
                else:
                    if ha == hb:
                        widx[a, b], sign[a, b] = ca ^ cb, HAMILTON_SIGN[cb, ca]
                    elif ha == 0 and hb == 1:
                        widx[a, b], sign[a, b] = 4 + (ca ^ cb), HAMILTON_SIGN[cb, ca]
        return widx, sign
    raise ValueError(algebra)


def expand_weight(weights, algebra, linear=False):
    """Dense real weight the reference builds with torch.cat every call.
    conv: compact (O, I, *k) -> (ncomp*O, ncomp*I, *k);  linear: compact (I, O) -> (ncomp*I, ncomp*O)."""
    widx, sign = block_table(algebra)
    nc = widx.shape[0]
    w0 = np.asarray(weights[0])
    if linear:
        I, O = w0.shape
        W = np.zeros((nc * I, nc * O), w0.dtype)
        for a in range(nc):
            for b in range(nc):
                if widx[a, b] >= 0:
                    W[b * I:(b + 1) * I, a * O:(a + 1) * O] = sign[a, b] * np.asarray(weights[widx[a, b]])
        return W
    O, I = w0.shape[:2]
    W = np.zeros((nc * O, nc * I) + w0.shape[2:], w0.dtype)
    for a in range(nc):
        for b in range(nc):
            if widx[a, b] >= 0:
                W[a * O:(a + 1) * O, b * I:(b + 1) * I] = sign[a, b] * np.asarray(weights[widx[a, b]])
    return W


def _pair(v):
    return (v, v) if np.isscalar(v) else tuple(v)


def conv_nd(x, W, bias=None, stride=1, padding=0, dilation=1):
    """Cross-correlation as torch.nn.functional.conv1d/conv2d define it (groups=1).
    x (N,C,L) or (N,C,H,W); W (Co,Ci,k) or (Co,Ci,kh,kw)."""
    x = np.asarray(x)
    one_d = x.ndim == 3
    if one_d:
        x = x[:, :, None, :]
        W = W[:, :, None, :]
        stride, padding, dilation = (1, stride), (0, padding), (1, dilation)
    sh, sw = _pair(stride)
    ph, pw = _pair(padding)
    dh, dw = _pair(dilation)
    N, C, H, Wd = x.shape
    Co, Ci, kh, kw = W.shape
    assert Ci == C
    OH = (H + 2 * ph - dh * (kh - 1) - 1) // sh + 1
    OW = (Wd + 2 * pw - dw * (kw - 1) - 1) // sw + 1
    xp = np.zeros((N, C, H + 2 * ph, Wd + 2 * pw), x.dtype)
    xp[:, :, ph:ph + H, pw:pw + Wd] = x
    y = np.zeros((N, Co, OH, OW), np.result_type(x.dtype, W.dtype))
    for i in range(kh):
        for j in range(kw):
            patch = xp[:, :, i * dh:i * dh + (OH - 1) * sh + 1:sh, j * dw:j * dw + (OW - 1) * sw + 1:sw]
            y += np.einsum("oc,nchw->nohw", W[:, :, i, j], patch, optimize=True)
    if bias is not None:
        y += np.asarray(bias)[None, :, None, None]
    return y[:, :, 0, :] if one_d else y


def conv_nd_backward(x, W, gy, stride=1, padding=0, dilation=1):
    """(gx, gW, gb) of conv_nd w.r.t. dense W -- adjoint of the loop above."""
    x = np.asarray(x)
    one_d = x.ndim == 3
    if one_d:
        x, W, gy = x[:, :, None, :], W[:, :, None, :], gy[:, :, None, :]
        stride, padding, dilation = (1, stride), (0, padding), (1, dilation)
    sh, sw = _pair(stride)
    ph, pw = _pair(padding)
    dh, dw = _pair(dilation)
    N, C, H, Wd = x.shape
    Co, Ci, kh, kw = W.shape
    OH, OW = gy.shape[2:]
    xp = np.zeros((N, C, H + 2 * ph, Wd + 2 * pw), x.dtype)
    xp[:, :, ph:ph + H, pw:pw + Wd] = x
    gxp = np.zeros_like(xp)
    gW = np.zeros_like(W)
    for i in range(kh):
        for j in range(kw):
            sl = (slice(None), slice(None), slice(i * dh, i * dh + (OH - 1) * sh + 1, sh),
                  slice(j * dw, j * dw + (OW - 1) * sw + 1, sw))
            gW[:, :, i, j] = np.einsum("nohw,nchw->oc", gy, xp[sl], optimize=True)
            gxp[sl] += np.einsum("oc,nohw->nchw", W[:, :, i, j], gy, optimize=True)
    gx = gxp[:, :, ph:ph + H, pw:pw + Wd]
    gb = gy.sum(axis=(0, 2, 3))
    if one_d:
        return gx[:, :, 0, :], gW[:, :, 0, :], gb
    return gx, gW, gb


def compact_grads(gW_dense, algebra, O, I, linear=False):
    """Adjoint of expand_weight: scatter-sum the dense gradient back onto the compact tensors
    (what autograd does through torch.cat / neg in the reference)."""
    widx, sign = block_table(algebra)
    nc = widx.shape[0]
    nw = int(widx.max()) + 1
    out = [None] * nw
    for a in range(nc):
        for b in range(nc):
            e = widx[a, b]
            if e < 0:
                continue
            blk = (gW_dense[b * I:(b + 1) * I, a * O:(a + 1) * O] if linear
                   else gW_dense[a * O:(a + 1) * O, b * I:(b + 1) * I])
            out[e] = sign[a, b] * blk if out[e] is None else out[e] + sign[a, b] * blk
    return out


def qconv(x, weights, bias=None, stride=1, padding=0, dilation=1, algebra="Q"):
    """quaternion_conv / dual_quaternion_conv (quaternion_ops.py:125-147; dual_quaternion_ops.py:111-153)."""
    return conv_nd(x, expand_weight(weights, algebra), bias, stride, padding, dilation)


def qconv_backward(x, weights, gy, stride=1, padding=0, dilation=1, algebra="Q"):
    W = expand_weight(weights, algebra)
    gx, gW, gb = conv_nd_backward(x, W, gy, stride, padding, dilation)
    O, I = np.asarray(weights[0]).shape[:2]
    return gx, compact_grads(gW, algebra, O, I), gb


def _transpose_out_shape(in_sp, ks, stride, padding, dilation, output_padding):
    nd = len(in_sp)
    s, p, d, op = (v if nd == 2 else (v,) for v in (_pair(stride) if nd == 2 else stride, _pair(padding) if nd == 2 else padding,
                                                    _pair(dilation) if nd == 2 else dilation,
                                                    _pair(output_padding) if nd == 2 else output_padding))
    return tuple((in_sp[a] - 1) * s[a] - 2 * p[a] + d[a] * (ks[a] - 1) + op[a] + 1 for a in range(nd))


def conv_transpose_nd(x, W, bias=None, stride=1, padding=0, dilation=1, output_padding=0):
    """F.conv_transpose1d / 2d of a dense (in, out, k...) weight: the input gradient of the convolution with that weight
    read as (out', in') -- stride, padding and dilation the same -- evaluated at x; output_padding only picks the output
    size among those the strided convolution maps onto x's."""
    x = np.asarray(x, np.float64)
    nd = x.ndim - 2
    out_sp = _transpose_out_shape(x.shape[2:], W.shape[2:], stride, padding, dilation, output_padding)
    gx, _, _ = conv_nd_backward(np.zeros((x.shape[0], W.shape[1]) + out_sp), np.asarray(W, np.float64), x, stride, padding,
                                dilation)
    return gx if bias is None else gx + np.asarray(bias, np.float64).reshape((1, -1) + (1,) * nd)


def qconv_transpose(x, weights, bias=None, padding=0, dilation=1, stride=1, output_padding=0):
    """quaternion_transpose_conv (quaternion_ops.py:149-172): F.conv_transpose of the expanded (in, out, k...) weight,
    whose blocks follow the convolution's table."""
    return conv_transpose_nd(x, expand_weight(weights, "Q"), bias, stride, padding, dilation, output_padding)


def qlinear(x, weights, bias=None, algebra="Q"):
    """quaternion_linear (quaternion_ops.py:299-327) for algebra='Q';
    dual_quaternion_linear (dual_quaternion_ops.py:156-203) for algebra='DQ_LINEAR'."""
    y = np.asarray(x) @ expand_weight(weights, algebra, linear=True)
    return y if bias is None else y + np.asarray(bias)


def qlinear_backward(x, weights, gy, algebra="Q"):
    W = expand_weight(weights, algebra, linear=True)
    x2, g2 = np.asarray(x).reshape(-1, W.shape[0]), np.asarray(gy).reshape(-1, W.shape[1])
    gx = (g2 @ W.T).reshape(np.asarray(x).shape)
    gW = x2.T @ g2
    I, O = np.asarray(weights[0]).shape
    return gx, compact_grads(gW, algebra, O, I, linear=True), g2.sum(0)


# ---- rotation variants and point-wise quaternion operators ------------------------------------------
def rotation_weight(weights, quaternion_format=False, dtype=np.float64):
    """The real weight of quaternion_conv_rotation / quaternion_transpose_conv_rotation / quaternion_linear_rotation
    (quaternion_ops.py:188-220, :249-281, :344-376): compact (d0, d1, k...) -> (nc d0, nc d1, k...), nc = 4 with
    quaternion_format (zero block row / column first), else 3.  Restated as written: norm_factor = 2 |q| (a product,
    not the 2 / |q|^2 of the textbook formula), products as (f * a) * b, sums left to right -- in float32 every
    operation rounds as torch's does, so dtype=np.float32 reproduces the reference's weight bit for bit."""
    r, i, j, k = (np.asarray(w, dtype) for w in weights)
    one, two = dtype(1.0), dtype(2.0)
    norm = np.sqrt(r * r + i * i + j * j + k * k)
    f = two * norm
    si, sj, sk = f * (i * i), f * (j * j), f * (k * k)
    ri, rj, rk = f * r * i, f * r * j, f * r * k
    ij, ik, jk = f * i * j, f * i * k, f * j * k
    cols = [[one - (sj + sk), ij - rk, ik + rj],
            [ij + rk, one - (si + sk), jk - ri],
            [ik - rj, jk + ri, one - (si + sj)]]
    if quaternion_format:
        z = np.zeros_like(r)
        cols = [[z, z, z, z]] + [[z] + c for c in cols]
    return np.concatenate([np.concatenate(c, axis=0) for c in cols], axis=1)


def rotation_weight_backward(weights, gW, quaternion_format=False):
    """Gradients of (r, i, j, k) given the gradient of rotation_weight's output: entry = delta_ab + f P_ab(q), f = 2 |q|,
    so d entry / d q_c = (2 q_c / |q|) P_ab + f dP_ab / dq_c."""
    q = [np.asarray(w, np.float64) for w in weights]
    r, i, j, k = q
    d0, d1 = r.shape[:2]
    z = 1 if quaternion_format else 0
    G = [[np.asarray(gW, np.float64)[(a + z) * d0:(a + z + 1) * d0, (b + z) * d1:(b + z + 1) * d1] for b in range(3)]
         for a in range(3)]
    n = np.sqrt(r * r + i * i + j * j + k * k)
    f = 2.0 * n
    P = [[-(j * j + k * k), i * j + r * k, i * k - r * j],
         [i * j - r * k, -(i * i + k * k), j * k + r * i],
         [i * k + r * j, j * k - r * i, -(i * i + j * j)]]
    S = sum(G[a][b] * P[a][b] for a in range(3) for b in range(3))
    zero = np.zeros_like(r)
    # dP[c][a][b] = d P_ab / d q_c
    dP = [[[zero, k, -j], [-k, zero, i], [j, -i, zero]],
          [[zero, j, k], [j, -2 * i, r], [k, -r, -2 * i]],
          [[-2 * j, i, -r], [i, zero, k], [r, k, -2 * j]],
          [[-2 * k, r, i], [-r, -2 * k, j], [i, j, zero]]]
    return [2.0 * q[c] / n * S + f * sum(G[a][b] * dP[c][a][b] for a in range(3) for b in range(3)) for c in range(4)]


def qconv_rotation(x, weights, bias=None, stride=1, padding=0, dilation=1, quaternion_format=False):
    """quaternion_conv_rotation (quaternion_ops.py:174-232)."""
    return conv_nd(x, rotation_weight(weights, quaternion_format), bias, stride, padding, dilation)


def qconv_rotation_backward(x, weights, gy, stride=1, padding=0, dilation=1, quaternion_format=False):
    gx, gW, gb = conv_nd_backward(x, rotation_weight(weights, quaternion_format), gy, stride, padding, dilation)
    return gx, rotation_weight_backward(weights, gW, quaternion_format), gb


def qconv_transpose_rotation(x, weights, bias=None, padding=0, dilation=1, quaternion_format=False, stride=1,
                             output_padding=0):
    """quaternion_transpose_conv_rotation (quaternion_ops.py:235-295): F.conv_transpose of the (in, out, k...) rotation
    weight."""
    return conv_transpose_nd(x, rotation_weight(weights, quaternion_format), bias, stride, padding, dilation, output_padding)


def qlinear_rotation(x, weights, bias=None, quaternion_format=False):
    """quaternion_linear_rotation (quaternion_ops.py:330-388)."""
    y = np.asarray(x, np.float64) @ rotation_weight(weights, quaternion_format)
    return y if bias is None else y + np.asarray(bias, np.float64)


def _components(x):
    """get_r / get_i / get_j / get_k of dual_quaternion_ops.py:34-85 for 2-d and >= 4-d inputs: quarters of dim 1."""
    x = np.asarray(x)
    n = x.shape[1] // 4
    return [x[:, c * n:(c + 1) * n] for c in range(4)]


def hamilton_product(q0, q1, dtype=np.float64):
    """quaternion_ops.py:467-507 / dual_quaternion_ops.py:374-414."""
    a, b = _components(np.asarray(q0, dtype)), _components(np.asarray(q1, dtype))
    r = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3]
    i = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2]
    j = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1]
    k = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]
    return np.concatenate([r, i, j, k], axis=1)


def q_normalize(x, dtype=np.float64):
    """dual_quaternion_ops.py:206-223 (channel = 1)."""
    r, i, j, k = _components(np.asarray(x, dtype))
    norm = np.sqrt(r * r + i * i + j * j + k * k + dtype(0.0001))
    return np.concatenate([r / norm, i / norm, j / norm, k / norm], axis=1)


def quaternion_exp(x, dtype=np.float64):
    """dual_quaternion_ops.py:227-246."""
    r, i, j, k = _components(np.asarray(x, dtype))
    nv = np.sqrt(i * i + j * j + k * k) + dtype(0.0001)
    e = np.exp(r)
    return np.concatenate([e * np.cos(nv), e * ((i / nv) * np.sin(nv)), e * ((j / nv) * np.sin(nv)),
                           e * ((k / nv) * np.sin(nv))], axis=1)


def hamming_periodic(n):
    """scipy.signal.get_window('hamming', n) (fftbins=True)."""
    return 0.54 - 0.46 * np.cos(2.0 * np.pi * np.arange(n) / n)


def spectrum_fast(x, nperseg=512, noverlap=128, window="hamming", cut_dc=True,
                  output_phase=True, cut_last_timeframe=True):
    """utility_functions.py:129-155 on top of scipy.signal.stft defaults (boundary='zeros',
    padded=True, detrend=False, return_onesided=True, scaling='spectrum')."""
    if window != "hamming":
        raise NotImplementedError("only the Hamming window of the reference path is restated")
    x = np.asarray(x)
    if x.ndim != 2:
        raise ValueError("spectrum_fast expects (channels, samples)")
    w = hamming_periodic(nperseg)
    hop = nperseg - noverlap
    xp = np.concatenate([np.zeros((x.shape[0], nperseg // 2), x.dtype), x,
                         np.zeros((x.shape[0], nperseg // 2), x.dtype)], axis=1)
    n = xp.shape[1]
    nadd = (-(n - nperseg) % hop) % nperseg
    xp = np.concatenate([xp, np.zeros((x.shape[0], nadd), x.dtype)], axis=1)
    nfr = (xp.shape[1] - nperseg) // hop + 1
    idx = np.arange(nperseg)[None, :] + hop * np.arange(nfr)[:, None]
    frames = xp[:, idx] * w                                   # (C, nfr, nperseg)
    Z = np.fft.rfft(frames, axis=-1) / w.sum()                # (C, nfr, nperseg//2+1)
    Z = np.swapaxes(Z, 1, 2)                                  # (C, F, T)
    out = np.abs(Z)
    if output_phase:
        out = np.concatenate((out, np.angle(Z)), axis=-3)
    if cut_dc:
        out = out[:, 1:, :]
    if cut_last_timeframe:
        out = out[:, :, :-1]
    return out


def rel_err(a, b):
    """max-abs-normalised error used by every parity test."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
