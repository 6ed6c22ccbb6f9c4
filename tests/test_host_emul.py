"""Runs the SIMT kernels' __host__ __device__ phase functions on the CPU (tests/host_emul/emul.cpp)
with the library's launch geometry and checks them against the golden fixtures.  This validates
the index arithmetic of csrc/conv_simt.cuh and csrc/stft.cuh without a GPU; the GPU tests then
check the real launches."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden
from oracle import algebra as A

HERE = os.path.join(ROOT, "tests", "host_emul")
NW = {"Q": 4, "DQ": 8}
ALG = {"R": 0, "Q": 1, "DQ": 2}


class ConvDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "algebra", "precision", "ndim", "batch", "cin", "cout", "in_h", "in_w", "k_h", "k_w",
        "stride_h", "stride_w", "pad_h", "pad_w", "dil_h", "dil_w")]


class LinearDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("algebra", "precision", "rows", "in_features", "out_features")]


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(HERE, "libseldq_emul.so")
    src = os.path.join(HERE, "emul.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I/usr/local/cuda/include", src, "-o", so])
    lib = ctypes.CDLL(so)
    lib.emul_last_error.restype = ctypes.c_char_p
    return lib


def fptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def ptr_array(arrs):
    return (ctypes.POINTER(ctypes.c_float) * len(arrs))(*[fptr(a) for a in arrs])


def conv_desc(meta, d):
    x = d["x"]
    w = d["w0"]
    nc = NW[meta["algebra"]]
    one_d = meta["ndim"] == 1
    s, p, dl = meta["stride"], meta["padding"], meta["dilation"]
    return ConvDesc(ALG[meta["algebra"]], 0, meta["ndim"], x.shape[0], x.shape[1], w.shape[0] * nc,
                    1 if one_d else x.shape[2], x.shape[-1], 1 if one_d else w.shape[2], w.shape[-1],
                    1 if one_d else s, s, 0 if one_d else p, p, 1 if one_d else dl, dl)


@pytest.mark.parametrize("name", golden_names("conv"))
def test_emulated_conv_kernels_match_golden(emul, name):
    meta, d = load_golden(name)
    nw = NW[meta["algebra"]]
    desc = conv_desc(meta, d)
    x = np.ascontiguousarray(d["x"], np.float32)
    gy = np.ascontiguousarray(d["gy"], np.float32)
    ws = [np.ascontiguousarray(d["w%d" % i], np.float32) for i in range(nw)]
    b = np.ascontiguousarray(d["b"], np.float32) if meta["bias"] else None
    y = np.full(d["y"].shape, np.nan, np.float32)
    assert emul.emul_conv(ctypes.byref(desc), 0, fptr(x), ptr_array(ws), fptr(b) if b is not None else None, fptr(y)) == 0
    assert A.rel_err(y, d["y"]) < 1e-5
    gx = np.full(x.shape, np.nan, np.float32)
    assert emul.emul_conv(ctypes.byref(desc), 1, fptr(gy), ptr_array(ws), None, fptr(gx)) == 0
    assert A.rel_err(gx, d["gx"]) < 1e-5
    gws = [np.zeros_like(w) for w in ws]
    assert emul.emul_conv_wgrad(ctypes.byref(desc), fptr(x), fptr(gy), ptr_array(gws), 3) == 0
    for i in range(nw):
        assert A.rel_err(gws[i], d["gw%d" % i]) < 1e-5, i


@pytest.mark.parametrize("name", golden_names("linear"))
def test_emulated_linear_kernels_match_golden(emul, name):
    meta, d = load_golden(name)
    nw = NW[meta["algebra"]]
    x = np.ascontiguousarray(d["x"], np.float32)
    gy = np.ascontiguousarray(d["gy"], np.float32)
    ws = [np.ascontiguousarray(d["w%d" % i], np.float32) for i in range(nw)]
    b = np.ascontiguousarray(d["b"], np.float32)
    desc = LinearDesc(ALG[meta["algebra"]], 0, x.shape[0], x.shape[1], gy.shape[1])
    y = np.full(gy.shape, np.nan, np.float32)
    assert emul.emul_linear(ctypes.byref(desc), 0, fptr(x), ptr_array(ws), fptr(b), fptr(y)) == 0
    assert A.rel_err(y, d["y"]) < 1e-5
    gx = np.full(x.shape, np.nan, np.float32)
    assert emul.emul_linear(ctypes.byref(desc), 1, fptr(gy), ptr_array(ws), None, fptr(gx)) == 0
    assert A.rel_err(gx, d["gx"]) < 1e-5
    gws = [np.zeros_like(w) for w in ws]
    assert emul.emul_linear_wgrad(ctypes.byref(desc), fptr(x), fptr(gy), ptr_array(gws), 2) == 0
    for i in range(nw):
        assert A.rel_err(gws[i], d["gw%d" % i]) < 1e-5, i


@pytest.mark.parametrize("name", golden_names("stft"))
def test_emulated_stft_kernel_matches_golden(emul, name):
    meta, d = load_golden(name)
    x = np.ascontiguousarray(d["x"], np.float32)
    out = np.full(d["out"].shape, np.nan, np.float32)
    C = x.shape[0]
    rc = emul.emul_stft(fptr(x), 1, C, ctypes.c_longlong(x.shape[1]), meta["nperseg"], meta["noverlap"], 1,
                        int(meta["output_phase"]), 1, fptr(out))
    assert rc == 0, emul.emul_last_error()
    assert A.rel_err(out[:C], d["out"][:C]) < 1e-5
    if meta["output_phase"]:
        mag = d["out"][:C]
        dphi = np.abs(np.angle(np.exp(1j * (out[C:].astype(np.float64) - d["out"][C:]))))
        # phase of a bin is only defined to ~eps_fp32 * max|Z| / |Z|
        assert (dphi * mag / mag.max()).max() < 1e-5


@pytest.mark.parametrize("pairs", [16, 28])
@pytest.mark.parametrize("name", golden_names("stft"))
def test_emulated_frame_pair_stft_kernel_matches_golden(emul, name, pairs):
    """csrc/stft_pair.cuh (two consecutive frames per packed register, 16 threads per frame pair): the magnitude plane
    of every STFT fixture, with a batch size that divides the frame count (16 pairs) and one that does not (28), and
    the phase plane (packed polynomial atan2, 12 pairs per batch) where the fixture has one."""
    meta, d = load_golden(name)
    x = np.ascontiguousarray(d["x"], np.float32)
    C = x.shape[0]
    out = np.full((C,) + d["out"].shape[1:], np.nan, np.float32)
    rc = emul.emul_stft_pairs(fptr(x), 1, C, ctypes.c_longlong(x.shape[1]), meta["nperseg"], meta["noverlap"], 1, 0, 1,
                              pairs, fptr(out))
    assert rc == 0, emul.emul_last_error()
    assert A.rel_err(out, d["out"][:C]) < 1e-5
    if meta["output_phase"] and pairs == 16:
        out = np.full(d["out"].shape, np.nan, np.float32)
        rc = emul.emul_stft_pairs(fptr(x), 1, C, ctypes.c_longlong(x.shape[1]), meta["nperseg"], meta["noverlap"], 1, 1,
                                  1, 12, fptr(out))
        assert rc == 0, emul.emul_last_error()
        assert A.rel_err(out[:C], d["out"][:C]) < 1e-5
        mag = d["out"][:C]
        dphi = np.abs(np.angle(np.exp(1j * (out[C:].astype(np.float64) - d["out"][C:]))))
        assert (dphi * mag / mag.max()).max() < 1e-5


@pytest.mark.parametrize("n_samples", [700, 5000, 12345, 400 * 7 + 13])
def test_emulated_frame_pair_stft_ragged_lengths_against_oracle(emul, n_samples):
    """Frame-pair kernel on lengths that leave odd frame counts, partial frame pairs and partial batches, signals that
    start / end inside a frame (zero boundary extension), one and three channels: magnitude (16 and 28 pairs per batch)
    and magnitude + phase (12 pairs) against oracle.spectrum_fast."""
    rng = np.random.default_rng(n_samples)
    for C in (1, 3):
        x = (0.3 * rng.standard_normal((C, n_samples))).astype(np.float32)
        for phase in (0, 1):
            ref = A.spectrum_fast(x.astype(np.float64), nperseg=512, noverlap=112, output_phase=bool(phase))
            for pairs in ((12,) if phase else (16, 28)):
                out = np.full(ref.shape, np.nan, np.float32)
                rc = emul.emul_stft_pairs(fptr(x), 1, C, ctypes.c_longlong(n_samples), 512, 112, 1, phase, 1, pairs, fptr(out))
                assert rc == 0, emul.emul_last_error()
                assert np.isfinite(out).all()
                assert A.rel_err(out[:C], ref[:C]) < 1e-5
                if phase:
                    mag = ref[:C]
                    dphi = np.abs(np.angle(np.exp(1j * (out[C:].astype(np.float64) - ref[C:]))))
                    assert (dphi * mag / mag.max()).max() < 1e-5


def test_dq_linear_runs_as_a_1x1_convolution_with_its_own_block_table(emul):
    """functional._linear_as_conv: dual_quaternion_linear (dual_quaternion_ops.py:156-203) on (rows, in) equals a
    1x1 convolution over the transposed matrices with the block table SELDQ_ALG_DQ_LINEAR (= 3) and transposed
    compact weights; checked here through the emulated FFMA kernels, which read the same table."""
    meta, d = load_golden("linear_dq_c48")
    x = np.ascontiguousarray(d["x"], np.float32)                      # (rows, in)
    gy = np.ascontiguousarray(d["gy"], np.float32)                    # (rows, out)
    rows, fin = x.shape
    fout = gy.shape[1]
    ws = [np.ascontiguousarray(d["w%d" % i].T[:, :, None], np.float32) for i in range(8)]     # (out/8, in/8, 1)
    desc = ConvDesc(3, 0, 1, 1, fin, fout, 1, rows, 1, 1, 1, 1, 0, 0, 1, 1)
    xc = np.ascontiguousarray(x.T[None], np.float32)                  # (1, in, rows)
    y = np.full((1, fout, rows), np.nan, np.float32)
    b = np.ascontiguousarray(d["b"], np.float32)
    assert emul.emul_conv(ctypes.byref(desc), 0, fptr(xc), ptr_array(ws), fptr(b), fptr(y)) == 0
    assert A.rel_err(y[0].T, d["y"]) < 1e-5
    gyc = np.ascontiguousarray(gy.T[None], np.float32)
    gx = np.full(xc.shape, np.nan, np.float32)
    assert emul.emul_conv(ctypes.byref(desc), 1, fptr(gyc), ptr_array(ws), None, fptr(gx)) == 0
    assert A.rel_err(gx[0].T, d["gx"]) < 1e-5
    gws = [np.zeros_like(w) for w in ws]
    assert emul.emul_conv_wgrad(ctypes.byref(desc), fptr(xc), fptr(gyc), ptr_array(gws), 2) == 0
    for i in range(8):
        assert A.rel_err(gws[i][:, :, 0].T, d["gw%d" % i]) < 1e-5, i


@pytest.mark.parametrize("name,alg", [("linear_dq_c48", 5), ("linear_dq", 5), ("linear_q", 4)])
def test_linear_layers_run_as_1x1_convolutions_on_their_own_tensors(emul, name, alg):
    """SELDQ_ALG_Q_LINEAR_IO (4) / SELDQ_ALG_DQ_LINEAR_IO (5): the same 1x1 convolution reading -- and, in the weight
    gradient, writing -- the linear layer's compact tensors in their own (in/nc, out/nc) layout: no transposed copies
    (functional._linear_as_conv), against the reference fixtures of quaternion_linear / dual_quaternion_linear."""
    meta, d = load_golden(name)
    nw = 8 if alg == 5 else 4
    x = np.ascontiguousarray(d["x"], np.float32)                      # (rows, in)
    gy = np.ascontiguousarray(d["gy"], np.float32)                    # (rows, out)
    rows, fin = x.shape
    fout = gy.shape[1]
    ws = [np.ascontiguousarray(d["w%d" % i], np.float32) for i in range(nw)]                   # (in/nc, out/nc) as stored
    desc = ConvDesc(alg, 0, 1, 1, fin, fout, 1, rows, 1, 1, 1, 1, 0, 0, 1, 1)
    xc = np.ascontiguousarray(x.T[None], np.float32)                  # (1, in, rows)
    y = np.full((1, fout, rows), np.nan, np.float32)
    b = np.ascontiguousarray(d["b"], np.float32) if "b" in d else None
    assert emul.emul_conv(ctypes.byref(desc), 0, fptr(xc), ptr_array(ws), None if b is None else fptr(b), fptr(y)) == 0
    assert A.rel_err(y[0].T, d["y"]) < 1e-5
    gyc = np.ascontiguousarray(gy.T[None], np.float32)
    gx = np.full(xc.shape, np.nan, np.float32)
    assert emul.emul_conv(ctypes.byref(desc), 1, fptr(gyc), ptr_array(ws), None, fptr(gx)) == 0
    assert A.rel_err(gx[0].T, d["gx"]) < 1e-5
    gws = [np.zeros_like(w) for w in ws]
    assert emul.emul_conv_wgrad(ctypes.byref(desc), fptr(xc), fptr(gyc), ptr_array(gws), 2) == 0
    for i in range(nw):
        assert A.rel_err(gws[i], d["gw%d" % i]) < 1e-5, i


# ---- channels-last tensor-core path: the host plan (csrc/conv_cl_plan.h) through a CPU model of the kernel's data flow ---
# tests/host_emul/emul.cpp run_cl_fprop: real plan_fprop (fusion sets, MMA op table, epilogue column / sign table, row-shared
# taps, chunk masks, unit schedule) + real pack mapping, plain-loop "MMAs" in double.  A wrong table entry, tile offset, slot
# or sign gives an O(1) error; the GPU tests then only have to vouch for the hardware semantics (descriptors, TMA, TMEM).
CL_GOLDEN = ["conv1d_dq_c48_d3", "conv1d_q_c32_d2", "conv2d_dq_c24", "conv2d_q_c16"]
POLICIES = {"default": {}, "unfused": {"SELDQ_PAIR_FUSE": "0"}, "pairs": {"SELDQ_QUAD_FUSE": "0"},
            "pairs_double": {"SELDQ_QUAD_FUSE": "2", "SELDQ_ACC_DOUBLE": "1"}, "no_row_sharing": {"SELDQ_NO_RS": "1"}}


class _Env(object):
    def __init__(self, env):
        self.env = env

    def __enter__(self):
        self.saved = {k: os.environ.get(k) for k in ("SELDQ_PAIR_FUSE", "SELDQ_QUAD_FUSE", "SELDQ_ACC_DOUBLE", "SELDQ_NO_RS")}
        for k in self.saved:
            os.environ.pop(k, None)
        os.environ.update(self.env)

    def __exit__(self, *exc):
        for k, v in self.saved.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v


def _cl_conv(emul, desc, pass_, inp, ws, out_shape, n_sms=148):
    out = np.full(out_shape, np.nan, np.float32)
    info = (ctypes.c_int32 * 10)()
    rc = emul.emul_cl_conv(ctypes.byref(desc), pass_, fptr(inp), ptr_array(ws), fptr(out), n_sms, info)
    assert rc == 0, emul.emul_last_error()
    return out, dict(zip(("fuse", "pair_xor", "gc", "ngroups", "rs", "tps", "acc_cols", "acc_stages", "nstages", "smem"), info))


@pytest.mark.parametrize("policy", sorted(POLICIES))
@pytest.mark.parametrize("name", CL_GOLDEN)
def test_tensor_core_plan_reproduces_golden_conv(emul, name, policy):
    meta, d = load_golden(name)
    nw = NW[meta["algebra"]]
    desc = conv_desc(meta, d)
    desc.precision = 1
    x = np.ascontiguousarray(d["x"], np.float32)
    gy = np.ascontiguousarray(d["gy"], np.float32)
    ws = [np.ascontiguousarray(d["w%d" % i], np.float32) for i in range(nw)]
    bias = d["b"].reshape((1, -1) + (1,) * (x.ndim - 2)) if meta["bias"] else 0.0
    with _Env(POLICIES[policy]):
        for n_sms in (148, 2):                      # few SMs: no splitting of the out components over CTAs
            y, info = _cl_conv(emul, desc, 0, x, ws, d["y"].shape, n_sms)
            assert A.rel_err(y, d["y"] - bias) < 1e-5, (info, A.rel_err(y, d["y"] - bias))
            gx, info_d = _cl_conv(emul, desc, 1, gy, ws, x.shape, n_sms)
            assert A.rel_err(gx, d["gx"]) < 1e-5, (info_d, A.rel_err(gx, d["gx"]))
            if policy == "unfused":
                assert info["fuse"] == 0 and info_d["fuse"] == 0
            if policy == "pairs":
                assert info["fuse"] == 2 and info_d["fuse"] == 2 and info["pair_xor"] != info_d["pair_xor"]


@pytest.mark.parametrize("alg,cc_in,cc_out,ks,dil,shape", [
    ("DQ", 48, 48, (3,), 5, (1, 300)),           # TCN residual block: pairs (four quad sets would need 768 columns)
    ("DQ", 48, 48, (3,), 55, (1, 400)),          # the widest dilation of the stack: a 238-row shared box
    ("DQ", 48, 48, (1,), 1, (2, 131)),           # skip / residual convolution
    ("DQ", 24, 24, (3, 3), 1, (1, 5, 140)),      # CNN block: quads
    ("Q", 16, 16, (3, 3), 1, (1, 4, 130)),       # quaternion model, CNN block
    ("Q", 32, 32, (3,), 2, (2, 200)),            # quaternion model, residual block
    ("DQ", 8, 16, (3,), 3, (1, 260)),            # unequal channel counts
    ("Q", 40, 24, (1,), 1, (1, 129)),            # channels per component not a multiple of 16 on the K side
])
def test_tensor_core_plan_matches_oracle_on_random_layers(emul, alg, cc_in, cc_out, ks, dil, shape):
    rng = np.random.default_rng(7)
    nc = NW[alg]
    nd = len(ks)
    x = rng.standard_normal((shape[0], nc * cc_in) + shape[1:]).astype(np.float32)
    ws = [(0.3 * rng.standard_normal((cc_out, cc_in) + ks)).astype(np.float32) for _ in range(nc)]
    pad = dil * (ks[-1] - 1) // 2
    y_ref = A.qconv(x.astype(np.float64), [w.astype(np.float64) for w in ws], None, 1, pad, dil, alg)
    gy = rng.standard_normal(y_ref.shape).astype(np.float32)
    gx_ref, _, _ = A.qconv_backward(x.astype(np.float64), [w.astype(np.float64) for w in ws], gy.astype(np.float64), 1, pad,
                                    dil, alg)
    one_d = nd == 1
    desc = ConvDesc(ALG[alg], 1, nd, x.shape[0], x.shape[1], nc * cc_out, 1 if one_d else x.shape[2], x.shape[-1],
                    1 if one_d else ks[0], ks[-1], 1, 1, 0 if one_d else pad, pad, 1 if one_d else dil, dil)
    for policy in ("default", "pairs_double", "unfused"):
        with _Env(POLICIES[policy]):
            y, info = _cl_conv(emul, desc, 0, x, ws, y_ref.shape)
            assert A.rel_err(y, y_ref) < 1e-5, (policy, info, A.rel_err(y, y_ref))
            gx, info_d = _cl_conv(emul, desc, 1, gy, ws, x.shape)
            assert A.rel_err(gx, gx_ref) < 1e-5, (policy, info_d, A.rel_err(gx, gx_ref))
            if policy == "default":
                want = 4 if cc_out <= 32 and cc_out % 8 == 0 else (2 if cc_out % 8 == 0 else 0)
                assert info["fuse"] == want, info
                assert info["acc_cols"] * info["acc_stages"] <= 512 and info["smem"] <= 227 * 1024 - 12 * 1024, info
                if ks[-1] == 3 and dil <= 64 and cc_in * nc >= 64:
                    assert info["rs"] == 1 and info["tps"] == 3, info


@pytest.mark.parametrize("cc,ks,dil,shape", [(48, (3,), 5, (1, 4800)), (48, (1,), 1, (1, 4800)), (48, (3,), 55, (4, 520)),
                                             (16, (3,), 2, (2, 200))])
def test_tensor_core_plan_of_a_sibling_launch(emul, cc, ks, dil, shape):
    """seldq_conv_pair (conv_cl.h `nprob`): filter / gate or skip / residual convolutions of a residual block in one
    launch.  The plan for n_sms / 2 CTAs per problem must reproduce the convolution, visit every unit of both
    problems exactly once (checked inside the emulation), and double-buffer its accumulators."""
    rng = np.random.default_rng(9)
    x = rng.standard_normal((shape[0], 8 * cc, shape[1])).astype(np.float32)
    ws = [(0.3 * rng.standard_normal((cc, cc) + ks)).astype(np.float32) for _ in range(8)]
    pad = dil * (ks[-1] - 1) // 2
    desc = ConvDesc(ALG["DQ"], 1, 1, x.shape[0], x.shape[1], 8 * cc, 1, x.shape[-1], 1, ks[-1], 1, 1, 0, pad, 1, dil)
    if shape[1] > 1000:       # full TCN length: the oracle is slow there, compare with the single-launch emulation
        y_ref, _ = _cl_conv(emul, desc, 0, x, ws, x.shape)
        gx_ref = None
    else:
        y_ref = A.qconv(x.astype(np.float64), [w.astype(np.float64) for w in ws], None, 1, pad, dil, "DQ")
        gx_ref, _, _ = A.qconv_backward(x.astype(np.float64), [w.astype(np.float64) for w in ws], x.astype(np.float64), 1,
                                        pad, dil, "DQ")
    for pass_, ref in ((0, y_ref), (1, gx_ref)):
        if ref is None:
            continue
        out = np.full(x.shape, np.nan, np.float32)
        info = (ctypes.c_int32 * 10)()
        rc = emul.emul_cl_conv_pair(ctypes.byref(desc), pass_, fptr(x), ptr_array(ws), fptr(out), 148, info)
        assert rc == 0, emul.emul_last_error()
        assert A.rel_err(out, ref) < 1e-5
        if cc == 48 and shape == (1, 4800):
            # batch-1 TCN layer: 38 tiles x 4 pair groups = 152 units for 74 CTAs, accumulators double-buffered
            assert info[2] == 2 and info[3] == 4 and info[7] == 2, list(info)


def test_tensor_core_plan_dq_linear_as_1x1_convolution(emul):
    """dual_quaternion_linear's block table (the transpose of the convolution's) through the same planner."""
    meta, d = load_golden("linear_dq_c48")
    x = np.ascontiguousarray(d["x"], np.float32)
    gy = np.ascontiguousarray(d["gy"], np.float32)
    ws = [np.ascontiguousarray(d["w%d" % i], np.float32) for i in range(8)]
    desc = LinearDesc(3, 1, x.shape[0], x.shape[1], gy.shape[1])          # SELDQ_ALG_DQ_LINEAR
    for policy in ("default", "unfused"):
        with _Env(POLICIES[policy]):
            y = np.full(gy.shape, np.nan, np.float32)
            info = (ctypes.c_int32 * 10)()
            assert emul.emul_cl_linear(ctypes.byref(desc), 0, fptr(x), ptr_array(ws), fptr(y), 148, info) == 0, emul.emul_last_error()
            assert A.rel_err(y, d["y"] - d["b"]) < 1e-5, (policy, list(info))
            gx = np.full(x.shape, np.nan, np.float32)
            assert emul.emul_cl_linear(ctypes.byref(desc), 1, fptr(gy), ptr_array(ws), fptr(gx), 148, info) == 0, emul.emul_last_error()
            assert A.rel_err(gx, d["gx"]) < 1e-5, (policy, list(info))


@pytest.mark.parametrize("no_tps", [False, True])
def test_tensor_core_plan_first_layer_dense_mode(emul, no_tps):
    """The first CNN layer (one input channel per component): the activation is ONE 16-channel-padded component, the
    signed expanded weight tile is built in shared memory, and all nine taps share one stage (tps = 9)."""
    rng = np.random.default_rng(11)
    x = rng.standard_normal((1, 8, 6, 140)).astype(np.float32)
    ws = [(0.3 * rng.standard_normal((24, 1, 3, 3))).astype(np.float32) for _ in range(8)]
    y_ref = A.qconv(x.astype(np.float64), [w.astype(np.float64) for w in ws], None, 1, 1, 1, "DQ")
    desc = ConvDesc(ALG["DQ"], 1, 2, 1, 8, 192, 6, 140, 3, 3, 1, 1, 1, 1, 1, 1)
    saved = os.environ.pop("SELDQ_NO_TPS", None)
    try:
        if no_tps:
            os.environ["SELDQ_NO_TPS"] = "1"
        y, info = _cl_conv(emul, desc, 0, x, ws, y_ref.shape)
    finally:
        os.environ.pop("SELDQ_NO_TPS", None)
        if saved is not None:
            os.environ["SELDQ_NO_TPS"] = saved
    assert A.rel_err(y, y_ref) < 1e-5, (info, A.rel_err(y, y_ref))
    assert info["fuse"] == 0 and info["tps"] == (1 if no_tps else 9) and info["rs"] == 0, info
    # the golden first-layer fixture, forward and input gradient (both passes are dense here)
    meta, d = load_golden("conv2d_dq_first")
    gd = conv_desc(meta, d)
    gd.precision = 1
    gx_in = np.ascontiguousarray(d["x"], np.float32)
    gws = [np.ascontiguousarray(d["w%d" % i], np.float32) for i in range(8)]
    bias = d["b"].reshape(1, -1, 1, 1) if meta["bias"] else 0.0
    y2, _ = _cl_conv(emul, gd, 0, gx_in, gws, d["y"].shape)
    assert A.rel_err(y2, d["y"] - bias) < 1e-5
    gx2, _ = _cl_conv(emul, gd, 1, np.ascontiguousarray(d["gy"], np.float32), gws, gx_in.shape)
    assert A.rel_err(gx2, d["gx"]) < 1e-5


@pytest.mark.parametrize("n_sms", [148, 5])
@pytest.mark.parametrize("name", CL_GOLDEN)
def test_tensor_core_wgrad_plan_reproduces_golden(emul, name, n_sms):
    """csrc/wgrad_cl.cu's host plan (o tiles, tap groups, split-K ranges, the (a, b) -> compact tensor fold table)
    through a CPU model of the kernel: dense accumulator per CTA and split, sign-weighted fold, accumulation across
    splits.  n_sms = 5 forces few or no splits."""
    meta, d = load_golden(name)
    nw = NW[meta["algebra"]]
    desc = conv_desc(meta, d)
    desc.precision = 1
    x = np.ascontiguousarray(d["x"], np.float32)
    gy = np.ascontiguousarray(d["gy"], np.float32)
    gws = [np.zeros(d["w%d" % i].shape, np.float32) for i in range(nw)]
    info = (ctypes.c_int32 * 8)()
    rc = emul.emul_cl_conv_wgrad(ctypes.byref(desc), fptr(x), fptr(gy), ptr_array(gws), n_sms, info)
    assert rc == 0, emul.emul_last_error()
    for i in range(nw):
        assert A.rel_err(gws[i], d["gw%d" % i]) < 1e-5, (i, list(info), A.rel_err(gws[i], d["gw%d" % i]))
    assert info[5] <= 227 * 1024 and info[2] * info[7] <= info[6] <= 512, list(info)


# ---- rotation variants and quaternion point-wise operators (csrc/rotation.cuh) --------------------------
ROT_FIXTURES = golden_names("rot_conv") + golden_names("rot_convT") + golden_names("rot_linear")


def _rot_weights(d):
    ws = [np.ascontiguousarray(d["w%d" % i], np.float32) for i in range(4)]
    d0, d1 = ws[0].shape[:2]
    return ws, d0, d1, int(np.prod(ws[0].shape[2:], dtype=np.int64))


def _emul_rotation_weight(emul, ws, qf, tr):
    d0, d1 = ws[0].shape[:2]
    taps = int(np.prod(ws[0].shape[2:], dtype=np.int64))
    nc = 4 if qf else 3
    out = np.full(((nc * d1, nc * d0) if tr else (nc * d0, nc * d1)) + ws[0].shape[2:], np.nan, np.float32)
    assert emul.emul_rotation_weight(ptr_array(ws), ctypes.c_longlong(d0), ctypes.c_longlong(d1), ctypes.c_longlong(taps),
                                     int(qf), int(tr), fptr(out)) == 0
    return out


@pytest.mark.parametrize("name", ROT_FIXTURES)
def test_emulated_rotation_weight_is_the_oracles_float32_weight(emul, name):
    """The kernel's element function, thread index by thread index: bit-identical to the oracle's float32 restatement;
    against the weight the reference itself built in float32 (W32, read back through an identity input) to the last bit
    or the one before it (torch's vectorised CPU square root is not correctly rounded)."""
    meta, d = load_golden(name)
    ws, d0, d1, taps = _rot_weights(d)
    qf = meta["quaternion_format"]
    want = A.rotation_weight(ws, qf, np.float32)
    got = _emul_rotation_weight(emul, ws, qf, False)
    assert np.array_equal(got, want)
    assert np.abs(got.reshape(d["W32"].shape).astype(np.float64) - d["W32"]).max() <= 1.5e-7 * np.abs(d["W32"]).max()
    got_t = _emul_rotation_weight(emul, ws, qf, True)
    assert np.array_equal(got_t, np.swapaxes(want, 0, 1))


@pytest.mark.parametrize("name", ROT_FIXTURES)
def test_emulated_rotation_variants_match_golden(emul, name):
    """Rotation weight (emulated kernel) -> real-algebra convolution through the emulated fp32 kernels and, for stride 1,
    through the tensor-core path's host plan (dense mode) -> compact gradients through the emulated backward kernel,
    against the reference's outputs and autograd gradients."""
    meta, d = load_golden(name)
    ws, d0, d1, taps = _rot_weights(d)
    qf, kind = meta["quaternion_format"], meta["kind"]
    x = np.ascontiguousarray(d["x"], np.float32)
    gy = np.ascontiguousarray(d["gy"], np.float32)
    bias = np.ascontiguousarray(d["b"], np.float32) if meta["bias"] else None
    if kind == "rot_linear":
        # 1 x 1 convolution over the flattened rows; its (out, in, 1) weight is the transposed rotation weight
        W = _emul_rotation_weight(emul, [w[..., None] for w in ws], qf, True)
        rows = int(np.prod(x.shape[:-1]))
        xin = np.ascontiguousarray(x.reshape(rows, -1).T[None])
        gout = np.ascontiguousarray(gy.reshape(rows, -1).T[None])
        desc = ConvDesc(ALG["R"], 0, 1, 1, xin.shape[1], W.shape[0], 1, rows, 1, 1, 1, 1, 0, 0, 1, 1)
        to_ref = lambda t: t[0].T.reshape(d["y"].shape if t.shape[1] == W.shape[0] else x.shape)
    else:
        W = _emul_rotation_weight(emul, ws, qf, False)
        one_d = meta["ndim"] == 1
        s, p, dl = meta["stride"], meta["padding"], meta["dilation"]
        to_ref = lambda t: t
        if kind == "rot_conv":
            xin, gout = x, gy
            desc = ConvDesc(ALG["R"], 0, meta["ndim"], x.shape[0], x.shape[1], W.shape[0], 1 if one_d else x.shape[2],
                            x.shape[-1], 1 if one_d else W.shape[2], W.shape[-1], 1 if one_d else s, s, 0 if one_d else p, p,
                            1 if one_d else dl, dl)
        else:
            # transposed convolution = input-gradient pass of the convolution (N, cout, out) -> (N, cin, in) whose weight
            # is W read as (out', in'); its "input" is the fixture's output and the other way round
            xin, gout = gy, x
            yshape = d["y"].shape
            desc = ConvDesc(ALG["R"], 0, meta["ndim"], yshape[0], yshape[1], W.shape[0], 1 if one_d else yshape[2],
                            yshape[-1], 1 if one_d else W.shape[2], W.shape[-1], 1 if one_d else s, s, 0 if one_d else p, p,
                            1 if one_d else dl, dl)
    fwd_in, fwd_ref = (xin, d["y"]) if kind != "rot_convT" else (gout, d["y"])
    bwd_in, bwd_ref = (gout, d["gx"]) if kind != "rot_convT" else (xin, d["gx"])
    b_arr = (bias.reshape((1, -1) + (1,) * (d["y"].ndim - 2)) if kind != "rot_linear" else bias) if bias is not None else 0.0
    for prec in (0, 1):
        if prec == 1 and meta["stride"] != 1:
            continue
        desc.precision = prec
        # pass 0 of the descriptor's convolution, pass 1 = its input gradient
        p_y, p_gx = (0, 1) if kind != "rot_convT" else (1, 0)
        y_shape, gx_shape = ((1, W.shape[0], xin.shape[2]), xin.shape) if kind == "rot_linear" else (d["y"].shape, x.shape)
        if prec == 0:
            y = np.full(y_shape, np.nan, np.float32)
            assert emul.emul_conv(ctypes.byref(desc), p_y, fptr(fwd_in), ptr_array([W]), None, fptr(y)) == 0, emul.emul_last_error()
            gx = np.full(gx_shape, np.nan, np.float32)
            assert emul.emul_conv(ctypes.byref(desc), p_gx, fptr(bwd_in), ptr_array([W]), None, fptr(gx)) == 0, emul.emul_last_error()
            gW = np.zeros_like(W)
            assert emul.emul_conv_wgrad(ctypes.byref(desc), fptr(xin), fptr(gout), ptr_array([gW]), 3) == 0, emul.emul_last_error()
        else:
            y, _ = _cl_conv(emul, desc, p_y, fwd_in, [W], y_shape)
            gx, _ = _cl_conv(emul, desc, p_gx, bwd_in, [W], gx_shape)
            gW = np.zeros_like(W)
            info = (ctypes.c_int32 * 8)()
            rc = emul.emul_cl_conv_wgrad(ctypes.byref(desc), fptr(xin), fptr(gout), ptr_array([gW]), 148, info)
            if rc != 0:                     # the channels-last weight-gradient kernel does not serve this layer: the
                gW = None                   # library routes it to csrc/wgrad_umma.cu, which has no host model
        assert A.rel_err(to_ref(y) + b_arr, d["y"]) < 1e-5, (prec, A.rel_err(to_ref(y) + b_arr, d["y"]))
        assert A.rel_err(to_ref(gx), d["gx"]) < 1e-5, (prec, A.rel_err(to_ref(gx), d["gx"]))
        if gW is None:
            continue
        gws = [np.full_like(w, np.nan) for w in ws]
        if kind == "rot_linear":
            assert emul.emul_rotation_weight_bwd(ptr_array(ws), fptr(gW), ctypes.c_longlong(d0), ctypes.c_longlong(d1),
                                                 ctypes.c_longlong(1), int(qf), 1, ptr_array(gws)) == 0
        else:
            assert emul.emul_rotation_weight_bwd(ptr_array(ws), fptr(gW), ctypes.c_longlong(d0), ctypes.c_longlong(d1),
                                                 ctypes.c_longlong(taps), int(qf), 0, ptr_array(gws)) == 0
        for i in range(4):
            assert A.rel_err(gws[i], d["gw%d" % i]) < 2e-5, (prec, i, A.rel_err(gws[i], d["gw%d" % i]))


@pytest.mark.parametrize("name", golden_names("qpointwise"))
def test_emulated_quaternion_pointwise_operators_match_golden(emul, name):
    meta, d = load_golden(name)
    a, b, g = (np.ascontiguousarray(d[k], np.float32) for k in ("a", "b", "g"))
    outer, m = a.shape[0], a.size // (4 * a.shape[0])

    def run(op, p, q):
        out = np.full_like(a, np.nan)
        assert emul.emul_qpointwise(op, fptr(p), fptr(q) if q is not None else None, fptr(out), ctypes.c_longlong(outer),
                                    ctypes.c_longlong(m)) == 0
        return out

    ham = run(0, a, b)
    assert np.array_equal(ham, A.hamilton_product(a, b, np.float32))          # same products, same order of the sums
    assert A.rel_err(ham, d["ham"]) < 1e-6
    assert A.rel_err(run(1, g, b), d["ham_ga"]) < 1e-6
    assert A.rel_err(run(2, a, g), d["ham_gb"]) < 1e-6
    assert np.array_equal(run(3, a, None), A.q_normalize(a, np.float32))
    assert A.rel_err(run(3, a, None), d["norm"]) < 1e-6
    assert A.rel_err(run(4, a, g), d["norm_g"]) < 1e-5
    assert A.rel_err(run(5, a, None), d["exp"]) < 1e-6
    assert A.rel_err(run(6, a, g), d["exp_g"]) < 1e-5


@pytest.mark.parametrize("name", golden_names("convT"))
def test_emulated_transposed_conv_is_the_input_gradient_pass(emul, name):
    """quaternion_transpose_conv (quaternion_ops.py:149-172) as functional.block_conv_transpose runs it: the input-gradient
    pass of the convolution (N, cout, out) -> (N, cin, in) with the same stride / padding / dilation on the compact
    tensors read as (out', in'); its own gradients are that convolution's forward and weight-gradient passes with the
    two activations swapped.  Includes stride > 1 with output_padding (fp32 kernels)."""
    meta, d = load_golden(name)
    ws = [np.ascontiguousarray(d["w%d" % i], np.float32) for i in range(4)]
    x = np.ascontiguousarray(d["x"], np.float32)
    gy = np.ascontiguousarray(d["gy"], np.float32)
    one_d = meta["ndim"] == 1
    s, p, dl = meta["stride"], meta["padding"], meta["dilation"]
    ys = d["y"].shape
    desc = ConvDesc(ALG["Q"], 0, meta["ndim"], ys[0], ys[1], x.shape[1], 1 if one_d else ys[2], ys[-1],
                    1 if one_d else ws[0].shape[2], ws[0].shape[-1], 1 if one_d else s, s, 0 if one_d else p, p,
                    1 if one_d else dl, dl)
    bias = d["b"].reshape((1, -1) + (1,) * (x.ndim - 2)) if meta["bias"] else 0.0
    y = np.full(ys, np.nan, np.float32)
    assert emul.emul_conv(ctypes.byref(desc), 1, fptr(x), ptr_array(ws), None, fptr(y)) == 0, emul.emul_last_error()
    assert A.rel_err(y + bias, d["y"]) < 1e-5
    gx = np.full(x.shape, np.nan, np.float32)
    assert emul.emul_conv(ctypes.byref(desc), 0, fptr(gy), ptr_array(ws), None, fptr(gx)) == 0, emul.emul_last_error()
    assert A.rel_err(gx, d["gx"]) < 1e-5
    gws = [np.zeros_like(w) for w in ws]
    assert emul.emul_conv_wgrad(ctypes.byref(desc), fptr(gy), fptr(x), ptr_array(gws), 3) == 0, emul.emul_last_error()
    for i in range(4):
        assert A.rel_err(gws[i], d["gw%d" % i]) < 1e-5, i
