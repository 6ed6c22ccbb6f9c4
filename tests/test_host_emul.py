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
