"""Bring-up probes for the tcgen05 / TMA path (run on a B200: `python tools/umma_probe.py`).

Feeds host-built shared-memory images and raw descriptors to seldq_probe_umma / seldq_probe_tma_load
and reports which layout hypotheses reproduce A @ B^T.  Results go to gpurun_out/umma_probe.json.
Nothing in the product path depends on this file; it documents (and re-checks) the descriptor
conventions conv_umma.cu / wgrad_umma.cu rely on.
"""
import ctypes
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sound-event-localization-and-detection_b200")
L = pkg._lib
sys.path.insert(0, os.path.join(ROOT, "tools", "probe"))
import probe_lib  # noqa: E402
lib = probe_lib.lib()

_P = ctypes.c_void_p
lib.seldq_probe_tensor_map.argtypes = [_P, _P, ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_uint64),
                                       ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint32), ctypes.c_int32]
lib.seldq_probe_tma_load.argtypes = [_P, _P, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), ctypes.c_uint32,
                                     ctypes.c_uint32, _P, ctypes.c_uint32, _P]
lib.seldq_probe_umma.argtypes = [_P, ctypes.c_uint32, _P, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64,
                                 ctypes.c_uint32, ctypes.c_int32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int32, _P, _P]
lib.seldq_probe_last_error.restype = ctypes.c_char_p
SW_NONE, SW_128 = 0, 2


def check(rc):
    assert rc == 0, (rc, lib.seldq_probe_last_error())



def bf16_bits(a):
    t = torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(torch.bfloat16)
    return t.view(torch.int16).numpy().astype(np.uint16), t.float().numpy()


def desc_hi(lbo, sbo, swizzle):
    return (((lbo >> 4) & 0x3fff) << 16) | (((sbo >> 4) & 0x3fff) << 32) | (1 << 46) | ((swizzle & 7) << 61)


def idesc(m, n, a_mn, b_mn, neg_a=0, neg_b=0):
    return (1 << 4) | (1 << 7) | (1 << 10) | (neg_a << 13) | (neg_b << 14) | (a_mn << 15) | (b_mn << 16) | \
        ((n >> 3) << 17) | ((m >> 4) << 24)


def swz128(byte_off):
    """Swizzle<3,4,3> on a byte offset inside a 1024-byte aligned region."""
    return byte_off ^ (((byte_off >> 7) & 7) << 4)


def run_umma(a_img, b_img, a_desc, b_desc, idsc, n_mma, a_step, b_step, n_cols):
    a = torch.from_numpy(a_img.view(np.uint8).copy()).cuda()
    b = torch.from_numpy(b_img.view(np.uint8).copy()).cuda()
    out = torch.full((128, n_cols), float("nan"), device="cuda")
    rc = lib.seldq_probe_umma(a.data_ptr(), a.numel(), b.data_ptr(), b.numel(), a_desc, b_desc, idsc, n_mma,
                              a_step, b_step, n_cols, out.data_ptr(), None)
    check(rc)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def image_mn_major_sw128(A_bits, katoms):
    """A is (128 m, K) -> [m_half][k_atom][8 k rows x 128 B], 16-byte chunks XOR-swizzled by the row."""
    img = np.zeros(2 * katoms * 1024 // 2, np.uint16)
    for m in range(128):
        for k in range(katoms * 8):
            mh, mi = divmod(m, 64)
            ka, kr = divmod(k, 8)
            off = mh * katoms * 1024 + ka * 1024 + swz128(kr * 128 + mi * 2)
            img[off // 2] = A_bits[m, k]
    return img


def image_k_major_noswizzle(B_bits, n_rows, kpairs):
    """B is (n_rows, K) -> per kpair tile: [k chunk (2)][row group][8 rows x 16 B]."""
    tile = n_rows * 32
    img = np.zeros(kpairs * tile // 2, np.uint16)
    for n in range(n_rows):
        for k in range(kpairs * 16):
            kp, kk = divmod(k, 16)
            kc, kj = divmod(kk, 8)
            off = kp * tile + kc * (n_rows // 8) * 128 + (n // 8) * 128 + (n % 8) * 16 + kj * 2
            img[off // 2] = B_bits[n, k]
    return img


def image_k_major_sw128(M_bits, rows):
    """(rows, 64 k) -> row r at r*128 B, chunks swizzled (what a [64 k x rows] TMA box writes)."""
    img = np.zeros(rows * 64, np.uint16)
    for r in range(rows):
        for k in range(64):
            img[swz128(r * 128 + k * 2) // 2] = M_bits[r, k]
    return img


def err(out, ref):
    return float(np.abs(out - ref).max() / max(np.abs(ref).max(), 1e-9))


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else "all"
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    rng = np.random.default_rng(0)
    res = {}
    if only in ("all", "t1"):
        probe_t1(res)
    if only in ("all", "u1"):
        probe_u1(res, rng)
    if only in ("all", "u2"):
        probe_u2(res, rng)
    if only in ("all", "u3"):
        probe_u3(res, rng)
    for k, v in res.items():
        print("%-24s %.3e %s" % (k, v, "PASS" if v < 1e-2 else "fail"))
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "umma_probe_%s.json" % only), "w"), indent=1)


def probe_u1(res, rng):

    # ---- U1: fprop operand conventions: A MN-major SW128, B K-major no swizzle ---------------------
    N, katoms = 48, 6
    K = katoms * 8
    A = rng.standard_normal((128, K))
    B = rng.standard_normal((N, K))
    Ab, Af = bf16_bits(A)
    Bb, Bf = bf16_bits(B)
    ref = Af @ Bf.T
    a_img = image_mn_major_sw128(Ab, katoms)
    b_img = image_k_major_noswizzle(Bb, N, katoms // 2)
    half = katoms * 1024
    variants = {
        "as_designed": (half, 1024, N * 16, 128),
        "a_swapped": (1024, half, N * 16, 128),
        "b_swapped": (half, 1024, 128, N * 16),
        "both_swapped": (1024, half, 128, N * 16),
    }
    pick = os.environ.get("U1_VARIANT")
    for name, (a_lbo, a_sbo, b_lbo, b_sbo) in variants.items():
        if pick and name != pick:
            continue
        out = run_umma(a_img, b_img, desc_hi(a_lbo, a_sbo, SW_128), desc_hi(b_lbo, b_sbo, SW_NONE),
                       idesc(128, N, 1, 0), katoms // 2, 2048 // 16, (N * 32) // 16, N)
        res["U1_" + name] = err(out, ref)
    if pick and pick != "as_designed":
        return
    # negate-B bit
    out = run_umma(a_img, b_img, desc_hi(half, 1024, SW_128), desc_hi(N * 16, 128, SW_NONE),
                   idesc(128, N, 1, 0, 0, 1), katoms // 2, 2048 // 16, (N * 32) // 16, N)
    res["U1_negate_b"] = err(out, -ref)
    out = run_umma(a_img, b_img, desc_hi(half, 1024, SW_128), desc_hi(N * 16, 128, SW_NONE),
                   idesc(128, N, 1, 0, 1, 0), katoms // 2, 2048 // 16, (N * 32) // 16, N)
    res["U1_negate_a"] = err(out, -ref)



def probe_u2(res, rng):
    # ---- U2: wgrad operand conventions: both K-major SW128, K advanced by 32 B inside the row -------
    NW = 128
    G = rng.standard_normal((128, 64))
    X = rng.standard_normal((NW, 64))
    Gb, Gf = bf16_bits(G)
    Xb, Xf = bf16_bits(X)
    ref2 = Gf @ Xf.T
    g_img = image_k_major_sw128(Gb, 128)
    x_img = image_k_major_sw128(Xb, NW)
    for name, (lbo, sbo) in {"as_designed": (16, 1024), "lbo0": (0, 1024), "swapped": (1024, 16)}.items():
        out = run_umma(g_img, x_img, desc_hi(lbo, sbo, SW_128), desc_hi(lbo, sbo, SW_128), idesc(128, NW, 0, 0),
                       4, 2, 2, NW)
        res["U2_" + name] = err(out, ref2)



def probe_u3(res, rng):
    # ---- U3: A operand starting at a row that is not a multiple of 8 inside a K-major SW128 tile: the three
    #      w-taps of a 3x3 convolution as row-shifted views of ONE [130 w x 64 ch] TMA box.  Does the tensor core
    #      derive the swizzle phase from the address (base offset 0), or does the descriptor's base-offset field
    #      (bits 49..51) have to carry row & 7?
    N = 48
    A = rng.standard_normal((144, 64))
    B = rng.standard_normal((N, 64))
    Ab, Af = bf16_bits(A)
    Bb, Bf = bf16_bits(B)
    a_img = image_k_major_sw128(Ab, 144)
    b_img = image_k_major_noswizzle(Bb, N, 4)
    for r in (0, 1, 2, 5, 8, 11):
        ref = Af[r:r + 128] @ Bf.T
        for name, bo in (("addr", 0), ("baseoff", r & 7)):
            a_desc = desc_hi(16, 1024, SW_128) + r * 8 + (bo << 49)
            out = run_umma(a_img, b_img, a_desc, desc_hi(N * 16, 128, SW_NONE), idesc(128, N, 0, 0), 4, 2, (N * 32) // 16, N)
            res["U3_row%d_%s" % (r, name)] = err(out, ref)


def probe_t1(res):
    # ---- T1: TMA box of a bf16 (C, W) tensor: rank 2 / rank 4, no swizzle / 128B swizzle,
    #      descriptor passed as a __grid_constant__ kernel parameter or through global memory
    import time
    C, W = 40, 200
    src = torch.arange(C * W, dtype=torch.float32).reshape(C, W) % 251
    src16 = src.to(torch.bfloat16).cuda()
    variant = os.environ.get("T1_VARIANT", "r4_sw128_param")
    rank = 4 if variant.startswith("r4") else 2
    swz = 3 if "sw128" in variant else 0
    via_global = variant.endswith("global")
    tmap = (ctypes.c_uint8 * 128)()
    if rank == 4:
        dims = (ctypes.c_uint64 * 4)(W, 1, C, 1)
        strides = (ctypes.c_uint64 * 3)(W * 2, W * 2, C * W * 2)
        box = (ctypes.c_uint32 * 4)(64, 1, 16, 1)
    else:
        dims = (ctypes.c_uint64 * 2)(W, C)
        strides = (ctypes.c_uint64 * 1)(W * 2)
        box = (ctypes.c_uint32 * 2)(64, 16)
    check(lib.seldq_probe_tensor_map(tmap, src16.data_ptr(), 2, rank, dims, strides, box, swz))
    print("tensor map words:", [hex(w) for w in np.frombuffer(bytes(tmap), np.uint32)[:16]])
    dmap = torch.from_numpy(np.frombuffer(bytes(tmap), np.uint8).copy()).cuda() if via_global else None
    cases = {"interior": (8, 8), "right_oob": (176, 24), "chan_oob": (0, 32), "neg_aligned": (-8, 0),
             "neg_aligned_big": (-64, 8), "chan_neg": (16, -3), "pos_unaligned": (5, 0), "neg_unaligned": (-5, 0)}
    only_case = os.environ.get("T1_CASE")
    for name, (w0, c0) in cases.items():
        if only_case and name != only_case:
            continue
        dump = torch.zeros(16 * 64, dtype=torch.int16, device="cuda")
        coords = (ctypes.c_int32 * 4)(w0, 0, c0, 0) if rank == 4 else (ctypes.c_int32 * 4)(w0, c0, 0, 0)
        t0 = time.time()
        try:
            check(lib.seldq_probe_tma_load(tmap, dmap.data_ptr() if via_global else None, rank, coords, 16 * 128, 0,
                                             dump.data_ptr(), 16 * 128, None))
            torch.cuda.synchronize()
        finally:
            print("T1 %s %s: %.2f s" % (variant, name, time.time() - t0), flush=True)
        got = dump.view(torch.bfloat16).float().cpu().numpy()
        want = np.zeros(16 * 64, np.float32)
        for r in range(16):
            for j in range(64):
                w, c = w0 + j, c0 + r
                v = float(src[c, w]) if (0 <= w < W and 0 <= c < C) else 0.0
                off = r * 128 + j * 2
                want[(swz128(off) if swz else off) // 2] = v
        res["T1_" + name] = float(np.abs(got - want).max())


if __name__ == "__main__":
    main()
