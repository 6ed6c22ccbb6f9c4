"""Kernel-level timing of the TCN convolution launches (CUDA events, 200 back-to-back launches each):
single launches, sibling launches (seldq_conv_pair) and the fused-glue epilogue variants.
    python tools/pair_bench.py [--k 3 --dil 5]"""
import argparse
import ctypes
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sound-event-localization-and-detection_b200")
L, F = pkg._lib, pkg.functional
lib = L.lib()
ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--batch", type=int, default=1)
args = ap.parse_args()
N, C, T = args.batch, 384, 4800
dev = torch.device("cuda")
st = lambda: torch.cuda.current_stream().cuda_stream


def timeit(fn, iters=args.iters):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / iters


for k, dil in ((3, 5), (1, 1)):
    pad = dil * (k - 1) // 2
    d = L.ConvDesc(L.ALG_DQ, L.PREC_BF16, 1, N, C, C, 1, T, 1, k, 1, 1, 0, pad, 1, dil)
    x = torch.randn(N, C, T, device=dev)
    x_cl, _ = F.stage_operand(x, d, 0)
    x2_cl, _ = F.stage_operand(torch.randn(N, C, T, device=dev), d, 0)
    wa = [0.05 * torch.randn(C // 8, C // 8, k, device=dev) for _ in range(8)]
    wb = [0.05 * torch.randn(C // 8, C // 8, k, device=dev) for _ in range(8)]
    for pass_, pname in ((L.PASS_FWD, "fwd"), (L.PASS_DGRAD, "dgrad")):
        pa, pb = F.packed_weights(wa, d, pass_, cache=False), F.packed_weights(wb, d, pass_, cache=False)
        ya, yb = torch.zeros(N, C, T, device=dev), torch.zeros(N, C, T, device=dev)
        add = torch.randn(N, C, T, device=dev)
        sa, sb = torch.zeros(C, 2, dtype=torch.float64, device=dev), torch.zeros(C, 2, dtype=torch.float64, device=dev)
        wpa = L.ptr_array([w.data_ptr() for w in wa])

        def single():
            L.check(lib.seldq_conv_fwd(ctypes.byref(d), None, x_cl.data_ptr(), wpa, pa.data_ptr(), None, ya.data_ptr(), None, 0,
                                       st()) if pass_ == L.PASS_FWD else
                    lib.seldq_conv_dgrad(ctypes.byref(d), None, x_cl.data_ptr(), wpa, pa.data_ptr(), ya.data_ptr(), None, 0, st()))

        def epi(mode=0, addend=None, stats=None):
            e = L.ConvEpilogue()
            e.mode, e.addend, e.stats = mode, None if addend is None else addend.data_ptr(), None if stats is None else stats.data_ptr()
            return e

        def pair(ea, eb, second=x_cl):
            return lambda: L.check(lib.seldq_conv_pair(ctypes.byref(d), pass_, x_cl.data_ptr(), second.data_ptr(), pa.data_ptr(),
                                                       pb.data_ptr(), ya.data_ptr(), yb.data_ptr(), ctypes.byref(ea),
                                                       ctypes.byref(eb), st()))

        def one_epi(e):
            return lambda: L.check(lib.seldq_conv_epi(ctypes.byref(d), pass_, x_cl.data_ptr(), pa.data_ptr(), ya.data_ptr(),
                                                      ctypes.byref(e), st()))

        rows = [("single launch", single),
                ("single + stats", one_epi(epi(stats=sa))),
                ("single accumulate", one_epi(epi(1))),
                ("pair plain", pair(epi(), epi())),
                ("pair plain, two inputs", pair(epi(), epi(), x2_cl)),
                ("pair plain through the glue kernel", pair(epi(3), epi(3))),
                ("single plain through the glue kernel", one_epi(epi(3))),
                ("pair + stats both", pair(epi(stats=sa), epi(stats=sb))),
                ("pair accumulate | add", pair(epi(1), epi(2, add))),
                ("pair accumulate | add + stats", pair(epi(1), epi(2, add, sb)))]
        print("== k%d dil %d %s  N=%d" % (k, dil, pname, N))
        for name, fn in rows:
            print("   %-32s %7.2f us" % (name, timeit(fn)))
