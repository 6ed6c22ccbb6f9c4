"""Kernel-level timing of the TCN weight-gradient launch (seldq_conv_wgrad_pair) versus its split-K factor:
time = fixed part (prologue, epilogue: TMEM -> fold -> atomics) + K steps per CTA x time per K step."""
import ctypes, importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sound-event-localization-and-detection_b200")
L, F = pkg._lib, pkg.functional
lib = L.lib()
N, C, T = int(os.environ.get("N", 1)), 384, 4800
dev = torch.device("cuda")


def timeit(fn, iters=100):
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
    x_cl, _ = F.stage_operand(torch.randn(N, C, T, device=dev), d, 0)
    _, ga = F.stage_operand(torch.randn(N, C, T, device=dev), d, 1, want_cl=False, want_t16=True)
    _, gb = F.stage_operand(torch.randn(N, C, T, device=dev), d, 1, want_cl=False, want_t16=True)
    gwa = [torch.zeros(C // 8, C // 8, k, device=dev) for _ in range(8)]
    gwb = [torch.zeros(C // 8, C // 8, k, device=dev) for _ in range(8)]
    pa, pb = L.ptr_array([g.data_ptr() for g in gwa]), L.ptr_array([g.data_ptr() for g in gwb])
    st = torch.cuda.current_stream().cuda_stream
    fn = lambda: L.check(lib.seldq_conv_wgrad_pair(ctypes.byref(d), x_cl.data_ptr(), ga.data_ptr(), gb.data_ptr(), pa, pb, 1, st))
    for splits in (0, 1, 2, 4, 8, 16, 24):
        if splits:
            os.environ["SELDQ_WGRAD_SPLITS"] = str(splits)
        else:
            os.environ.pop("SELDQ_WGRAD_SPLITS", None)
        print("k%d dil %d  splits %2s : %7.2f us" % (k, dil, splits or "auto", timeit(fn)))
