"""Timeline of CTA (0, 0) of one TCN weight-gradient launch (seldq_debug_fprop_trace stamps in qconv_cl_wgrad_kernel):
where the ~17 us a launch costs "whatever its size" go."""
import ctypes, importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sound-event-localization-and-detection_b200")
L, F = pkg._lib, pkg.functional
lib = L.lib()
N, C = int(os.environ.get("N", 1)), 384
dev = torch.device("cuda")
st = lambda: torch.cuda.current_stream().cuda_stream
buf = torch.zeros(64, dtype=torch.int64, device=dev)


def timeit(fn, iters=50):
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


for k, dil, T in ((3, 5, 4800), (1, 1, 4800), (1, 1, 256)):
    pad = dil * (k - 1) // 2
    d = L.ConvDesc(L.ALG_DQ, L.PREC_BF16, 1, N, C, C, 1, T, 1, k, 1, 1, 0, pad, 1, dil)
    x_cl, _ = F.stage_operand(torch.randn(N, C, T, device=dev), d, 0)
    _, ga = F.stage_operand(torch.randn(N, C, T, device=dev), d, 1, want_cl=False, want_t16=True)
    _, gb = F.stage_operand(torch.randn(N, C, T, device=dev), d, 1, want_cl=False, want_t16=True)
    gwa = [torch.zeros(C // 8, C // 8, k, device=dev) for _ in range(8)]
    gwb = [torch.zeros(C // 8, C // 8, k, device=dev) for _ in range(8)]
    pa, pb = L.ptr_array([g.data_ptr() for g in gwa]), L.ptr_array([g.data_ptr() for g in gwb])
    run = lambda: L.check(lib.seldq_conv_wgrad_pair(ctypes.byref(d), x_cl.data_ptr(), ga.data_ptr(), gb.data_ptr(), pa, pb, 1, st()))
    for splits in (0, 1):
        if splits:
            os.environ["SELDQ_WGRAD_SPLITS"] = str(splits)
        else:
            os.environ.pop("SELDQ_WGRAD_SPLITS", None)
        us = timeit(run)
        lib.seldq_debug_fprop_trace(buf.data_ptr())
        buf.zero_()
        torch.cuda.synchronize()
        run()
        torch.cuda.synchronize()
        lib.seldq_debug_fprop_trace(None)
        t = buf.cpu().tolist()
        rel = lambda i: (t[i] - t[0]) / 1e3 if t[i] else float("nan")
        print("== k%d dil %d T %d splits %s: %.2f us per launch (back-to-back launches)" % (k, dil, T, splits or "auto", us))
        print("   set-up done %.2f | first stage requested %.2f | landed %.2f | last MMA issued %.2f | accumulator seen %.2f | "
              "fold done %.2f | exit %.2f" % tuple(rel(i) for i in range(1, 8)))
        print("   K steps seen at: " + " ".join("%.2f" % rel(8 + i) for i in range(24) if t[8 + i]))
os.environ.pop("SELDQ_WGRAD_SPLITS", None)
