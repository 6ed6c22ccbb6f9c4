"""Timeline of CTA 0 of one TCN convolution launch (seldq_debug_fprop_trace): where a launch's ~10 us go."""
import ctypes, importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sound-event-localization-and-detection_b200")
L, F = pkg._lib, pkg.functional
lib = L.lib()
N, C, T = int(os.environ.get("N", 1)), 384, 4800
dev = torch.device("cuda")
st = lambda: torch.cuda.current_stream().cuda_stream
buf = torch.zeros(64, dtype=torch.int64, device=dev)
for k, dil, pair in ((3, 5, False), (3, 5, True), (1, 1, True)):
    pad = dil * (k - 1) // 2
    d = L.ConvDesc(L.ALG_DQ, L.PREC_BF16, 1, N, C, C, 1, T, 1, k, 1, 1, 0, pad, 1, dil)
    x_cl, _ = F.stage_operand(torch.randn(N, C, T, device=dev), d, 0)
    wa = [0.05 * torch.randn(C // 8, C // 8, k, device=dev) for _ in range(8)]
    pa = F.packed_weights(wa, d, L.PASS_FWD, cache=False)
    ya, yb = torch.zeros(N, C, T, device=dev), torch.zeros(N, C, T, device=dev)
    wpa = L.ptr_array([w.data_ptr() for w in wa])
    e = L.ConvEpilogue()

    def run():
        if pair:
            L.check(lib.seldq_conv_pair(ctypes.byref(d), L.PASS_FWD, x_cl.data_ptr(), x_cl.data_ptr(), pa.data_ptr(), pa.data_ptr(),
                                        ya.data_ptr(), yb.data_ptr(), ctypes.byref(e), ctypes.byref(e), st()))
        else:
            L.check(lib.seldq_conv_fwd(ctypes.byref(d), None, x_cl.data_ptr(), wpa, pa.data_ptr(), None, ya.data_ptr(), None, 0, st()))

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    lib.seldq_debug_fprop_trace(buf.data_ptr())
    buf.zero_()
    torch.cuda.synchronize()
    run()
    torch.cuda.synchronize()
    lib.seldq_debug_fprop_trace(None)
    t = buf.cpu().tolist()
    t0 = t[0]
    rel = lambda i: (t[i] - t0) / 1e3 if t[i] else float("nan")
    print("== k%d dil %d %s (CTA 0; us since kernel entry)" % (k, dil, "pair" if pair else "single"))
    print("   set-up done %.2f | weights resident %.2f | first stage landed %.2f | exit %.2f" % (rel(1), rel(2), rel(3), rel(7)))
    for pc in range(3):
        if t[40 + 4 * pc]:
            print("   unit 0, piece %d of warp 2: start %.2f  TMEM loaded %.2f  combined %.2f  stored %.2f" % (
                pc, rel(40 + 4 * pc), rel(41 + 4 * pc), rel(42 + 4 * pc), rel(43 + 4 * pc)))
    for u in range(6):
        if t[8 + 4 * u]:
            print("   unit %d: MMA start %.2f  issued %.2f | epilogue start %.2f  done %.2f" % (
                u, rel(8 + 4 * u), rel(9 + 4 * u), rel(10 + 4 * u), rel(11 + 4 * u)))
