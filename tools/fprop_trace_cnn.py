"""CTA-0 timeline of the second CNN block's convolution: forward with the fp16 output of the fused path against the
input-gradient pass with its fp32 output (why does dgrad take twice as long?)."""
import ctypes, importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sound-event-localization-and-detection_b200")
L, F = pkg._lib, pkg.functional
lib = L.lib()
dev = torch.device("cuda")
st = lambda: torch.cuda.current_stream().cuda_stream
buf = torch.zeros(64, dtype=torch.int64, device=dev)
H, W = 32, 4800
d = L.ConvDesc(L.ALG_DQ, L.PREC_BF16, 2, 1, 192, 192, H, W, 3, 3, 1, 1, 1, 1, 1, 1)
x_cl, _ = F.stage_operand(torch.randn(1, 192, H, W, device=dev), d, 0)
ws = [0.05 * torch.randn(24, 24, 3, 3, device=dev) for _ in range(8)]
wp = L.ptr_array([w.data_ptr() for w in ws])
pf, pd = F.packed_weights(ws, d, L.PASS_FWD, cache=False), F.packed_weights(ws, d, L.PASS_DGRAD, cache=False)
y32 = torch.zeros(1, 192, H, W, device=dev)
y16 = torch.zeros(1, 192, H, W, device=dev, dtype=torch.float16)
cases = {
    "forward, fp16 output": lambda: lib.seldq_conv_fwd_bf16(ctypes.byref(d), None, x_cl.data_ptr(), wp, pf.data_ptr(), None, y16.data_ptr(), None, 0, st()),
    "forward, fp32 output": lambda: lib.seldq_conv_fwd(ctypes.byref(d), None, x_cl.data_ptr(), wp, pf.data_ptr(), None, y32.data_ptr(), None, 0, st()),
    "dgrad, fp32 output": lambda: lib.seldq_conv_dgrad(ctypes.byref(d), None, x_cl.data_ptr(), wp, pd.data_ptr(), y32.data_ptr(), None, 0, st()),
}
for name, run in cases.items():
    for _ in range(3):
        L.check(run())
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        L.check(run())
    b.record()
    torch.cuda.synchronize()
    us = 100 * a.elapsed_time(b)
    lib.seldq_debug_fprop_trace(buf.data_ptr())
    buf.zero_()
    torch.cuda.synchronize()
    L.check(run())
    torch.cuda.synchronize()
    lib.seldq_debug_fprop_trace(None)
    t = buf.cpu().tolist()
    rel = lambda i: (t[i] - t[0]) / 1e3 if t[i] else float("nan")
    print("== cnn1 %s: %.1f us per launch" % (name, us))
    for pc in range(3):
        if t[40 + 4 * pc]:
            print("   unit 0, piece %d of warp 2: start %.2f  TMEM loaded %.2f  combined %.2f  stored %.2f" % (
                pc, rel(40 + 4 * pc), rel(41 + 4 * pc), rel(42 + 4 * pc), rel(43 + 4 * pc)))
    for u in range(4):
        if t[8 + 4 * u]:
            print("   unit %d: MMA start %.2f  issued %.2f | epilogue start %.2f  done %.2f" % (
                u, rel(8 + 4 * u), rel(9 + 4 * u), rel(10 + 4 * u), rel(11 + 4 * u)))
