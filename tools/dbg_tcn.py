import ctypes, importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sound-event-localization-and-detection_b200")
L, F = pkg._lib, pkg.functional
lib = L.lib()
dil, pair, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
N, C, T, k = 1, 384, 4800, 3
d = L.ConvDesc(L.ALG_DQ, L.PREC_BF16, 1, N, C, C, 1, T, 1, k, 1, 1, 0, dil, 1, dil)
x_cl, _ = F.stage_operand(torch.randn(N, C, T, device="cuda"), d, 0)
ws = [0.05 * torch.randn(48, 48, k, device="cuda") for _ in range(8)]
wp = L.ptr_array([w.data_ptr() for w in ws])
pk = F.packed_weights(ws, d, L.PASS_FWD, cache=False)
ya, yb = torch.zeros(N, C, T, device="cuda"), torch.zeros(N, C, T, device="cuda")
st = torch.cuda.current_stream().cuda_stream
e = L.ConvEpilogue()
for i in range(reps):
    if pair:
        L.check(lib.seldq_conv_pair(ctypes.byref(d), L.PASS_FWD, x_cl.data_ptr(), x_cl.data_ptr(), pk.data_ptr(), pk.data_ptr(),
                                    ya.data_ptr(), yb.data_ptr(), ctypes.byref(e), ctypes.byref(e), st))
    else:
        L.check(lib.seldq_conv_fwd(ctypes.byref(d), None, x_cl.data_ptr(), wp, pk.data_ptr(), None, ya.data_ptr(), None, 0, st))
torch.cuda.synchronize()
print("dil", dil, "pair", pair, "reps", reps, "ok", float(ya.abs().mean()))
