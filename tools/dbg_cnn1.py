import ctypes, importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sound-event-localization-and-detection_b200")
L, F = pkg._lib, pkg.functional
lib = L.lib()
which = sys.argv[1]
H, W = int(os.environ.get("H", 32)), int(os.environ.get("W", 4800))
d = L.ConvDesc(L.ALG_DQ, L.PREC_BF16, 2, 1, 192, 192, H, W, 3, 3, 1, 1, 1, 1, 1, 1)
x = torch.randn(1, 192, H, W, device="cuda")
x_cl, _ = F.stage_operand(x, d, 0)
ws = [0.05 * torch.randn(24, 24, 3, 3, device="cuda") for _ in range(8)]
wp = L.ptr_array([w.data_ptr() for w in ws])
y = torch.zeros(1, 192, H, W, device="cuda")
st = torch.cuda.current_stream().cuda_stream
if which == "fwd":
    pk = F.packed_weights(ws, d, L.PASS_FWD, cache=False)
    L.check(lib.seldq_conv_fwd(ctypes.byref(d), None, x_cl.data_ptr(), wp, pk.data_ptr(), None, y.data_ptr(), None, 0, st))
else:
    pk = F.packed_weights(ws, d, L.PASS_DGRAD, cache=False)
    L.check(lib.seldq_conv_dgrad(ctypes.byref(d), None, x_cl.data_ptr(), wp, pk.data_ptr(), y.data_ptr(), None, 0, st))
torch.cuda.synchronize()
print(which, H, W, "ok", float(y.abs().mean()))
