"""Small driver for ncu: one forward + backward of selected DQSELD-TCN layers at per-GPU batch B."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sound-event-localization-and-detection_b200")
L = pkg._lib
B = int(os.environ.get("NCU_B", "1"))
which = os.environ.get("NCU_LAYERS", "cnn0,cnn1,tcn3,tcn1").split(",")
LAYERS = {
    "cnn0": ((B, 8, 256, 4800), 192, 3, 1, 1, False),
    "cnn1": ((B, 192, 32, 4800), 192, 3, 1, 1, True),
    "cnn2": ((B, 192, 4, 4800), 192, 3, 1, 1, True),
    "tcn3": ((B, 384, 4800), 384, 3, 5, 5, True),
    "tcn1": ((B, 384, 4800), 384, 1, 0, 1, True),
}
torch.manual_seed(0)
for name in which:
    xs, cout, k, pad, dil, need_gx = LAYERS[name]
    ks = (k,) * (len(xs) - 2)
    x = torch.randn(xs, device="cuda").requires_grad_(need_gx)
    ws = [(0.05 * torch.randn((cout // 8, xs[1] // 8) + ks, device="cuda")).requires_grad_(True) for _ in range(8)]
    with pkg.precision("bf16"):
        y = pkg.block_conv(x, ws, None, 1, pad, dil, L.ALG_DQ)
        y.backward(torch.randn_like(y))
    torch.cuda.synchronize()
    print(name, "done", flush=True)
