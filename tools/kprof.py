"""Device time per kernel (torch.profiler) of forward + backward of the DQSELD-TCN convolution layers,
one layer at a time:  python tools/kprof.py [--batch B] [--layers cnn0,cnn1,...]"""
import argparse
import importlib
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sound-event-localization-and-detection_b200")
L = pkg._lib

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--layers", default="cnn0,cnn1,cnn2,tcn3,tcn1")
ap.add_argument("--iters", type=int, default=5)
args = ap.parse_args()
B = args.batch
LAYERS = {
    "cnn0": ((B, 8, 256, 4800), 192, 3, 1, 1, False),
    "cnn1": ((B, 192, 32, 4800), 192, 3, 1, 1, True),
    "cnn2": ((B, 192, 4, 4800), 192, 3, 1, 1, True),
    "tcn3": ((B, 384, 4800), 384, 3, 5, 5, True),
    "tcn1": ((B, 384, 4800), 384, 1, 0, 1, True),
}
torch.manual_seed(0)
for name in args.layers.split(","):
    xs, cout, k, pad, dil, need_gx = LAYERS[name]
    ks = (k,) * (len(xs) - 2)
    x = torch.randn(xs, device="cuda").requires_grad_(need_gx)
    ws = [(0.05 * torch.randn((cout // 8, xs[1] // 8) + ks, device="cuda")).requires_grad_(True) for _ in range(8)]
    gy = None

    def run():
        global gy
        with pkg.precision("bf16"):
            y = pkg.block_conv(x, ws, None, 1, pad, dil, L.ALG_DQ)
            if gy is None:
                gy = torch.randn_like(y)
            y.backward(gy)
            with torch.no_grad():
                ws[0].add_(0.0)           # bump the version counter: weights are re-packed as in training

    for _ in range(2):
        run()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.iters):
            run()
        torch.cuda.synchronize()
    rows = {}
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            r = rows.setdefault(ev.name[:70], [0.0, 0])
            r[0] += ev.device_time
            r[1] += 1
    nz = 0.75
    pos = 1
    for s in xs[2:]:
        pos *= s
    flop = 2.0 * nz * cout * xs[1] * (k ** (len(xs) - 2)) * pos * xs[0]
    print("== %s B=%d  (%.2f GFLOP per pass)" % (name, B, flop / 1e9))
    for kn, (us, n) in sorted(rows.items(), key=lambda kv: -kv[1][0]):
        per = us / n
        extra = "  %.0f TFLOP/s" % (flop / per / 1e6) if ("fprop" in kn or "wgrad_kernel" in kn) else ""
        print("  %8.1f us x %4.1f/iter  %s%s" % (per, n / args.iters, kn, extra))
