"""Op-level sweep of the CUDA convolution against the numpy oracle over the odd shapes the small
models produce (short rows, big dilations, k = 1, dense mode).  Prints one line per case."""
import importlib
import itertools
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sound-event-localization-and-detection_b200")
from oracle import algebra as A  # noqa: E402


def case(alg, N, cin, cout, spatial, k, dil, prec, seed=0):
    rng = np.random.default_rng(seed)
    nc = 4 if alg == "Q" else 8
    nd = len(spatial)
    ws = [(0.2 * rng.standard_normal((cout // nc, cin // nc) + (k,) * nd)).astype(np.float32) for _ in range(nc)]
    x = rng.standard_normal((N, cin) + tuple(spatial)).astype(np.float32)
    pad = dil * (k - 1) // 2
    y_ref = A.qconv(x.astype(np.float64), [w.astype(np.float64) for w in ws], None, 1, pad, dil, alg)
    gy = rng.standard_normal(y_ref.shape).astype(np.float32)
    gx_ref, gw_ref, _ = A.qconv_backward(x.astype(np.float64), [w.astype(np.float64) for w in ws],
                                         gy.astype(np.float64), 1, pad, dil, alg)
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    wt = [torch.from_numpy(w).cuda().requires_grad_(True) for w in ws]
    with pkg.precision(prec):
        y = pkg.block_conv(xt, wt, None, 1, pad, dil, pkg._lib.ALG_Q if alg == "Q" else pkg._lib.ALG_DQ)
        y.backward(torch.from_numpy(gy).cuda())
    torch.cuda.synchronize()
    e_y = A.rel_err(y.detach().cpu().numpy(), y_ref)
    e_x = A.rel_err(xt.grad.cpu().numpy(), gx_ref)
    e_w = max(A.rel_err(wt[i].grad.cpu().numpy(), gw_ref[i]) for i in range(nc))
    return e_y, e_x, e_w


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    tol = 2e-2 if prec == "bf16" else 1e-4
    bad = 0
    cases = []
    for W, k, dil in itertools.product((8, 16, 32, 64, 100, 160, 264), (1, 3), (1, 5, 55)):
        if k == 1 and dil != 1:
            continue
        cases.append(("DQ", 2, 16, 16, (W,), k, dil))       # dense mode (2 channels per component)
        cases.append(("DQ", 2, 128, 128, (W,), k, dil))     # compact mode, 16 per component
        cases.append(("DQ", 1, 64, 128, (W,), k, dil))      # 8 -> 16 per component
        cases.append(("Q", 2, 32, 64, (W,), k, dil))        # 8 -> 16 per component
    for H, W in ((2, 64), (16, 160), (1, 40), (3, 8)):
        cases.append(("DQ", 2, 8, 64, (H, W), 3, 1))        # first layer of the mid model (dense)
        cases.append(("DQ", 2, 64, 64, (H, W), 3, 1))       # 8 per component
        cases.append(("DQ", 1, 16, 16, (H, W), 3, 1))       # tiny model (dense)
        cases.append(("Q", 1, 64, 64, (H, W), 3, 1))
    for c in cases:
        try:
            e = case(*c, prec)
            ok = max(e) < tol
        except Exception as ex:  # noqa: BLE001
            e, ok = (repr(ex)[:100],), False
        bad += not ok
        print(("ok  " if ok else "BAD ") + str(c) + " " + " ".join("%.2e" % v if isinstance(v, float) else v for v in e),
              flush=True)
    print("sweep %s: %d cases, %d bad" % (prec, len(cases), bad))


if __name__ == "__main__":
    main()
