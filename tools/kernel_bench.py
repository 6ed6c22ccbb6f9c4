"""Per-kernel timings on a B200 (CUDA events, warm-up, L2 flushed between iterations).
Usage: python tools/kernel_bench.py [--batch 1 4] [--out gpurun_out/kbench.json]"""
import argparse
import importlib
import json
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sound-event-localization-and-detection_b200")
L = pkg._lib

_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    _flush.fill_(1)


def timeit(fn, iters=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def conv_layers(B):
    # (name, algebra, x shape, cout, k, pad, dil) for DQSELD-TCN-S1-PHI_8ch
    T = 4800
    return [
        ("cnn0_8to192", L.ALG_DQ, (B, 8, 256, T), 192, 3, 1, 1),
        ("cnn1_192", L.ALG_DQ, (B, 192, 32, T), 192, 3, 1, 1),
        ("cnn2_192", L.ALG_DQ, (B, 192, 4, T), 192, 3, 1, 1),
        ("tcn_k3_d1", L.ALG_DQ, (B, 384, T), 384, 3, 1, 1),
        ("tcn_k3_d55", L.ALG_DQ, (B, 384, T), 384, 3, 55, 55),
        ("tcn_k1", L.ALG_DQ, (B, 384, T), 384, 1, 0, 1),
        ("q_tcn_k3_d5", L.ALG_Q, (B, 128, T), 128, 3, 5, 5),
    ]


def flops(alg, xs, cout, k):
    nz = 0.75 if alg == L.ALG_DQ else 1.0
    pos = 1
    for s in xs[2:]:
        pos *= s
    taps = k ** (len(xs) - 2)
    return 2.0 * nz * cout * xs[1] * taps * pos * xs[0]


def bench_conv(B, precs, res):
    for name, alg, xs, cout, k, pad, dil in conv_layers(B):
        nc = 4 if alg == L.ALG_Q else 8
        kshape = (k,) * (len(xs) - 2)
        x = torch.randn(xs, device="cuda")
        ws = [(0.05 * torch.randn((cout // nc, xs[1] // nc) + kshape, device="cuda")).requires_grad_(True)
              for _ in range(nc)]
        for prec in precs:
            if prec == "fp32" and name.startswith("cnn") and B > 1:
                continue
            key = "%s_B%d_%s" % (name, B, prec)
            try:
                with pkg.precision(prec):
                    xin = x.clone().requires_grad_(name != "cnn0_8to192")
                    y = pkg.block_conv(xin, ws, None, 1, pad, dil, alg)
                    gy = torch.randn_like(y)
                    f = flops(alg, xs, cout, k)
                    t_f = timeit(lambda: pkg.block_conv(xin, ws, None, 1, pad, dil, alg))

                    def fb():
                        for w in ws:
                            w.grad = None
                        xin.grad = None
                        yy = pkg.block_conv(xin, ws, None, 1, pad, dil, alg)
                        yy.backward(gy)
                    t_fb = timeit(fb)
                res[key] = dict(fwd_ms=t_f, fwd_bwd_ms=t_fb, fwd_tflops=f / t_f / 1e9,
                                train_tflops=(3 if xin.requires_grad else 2) * f / t_fb / 1e9)
                print(key, res[key], flush=True)
                del y, gy
            except Exception as e:  # keep going: one unsupported shape must not hide the others
                res[key] = dict(error=repr(e))
                print(key, "ERROR", repr(e), flush=True)
                traceback.print_exc()
        del x, ws
        torch.cuda.empty_cache()


def bench_stft(res):
    for B in (1, 4, 16):
        x = 0.1 * torch.randn(B, 8, 1_920_000, device="cuda")
        for phase in (False, True):
            t = timeit(lambda: pkg.stft_magphase(x, 512, 112, True, phase, True), iters=7)
            byt = B * (8 * 1_920_000 * 4 + (2 if phase else 1) * 8 * 256 * 4800 * 4)
            res["stft_B%d_%s" % (B, "magphase" if phase else "mag")] = dict(
                ms=t, us_per_clip=1e3 * t / B, gbs=byt / t / 1e6)
            print("stft", B, phase, res["stft_B%d_%s" % (B, "magphase" if phase else "mag")], flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, nargs="+", default=[1, 4])
    ap.add_argument("--prec", nargs="+", default=["bf16", "fp32"])
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "kbench.json"))
    ap.add_argument("--skip-stft", action="store_true")
    ap.add_argument("--skip-conv", action="store_true")
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    res = {}
    if not args.skip_stft:
        try:
            bench_stft(res)
        except Exception as e:
            res["stft_error"] = repr(e)
            traceback.print_exc()
    for B in ([] if args.skip_conv else args.batch):
        bench_conv(B, args.prec, res)
        json.dump(res, open(args.out, "w"), indent=1)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
