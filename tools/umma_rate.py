"""tcgen05.mma issue-rate probe (tools/probe/umma_probe.cu): cycles per MMA versus N for the product kernels'
operand layouts.  Run on a B200:  python tools/umma_rate.py"""
import ctypes
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("sound-event-localization-and-detection_b200")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "probe"))
import probe_lib  # noqa: E402
L = probe_lib.lib()
L.seldq_probe_umma_rate.argtypes = [ctypes.c_uint32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                    ctypes.c_void_p, ctypes.c_void_p]
L.seldq_probe_umma_rate.restype = ctypes.c_int


def run(n, n_mma, d_cycle, mode, blocks):
    out = torch.zeros(2 * blocks, dtype=torch.int64, device="cuda")
    for _ in range(2):
        rc = L.seldq_probe_umma_rate(n, n_mma, d_cycle, mode, blocks, out.data_ptr(), None)
        assert rc == 0, rc
        torch.cuda.synchronize()
    o = out.view(blocks, 2).float()
    return o[:, 0].mean().item() / n_mma, o[:, 1].mean().item() / n_mma


if __name__ == "__main__":
    for mode, label in ((8, "lane-parallel"), (16, "one thread + smem entries")):
        for n in (16, 32, 48, 96):
            for nl in (8, 24, 32):
                dc = max(1, min(4, 512 // n))
                iss, tot = run(n, 3840, dc | (nl << 8), mode, 1)
                print(f"{label:26s} N {n:3d} lanes/stage {nl:2d}: issue {iss:6.1f} cyc/MMA, issue+drain {tot:6.1f} cyc/MMA",
                      flush=True)
    for mode in (2,):
        for n in (16, 48, 96):
            iss, tot = run(n, 4096, 2, mode, 1)
            print(f"unrolled, descriptors fixed   N {n:3d}: issue {iss:6.1f} cyc/MMA, issue+drain {tot:6.1f} cyc/MMA", flush=True)
