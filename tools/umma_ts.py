"""A-from-tensor-memory probe (tools/probe/umma_probe.cu, probe::umma_ts_kernel): does a K = 16 slab copied once to
tensor memory (tcgen05.cp.128x256b) and read from there by G small MMAs beat G shared-memory-operand MMAs that
each re-read the slab?  Run on a B200:  python tools/umma_ts.py"""
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
L.seldq_probe_umma_ts.argtypes = [ctypes.c_uint32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                  ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
L.seldq_probe_umma_ts.restype = ctypes.c_int


def run(n, n_slabs, g, mode, nbuf, blocks=1):
    out = torch.zeros(2 * blocks, dtype=torch.int64, device="cuda")
    for _ in range(1 if mode == 0 else 2):
        if mode == 0:
            out.zero_()
        rc = L.seldq_probe_umma_ts(n, n_slabs, g, mode, nbuf, blocks, out.data_ptr(), None)
        assert rc == 0, rc
        torch.cuda.synchronize()
    return out.view(blocks, 2)


if __name__ == "__main__":
    for n in (32, 48, 96):
        for g in (1, 3, 4):
            for nbuf in (1, 2, 4):
                o = run(n, 4, g, 0, nbuf)
                print(f"check N {n:3d} G {g} nbuf {nbuf}: mismatches {int(o[0, 0])}  non-zero {int(o[0, 1])} of {128 * n}", flush=True)
    names = {1: "TS only (A resident)", 2: "cp + G TS MMAs", 3: "cp only", 4: "G SS MMAs (baseline)"}
    for blocks in (1, 148):
        for n in (16, 32, 48, 64, 96):
            for mode in (1, 2, 4, 3):
                for g in ((1,) if mode == 3 else (1, 2, 3, 4, 6, 8)):
                    for nbuf in ((1, 4) if mode in (2, 3) else (1,)):
                        slabs = 2048
                        o = run(n, slabs, g, mode, nbuf, blocks).float()
                        iss, tot = o[:, 0].mean().item() / slabs, o[:, 1].mean().item() / slabs
                        print(f"blocks {blocks:3d} N {n:3d} {names[mode]:22s} G {g} nbuf {nbuf}: issue {iss:7.1f} cyc/slab, "
                              f"issue+drain {tot:7.1f} cyc/slab = {tot / g:6.1f} cyc/MMA  (math floor {n / 2:.0f})", flush=True)
