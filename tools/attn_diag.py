"""Error budget of bf16 attention (SELDQ_ATTN_BF16=1) on the model fixtures: prints what the parity test asserts."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity as T
from oracle import algebra as A
seldq = importlib.import_module("sound-event-localization-and-detection_b200")
for name in ("model_dq_tiny", "model_dq_mid"):
    meta, d, sed, doa, loss, grads = T._run_model(seldq, name, "bf16")
    emu = "bf16emu16" if name == "model_dq_mid" else "bf16emu"
    print(name, "sed vs ref %.2e doa vs ref %.2e | sed vs emu %.2e doa vs emu %.2e" % (
        A.rel_err(sed, d["sed"]), A.rel_err(doa, d["doa"]), A.rel_err(sed, d[emu + "/sed"]), A.rel_err(doa, d[emu + "/doa"])))
    bad = []
    for k, g in grads.items():
        noise = A.rel_err(d[emu + "_grad/" + k], d["grad/" + k])
        e, tol = A.rel_err(g, d["grad/" + k]), max(2e-2, 2.0 * noise)
        if not e < tol:
            bad.append((k, e, tol))
    print("  gradient tensors outside max(2e-2, 2 x emulation noise): %d of %d" % (len(bad), len(grads)))
    for k, e, tol in sorted(bad, key=lambda t: -t[1] / t[2])[:8]:
        print("    %-60s err %.3e tol %.3e" % (k, e, tol))
