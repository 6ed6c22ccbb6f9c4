"""How close the bf16 whole-model gradient checks of tests/test_gpu_parity.py run to their gates: worst ratio
error / gate over all tensors of a fixture, over repeated runs (the tensor path has run-to-run noise)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity as T      # noqa: E402
from oracle import algebra as A  # noqa: E402
seldq = importlib.import_module("sound-event-localization-and-detection_b200")
runs = int(os.environ.get("RUNS", "5"))
for name in T.MODEL_FIXTURES:
    for fused in (False, True):
        worst = []
        for _ in range(runs):
            prev = seldq.fused.ENABLED
            seldq.fused.ENABLED = fused
            try:
                meta, d, sed, doa, loss, grads = T._run_model(seldq, name, "bf16")
            finally:
                seldq.fused.ENABLED = prev
            if meta["cfg"]["domain"] == "R":
                continue
            emu = "bf16emu16" if (fused and name in T._FUSED_CNN) else "bf16emu"
            w = (0.0, "")
            for k, g in grads.items():
                ref, em = d["grad/" + k].astype(np.float64), d[emu + "_grad/" + k].astype(np.float64)
                nrm = max(float(np.linalg.norm(ref)), 1e-300)
                e2, n2 = float(np.linalg.norm(g - ref)) / nrm, float(np.linalg.norm(em - ref)) / nrm
                e, noise = A.rel_err(g, ref), A.rel_err(em, ref)
                r = max(e2 / max(2e-2, 2.5 * n2), e / max(2e-2, 3.0 * noise))
                if r > w[0]:
                    w = (r, k)
            worst.append(w)
        if worst:
            print("%-24s fused=%-5s worst error/gate per run: %s  (%s)" % (
                name, fused, " ".join("%.2f" % r for r, _ in worst), max(worst)[1]), flush=True)

# fp32 mode (test_model_fp32_matches_reference_fixture): worst error / gate with the gate at max(floor, F x the
# reference's own float32 error) for F = 2 and F = 4
for name in T.MODEL_FIXTURES:
    rows = []
    for _ in range(min(runs, 3)):
        meta, d, sed, doa, loss, grads = T._run_model(seldq, name, "fp32")
        floor = 2e-3 if meta["cfg"]["domain"] == "R" else 5e-4
        w2 = w4 = (0.0, "")
        for k, g in grads.items():
            e, r32 = A.rel_err(g, d["grad/" + k]), float(d["ref32err/" + k])
            w2 = max(w2, (e / max(floor, 2.0 * r32), k))
            w4 = max(w4, (e / max(floor, 4.0 * r32), k))
        rows.append((w2, w4))
    print("%-24s fp32 worst error/gate per run, gate 2x: %s  gate 4x: %s  (%s)" % (
        name, " ".join("%.2f" % a[0] for a, _ in rows), " ".join("%.2f" % b[0] for _, b in rows), max(rows)[0][1]), flush=True)
