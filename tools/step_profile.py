"""Per-kernel breakdown of ONE eager training step (torch.profiler, CUDA activities): which kernels
the step spends its device time in, ours and PyTorch's.  Writes a table to stdout."""
import argparse
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="DQSELD-TCN-S1-PHI_8ch")
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--top", type=int, default=45)
ap.add_argument("--ops", action="store_true", help="also list the aten ops (with input shapes) by self device time")
args = ap.parse_args()
cfg = bench.CONFIGS[args.config]
pkg = importlib.import_module(bench.PKG)
trainer_mod = importlib.import_module(bench.PKG + ".trainer")
dev = torch.device("cuda", 0)
np.random.seed(1)
torch.manual_seed(1)
batch = args.batch or cfg["batch_size"]
model = pkg.SELD_Model(time_dim=bench.TIME_DIM, **bench.model_kwargs(cfg)).to(dev).train()
trainer = trainer_mod.Trainer(model, lr=1e-4, n_sed=bench.N_SED)
x, t = bench.synth_batch(pkg, cfg, batch, 1234, dev)
for _ in range(3):
    trainer.step(x, t)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=args.ops) as prof:
    trainer.step(x, t)
    torch.cuda.synchronize()
rows = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name[:90]
        r = rows.setdefault(name, [0.0, 0, []])
        r[0] += ev.device_time
        r[1] += 1
        r[2].append(ev.device_time)
total = sum(r[0] for r in rows.values())
print("total device time %.3f ms over %d launches" % (total / 1e3, sum(r[1] for r in rows.values())))
for name, (us, n, each) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:args.top]:
    print("%9.1f us %5d x %7.1f us  %5.1f%%  %s" % (us, n, us / n, 100 * us / total, name))
    if 1 < n <= 10 and us > 100:
        print("             each: " + " ".join("%.0f" % e for e in each))
if args.ops:
    print()
    print("aten ops by self device time (the part of the step that is not this repository's kernels):")
    ka = [e for e in prof.key_averages(group_by_input_shape=True) if e.key.startswith("aten::") and e.self_device_time_total > 0]
    tot = sum(e.self_device_time_total for e in ka)
    print("  total %.1f us over %d calls" % (tot, sum(e.count for e in ka)))
    for e in sorted(ka, key=lambda e: -e.self_device_time_total)[:45]:
        print("  %8.1f us %4d x  %-34s %s" % (e.self_device_time_total, e.count, e.key, str(e.input_shapes)[:110]))
    print()
    print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=60,
                                                             max_name_column_width=48, max_shapes_column_width=70))
