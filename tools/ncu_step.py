"""One eager training step of DQSELD-TCN-S1-PHI_8ch between cudaProfilerStart / Stop, for Nsight Compute:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \\
        --log-file gpurun_out/launches.csv python tools/ncu_step.py
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:<kernel> \\
        -o gpurun_out/prof python tools/ncu_step.py [--stft]
"""
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
ap.add_argument("--stft", action="store_true", help="also run the STFT front end inside the profiled range")
args = ap.parse_args()
cfg = bench.CONFIGS[args.config]
pkg = importlib.import_module(bench.PKG)
trainer_mod = importlib.import_module(bench.PKG + ".trainer")
dev = torch.device("cuda", 0)
np.random.seed(1)
torch.manual_seed(1)
model = pkg.SELD_Model(time_dim=bench.TIME_DIM, **bench.model_kwargs(cfg)).to(dev).train()
trainer = trainer_mod.Trainer(model, lr=1e-4, n_sed=bench.N_SED)
x, t = bench.synth_batch(pkg, cfg, cfg["batch_size"], 1234, dev)
wav = 0.1 * torch.randn(1, 8, 1_920_000, device=dev)
for _ in range(2):
    trainer.step(x, t)
    pkg.stft_magphase(wav, 512, 112, True, True, True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
trainer.step(x, t)
if args.stft:
    pkg.stft_magphase(wav, 512, 112, True, False, True)
    pkg.stft_magphase(wav, 512, 112, True, True, True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", flush=True)
