#!/usr/bin/env python
"""bench.py -- DQSELD-TCN training samples/s on N B200s of one node (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the reference's CPU arithmetic (oracle port)

A "step" is one full training step (train.py:546-561: zero_grad, forward, BCE+5*MSE loss, backward,
gradient all-reduce, Adam) on one synthetic L3DAS21-shaped batch per GPU.  One process per GPU
(torchrun for N > 1), batch-sharded data parallel, weak scaling.  Prints ONE JSON line on rank 0.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "sound-event-localization-and-detection_b200"

# hyper-parameters train.py passes for each config/SERVER_*.txt
COMMON = dict(freq_dim=256, output_classes=14, kernel_size_cnn_blocks=3, pool_size=[[8, 2], [8, 2], [2, 2]],
              pool_time="TCN", D=[10], dilation_mode="fibonacci", kernel_size_dilated_conv=3, V_kernel_size=3,
              fc_activations="linear", fc_dropout="Last", class_overlaps=3, use_bias_conv=0, use_bias_linear=1,
              batch_norm="BN", spatial_dropout_rate=0.5, dropout_perc=0.3)
CONFIGS = {
    "DQSELD-TCN-S1-PHI_8ch": dict(input_channels=8, domain="DQ", domain_classifier="DQ",
                                  cnn_filters=[192, 192, 192], G=384, U=384, V=[384, 384], fc_layers=[384],
                                  parallel_ConvTC_block="False", parallel_magphase=False, extra_name="_8ch",
                                  batch_size=1, conv_train_gflop=570.9),
    "DQSELD-TCN-S1-PHI_16chMagPhase": dict(input_channels=16, domain="DQ", domain_classifier="DQ",
                                           cnn_filters=[192, 192, 192], G=384, U=384, V=[384, 384],
                                           fc_layers=[384], parallel_ConvTC_block="False",
                                           parallel_magphase=False, extra_name="_16chMagPhase", batch_size=4,
                                           conv_train_gflop=621.9),
    "QSELD-TCN-S1-PHI_parallel_8ch": dict(input_channels=8, domain="Q", domain_classifier="R",
                                          cnn_filters=[64, 64, 64], G=128, U=128, V=[128, 128], fc_layers=[128],
                                          parallel_ConvTC_block="1", parallel_magphase=False,
                                          extra_name="_parallel_8ch", batch_size=4, conv_train_gflop=99.7),
    # config/SERVER_DQSELD-TCN-S1-PHI_micAMagPhaseParallelmicBMagPhase.txt: two DQ ConvTC branches (mic A / mic B,
    # magnitude + phase each, model.py:463-471), real-valued heads
    "DQSELD-TCN-S1-PHI_micAMagPhaseParallelmicBMagPhase": dict(
        input_channels=16, domain="DQ", domain_classifier="R", cnn_filters=[192, 192, 192], G=384, U=384,
        V=[384, 384], fc_layers=[128], parallel_ConvTC_block="2Parallel", parallel_magphase=True,
        extra_name="_micAMagPhaseParallelmicBMagPhase", batch_size=4, conv_train_gflop=1141.8),
    # config/SERVER_SELD-TCN-S1-PHI_8ch.txt: the real-valued baseline (nn.Conv* layers: none of the Q / DQ kernels)
    "SELD-TCN-S1-PHI_8ch": dict(input_channels=8, domain="R", domain_classifier="R", cnn_filters=[64, 64, 64],
                                G=128, U=128, V=[128, 128], fc_layers=[128], parallel_ConvTC_block="False",
                                parallel_magphase=False, extra_name="_8ch", batch_size=4, conv_train_gflop=99.7),
}
TIME_DIM, N_FRAMES_OUT, N_SED = 4800, 600, 42


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tflops=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, tflops=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def _nvml(self):
        """NVML handle for fast polling (a timed region of a few steps lasts tens of milliseconds; spawning
        nvidia-smi takes longer than that), or None."""
        try:
            import pynvml
            pynvml.nvmlInit()
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)
        except Exception:
            return None

    def run(self):
        nv = self._nvml()
        if nv is not None:
            pynvml, h = nv
            bits = ((0x8, 3), (0x40, 4), (0x20, 5), (0x4, 6))     # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap
            while not self._halt.is_set():
                try:
                    row = [str(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                           str(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)),
                           str(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3), "", "", "", ""]
                    try:
                        mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    except Exception:
                        mask = 0
                    for bit, col in bits:
                        row[col] = "Active" if mask & bit else "Not Active"
                    self.rows.append(row)
                except Exception:
                    pass
                self._halt.wait(0.01)
            return
        while not self._halt.is_set():
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                               "--format=csv,noheader,nounits"], timeout=5).decode().strip()
                self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(self.rows))


def synth_batch(pkg, cfg, batch, seed, device):
    """Synthetic L3DAS21-shaped batch: 0.1*randn waveforms (B, 8, 60 s @ 32 kHz) ->
    STFT magnitude(/phase) features via this repo's front-end kernel -> train.py-style mean/std
    normalisation; targets: SED Bernoulli(0.05), DOA U(-1,1) masked by SED."""
    g = torch.Generator().manual_seed(seed)
    phase = cfg["input_channels"] == 16
    wav = (0.1 * torch.randn(batch, 8, 1_920_000, generator=g)).to(device)
    feat = pkg.stft_magphase(wav, 512, 112, True, phase, True)            # (B, 8|16, 256, 4800)
    feat[:, :8] = (feat[:, :8] - feat[:, :8].mean()) / feat[:, :8].std()   # train.py:379-382
    if phase:
        feat[:, 8:] = (feat[:, 8:] - feat[:, 8:].mean()) / feat[:, 8:].std()   # train.py:397-400
    sed = (torch.rand(batch, N_FRAMES_OUT, N_SED, generator=g) < 0.05).float()
    doa = (2 * torch.rand(batch, N_FRAMES_OUT, 3 * N_SED, generator=g) - 1) * sed.repeat_interleave(3, -1)
    return feat.contiguous(), torch.cat([sed, doa], -1).to(device)


def frontend_roofline(pkg, device, peaks, batches=(1, 4, 16), iters=10):
    """STFT front end (utility_functions.py:129-155 -> csrc/stft_pair.cuh / csrc/stft.cuh), kernel-level: clips/s and achieved HBM GB/s
    against the measured copy bandwidth, for magnitude-only and magnitude + phase output, over a batch sweep.
    Algorithmic bytes per clip: 61.44 MB read + 39.32 MB (mag) [+ 39.32 MB (phase)] written.  Every
    iteration reads a different resident clip set (inputs of >= 61 MB per clip rotate through > 126 MB of L2)."""
    out = {"bound": "hbm", "unit": "GB/s", "peak": peaks["hbm_gbs"], "peak_source": peaks["source"], "sweep": []}
    g = torch.Generator().manual_seed(7)
    nb = max(batches)
    wav = (0.1 * torch.randn(nb + 2, 8, 1_920_000, generator=g)).to(device)
    for phase in (False, True):
        bytes_clip = 8 * 1_920_000 * 4 + (2 if phase else 1) * 8 * 256 * 4800 * 4
        for b in batches:
            for _ in range(2):
                pkg.stft_magphase(wav[:b], 512, 112, True, phase, True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for i in range(iters):
                o = i % 3 if b + 2 <= wav.shape[0] else 0
                pkg.stft_magphase(wav[o:o + b], 512, 112, True, phase, True)
            e1.record()
            torch.cuda.synchronize()
            us = 1e3 * e0.elapsed_time(e1) / iters
            gbs = b * bytes_clip / (us * 1e-6) / 1e9
            out["sweep"].append(dict(batch=b, phase=phase, us_per_launch=us, clips_per_s=b / (us * 1e-6),
                                     achieved=gbs, frac=gbs / peaks["hbm_gbs"]))
    head = next(r for r in out["sweep"] if r["batch"] == 1 and not r["phase"])
    best = max((r for r in out["sweep"] if not r["phase"]), key=lambda r: r["frac"])
    out.update(kernel="stft_pair_kernel (magnitude; csrc/stft_pair.cuh), stft_magphase_kernel (magnitude + phase; csrc/stft.cuh)",
               achieved=head["achieved"], frac=head["frac"],
               workload="8 ch x 60 s @ 32 kHz clip, nperseg 512, hop 400, magnitude only, batch 1",
               best=dict(batch=best["batch"], achieved=best["achieved"], frac=best["frac"]))
    del wav
    return out


def model_kwargs(cfg):
    kw = dict(COMMON)
    kw.update({k: v for k, v in cfg.items() if k not in ("batch_size", "conv_train_gflop")})
    return kw


def _cpu_train_step(model, opt, x, target):
    """The step body of train.py:546-561 with seld_loss of train.py:186-204 (BCE + 5 * MSE)."""
    import torch.nn.functional as tF
    opt.zero_grad()
    sed, doa = model(x)
    loss = (tF.binary_cross_entropy(torch.flatten(sed, 1), torch.flatten(target[:, :, :N_SED], 1))
            + 5.0 * tF.mse_loss(torch.flatten(doa, 1), torch.flatten(target[:, :, N_SED:], 1)))
    loss.backward()
    opt.step()
    return loss


def run_reference(args, cfg, rank, world):
    """The reference's own CPU implementation for the same config on the host cores of this box: the UNMODIFIED
    model.SELD_Model from the shipped copy of the reference (oracle/_ref, made by oracle/fetch_ref.sh; kind
    "reference"), or -- where that copy is absent -- the port of its arithmetic (oracle/cpu_model.py: torch.cat
    expansion + F.conv / mm, stock BatchNorm etc.; kind "port").  Rank 0 only."""
    if rank != 0:
        return
    from oracle import ref_import
    torch.set_num_threads(os.cpu_count())
    np.random.seed(1)
    torch.manual_seed(1)
    batch = args.batch or cfg["batch_size"]
    if ref_import.available() and not os.environ.get("SELDQ_REF_PORT"):
        kind = "reference"
        model = ref_import.load().model.SELD_Model(time_dim=TIME_DIM, **model_kwargs(cfg)).train()
        note = ("the reference's unmodified model.SELD_Model (%s), one replica on the host cores"
                % os.path.relpath(ref_import.REFERENCE_ROOT, ROOT))
    else:
        from oracle import cpu_model
        kind = "port"
        model = cpu_model.build_model(time_dim=TIME_DIM, **model_kwargs(cfg)).train()
        note = "CPU port of the reference arithmetic, one replica on the host cores"
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    g = torch.Generator().manual_seed(1234)
    # features of the same distribution as the GPU arm's (standardised STFT magnitudes): N(0,1) surrogate
    x = torch.randn(batch, cfg["input_channels"], 256, TIME_DIM, generator=g)
    sed = (torch.rand(batch, N_FRAMES_OUT, N_SED, generator=g) < 0.05).float()
    doa = (2 * torch.rand(batch, N_FRAMES_OUT, 3 * N_SED, generator=g) - 1) * sed.repeat_interleave(3, -1)
    target = torch.cat([sed, doa], -1)
    steps, warm = max(1, min(args.steps, args.ref_max_steps)), max(1, min(args.warmup, 1))
    for _ in range(warm):
        _cpu_train_step(model, opt, x, target)
    t0 = time.perf_counter()
    for _ in range(steps):
        _cpu_train_step(model, opt, x, target)
    dt = time.perf_counter() - t0
    value = steps * batch / dt
    line = dict(metric="train_samples_per_sec", value=value, unit="samples/s", n_gpus=args.gpus, steps=steps,
                warmup=warm, ms_per_step=1e3 * dt / steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=args.config, per_gpu_batch=batch, global_batch=batch,
                            note=note),
                cpu_baseline=dict(value=value, unit="samples/s", cores=os.cpu_count(), kind=kind,
                                  sample="%d warm-up + %d timed full training steps, batch %d" % (warm, steps, batch)),
                e2e=dict(value=value, unit="samples/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="DQSELD-TCN-S1-PHI_8ch", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the config file's batch_size)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--pool", type=int, default=4, help="distinct synthetic batches rotated through")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="drive every step from Python instead of one CUDA graph")
    ap.add_argument("--ref-max-steps", type=int, default=3)
    ap.add_argument("--model", default="mirror", choices=["mirror", "reference"],
                    help="mirror: this repository's seld_model.SELD_Model.  reference: the reference's UNMODIFIED "
                         "model.py (shipped copy oracle/_ref) imported on the drop-in layer modules, fused glue wired "
                         "in by install_dropin(fuse_model=True)")
    ap.add_argument("--no-fuse-model", action="store_true",
                    help="with --model reference: layers only, model.py's own forward code drives them one by one")
    ap.add_argument("--no-frontend", action="store_true", help="skip the STFT front-end measurement")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    pkg = importlib.import_module(PKG)
    pkg.set_precision(args.precision)
    F = pkg.functional
    trainer_mod = importlib.import_module(PKG + ".trainer")
    batch = args.batch or cfg["batch_size"]
    warm = max(3, args.warmup)

    np.random.seed(1)
    torch.manual_seed(1)                                   # train.py:214-221
    if args.model == "reference":
        # the caller is the reference's own model.py (host code only: every kernel it launches is this repository's)
        from oracle import ref_import
        model_cls = ref_import.load_model_on_dropin(pkg, fuse_model=not args.no_fuse_model).SELD_Model
    else:
        model_cls = pkg.SELD_Model
    model = model_cls(time_dim=TIME_DIM, **model_kwargs(cfg)).to(device).train()
    trainer = trainer_mod.Trainer(model, lr=1e-4, n_sed=N_SED)
    trainer.broadcast_parameters()
    torch.manual_seed(100 + rank)                          # dropout masks differ per replica

    pool_dev = [synth_batch(pkg, cfg, batch, 1234 + rank * 100 + i, device) for i in range(args.pool)]
    pool_host = [(x.cpu().pin_memory(), t.cpu().pin_memory()) for x, t in pool_dev]
    h2d = pool_host[0][0].numel() * 4 + pool_host[0][1].numel() * 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps):
        barrier()
        sampler = ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step_fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), (sampler.stop() if sampler else None)

    use_graph = not args.no_graph
    if use_graph:
        trainer.capture(*pool_dev[0], warmup=warm)
        do_step = trainer.step_graph
    else:
        do_step = trainer.step

    def step_resident(i):
        x, t = pool_dev[i % args.pool]
        do_step(x, t)

    e2e_state = {"staged": False}

    def step_e2e(i):
        # every step copies its own batch from pinned host memory and reads its loss back; with the captured
        # step the copy of batch i+1 is started before step i is replayed, so that it runs on the copy engine
        # while the SMs work on step i (trainer.stage / step_graph_staged)
        if use_graph:
            if not e2e_state["staged"]:
                trainer.stage(*pool_host[i % args.pool])
            loss = trainer.step_graph_staged()
            trainer.stage(*pool_host[(i + 1) % args.pool])
            e2e_state["staged"] = True
        else:
            hx, ht = pool_host[i % args.pool]
            loss = do_step(hx.to(device, non_blocking=True), ht.to(device, non_blocking=True))
        return float(loss.item())                          # device -> host read of the step's result

    for i in range(warm):
        step_resident(i)
    ms, clocks = timed(step_resident, args.steps)
    value = world * batch * args.steps / (ms / 1e3)

    for i in range(2):
        step_e2e(i)
    ms_e2e, _ = timed(step_e2e, args.steps)
    e2e_value = world * batch * args.steps / (ms_e2e / 1e3)

    # per-kernel durations: one eager step whose launches are queued behind a GPU-side sleep, so the
    # CUDA events around each of our kernels are not stretched by host launch latency
    # (one untimed eager step first: after the graph replays the caching allocator's default pool and the packed-weight
    # cache are cold, and a cudaMalloc in the measured step would synchronise with the sleep and leave the launches
    # behind it host-bound -- seen once as a backward pass "35 % slower" than in every other run)
    trainer.step(*pool_dev[0])
    torch.cuda.synchronize()
    F.profile_reset(enable=True)
    torch.cuda._sleep(int(2.5e9))
    trainer.step(*pool_dev[0])
    prof = F.profile_collect()
    F.profile_reset(enable=False)
    prof["launches"] *= args.steps                         # the same kernels run in each timed step

    frontend = None
    if rank == 0 and not args.no_frontend:
        try:
            frontend = frontend_roofline(pkg, device, load_peaks())
        except Exception as e:       # reported next to the headline, never required for it
            frontend = dict(error=repr(e))
    if rank == 0:
        peaks = load_peaks()
        k = prof["kernels"].get("qconv_cl_fprop_kernel" if args.precision == "bf16" else "conv_simt_kernel", None)
        roof = None
        if k and k["ms"] > 0:
            ach = k["flop"] / (k["ms"] / 1e3) / 1e12
            # DRAM bytes per launch of the same kernel from the committed Nsight Compute pass (profiles/, made by
            # tools/profile_trip.sh + tools/ncu_summary.py traffic): a recorded measurement, not taken in this run
            traffic, traffic_src = None, None
            tpath = os.path.join(ROOT, "profiles", "fprop_traffic.json")
            if os.path.exists(tpath):
                tj = json.load(open(tpath))
                if tj.get("kernel") == k["name"]:
                    traffic, traffic_src = tj["traffic_bytes_per_launch"], "profiles/fprop_traffic.json (ncu dram__bytes_read + write)"
            roof = dict(bound="tensor", kernel=k["name"], achieved=ach, peak=peaks["tflops"], unit="TFLOP/s",
                        frac=ach / peaks["tflops"], traffic=traffic, traffic_source=traffic_src, peak_source=peaks["source"],
                        launches_per_step=k["launches"], avg_launch_us=1e3 * k["ms"] / max(1, k["launches"]),
                        share_of_step=k["ms"] / (ms / args.steps),
                        other_kernels_ms_per_step={n: round(v["ms"], 3) for n, v in prof["kernels"].items()})
        line = dict(metric="train_samples_per_sec", value=value, unit="samples/s", n_gpus=world, steps=args.steps,
                    warmup=warm, ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype=args.precision, data="synthetic",
                    config=dict(workload=args.config, per_gpu_batch=batch, global_batch=batch * world,
                                time_frames=TIME_DIM, parallelism="dp%d" % world, cuda_graph=use_graph,
                                caller={"mirror": "seld_model.py (this repository's assembly)",
                                        "reference": "reference model.py, unmodified, on install_dropin()"}[args.model]
                                + (" layers only" if args.model == "reference" and args.no_fuse_model else ""),
                                l2="activations per step (>1 GB) exceed the 126 MB L2; %d distinct input batches "
                                   "are rotated" % args.pool),
                    clocks=clocks,
                    e2e=dict(value=e2e_value, unit="samples/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=4,
                             ms_per_step=ms_e2e / args.steps),
                    gpu_launches=prof["launches"], roofline=roof, roofline_frontend=frontend,
                    conv_tflops_per_gpu=cfg["conv_train_gflop"] * value / world / 1e3,
                    conv_frac_of_peak=cfg["conv_train_gflop"] * value / world / 1e3 / peaks["tflops"])
        if world == 1 and not args.no_cpu_baseline:
            try:
                out = subprocess.check_output([sys.executable, os.path.abspath(__file__), "--impl", "reference",
                                               "--config", args.config, "--steps", "2", "--warmup", "1",
                                               "--batch", str(batch)], timeout=900).decode().strip().splitlines()[-1]
                line["cpu_baseline"] = json.loads(out)["cpu_baseline"]
            except Exception as e:      # the baseline is reported, never required
                line["cpu_baseline"] = dict(error=repr(e))
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear down without destroy_process_group(): destroying the NCCL communicator while the captured training
        # graph (which holds the all-reduce) is alive can block forever.  All ranks are done and synchronised here.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
