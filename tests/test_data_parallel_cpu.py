"""Batch-sharded data-parallel step on CPU: 2 ranks over gloo (the N > 1 host logic of trainer.py).

The Q / DQ layers themselves have no CPU path, so the model is assembled with the oracle's CPU layer
classes (oracle/cpu_model.py, allowed in tests); what is under test is the trainer: flat gradient
bucket, one all-reduce per step, division by the world size, identical replicas after the step.
Correctness oracle: averaged gradients == mean of the gradients of the shards run one
after the other in a single process."""
import importlib
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, load_golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _build(meta, d):
    from oracle import cpu_model
    cfg = dict(meta["cfg"])
    m = cpu_model.build_model(time_dim=meta["time_dim"], spatial_dropout_rate=0, dropout_perc=0, **cfg)
    m.load_state_dict({k[6:]: torch.from_numpy(np.asarray(v)) for k, v in d.items() if k.startswith("param/")})
    return m.double().train()


def _shards(d, world):
    x = torch.from_numpy(np.asarray(d["x"], np.float64))
    t = torch.from_numpy(np.asarray(d["target"], np.float64))
    g = torch.Generator().manual_seed(5)
    xs = [x] + [x + 0.1 * torch.randn(x.shape, generator=g, dtype=torch.float64) for _ in range(world - 1)]
    return xs, [t] * world


def _rank_main(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    trainer_mod = importlib.import_module(PKG + ".trainer")
    meta, d = load_golden("model_dq_tiny")
    m = _build(meta, d)
    if rank != 0:                      # replicas start different: broadcast_parameters must fix that
        with torch.no_grad():
            for p in m.parameters():
                p.add_(1.0)
    tr = trainer_mod.Trainer(m, lr=1e-3, n_sed=42)
    tr.broadcast_parameters(src=0)
    xs, ts = _shards(d, world)
    tr.step(xs[rank], ts[rank])
    torch.save({"flat": tr.bucket.flat.clone(), "params": [p.detach().clone() for p in m.parameters()]},
               os.path.join(out_dir, "rank%d.pt" % rank))
    dist.destroy_process_group()


def test_two_rank_gloo_step_matches_sequential_mean(tmp_path):
    world = 2
    mp.spawn(_rank_main, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(str(tmp_path), "rank%d.pt" % r)) for r in range(world)]
    # the same shards, one after the other, in this process
    trainer_mod = importlib.import_module(PKG + ".trainer")
    meta, d = load_golden("model_dq_tiny")
    xs, ts = _shards(d, world)
    flats = []
    for r in range(world):
        m = _build(meta, d)
        tr = trainer_mod.Trainer(m, lr=0.0, n_sed=42)
        tr.step(xs[r], ts[r])
        flats.append(tr.bucket.flat.clone())
    mean = sum(flats) / world
    scale = float(mean.abs().max())
    for r in range(world):
        assert float((outs[r]["flat"] - mean).abs().max()) < 1e-10 * max(1.0, scale)
    # replicas hold identical parameters after the step, and they moved
    for p0, p1 in zip(outs[0]["params"], outs[1]["params"]):
        assert torch.equal(p0, p1)
    ref = _build(meta, d)
    moved = sum(float((p0 - q.detach()).abs().max()) > 0 for p0, q in zip(outs[0]["params"], ref.parameters()))
    assert moved > 0
