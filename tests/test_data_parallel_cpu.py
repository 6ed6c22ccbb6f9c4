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
    # the exchange ran in two parts: everything but the CNN front from the TC_Block's gradient hook, the rest at the end
    assert tr._early_launched and 0 < tr.bucket.n_late < tr.bucket.flat.numel()
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


def test_optimizer_state_in_the_reference_checkpoint_layout():
    """Trainer.optimizer_state_dict(): one Adam entry per parameter in model.parameters() order, the layout the
    reference's save_model / load_model write (train.py:26-81); loading it into a fresh trainer continues identically."""
    trainer_mod = importlib.import_module(PKG + ".trainer")
    meta, d = load_golden("model_dq_tiny")
    xs, ts = _shards(d, 1)
    m = _build(meta, d)
    tr = trainer_mod.Trainer(m, lr=1e-3, n_sed=42)
    tr.step(xs[0], ts[0])
    sd = tr.optimizer_state_dict()
    params = list(m.parameters())
    assert sd["param_groups"][0]["params"] == list(range(len(params))) and len(sd["state"]) == len(params)
    for i, p in enumerate(params):
        assert sd["state"][i]["exp_avg"].shape == p.shape
    # the same numbers a per-parameter Adam produces
    m2 = _build(meta, d)
    opt = torch.optim.Adam(m2.parameters(), lr=1e-3)
    from oracle import cpu_model
    cpu_model.train_step(m2, opt, xs[0], ts[0])
    for i, p in enumerate(m2.parameters()):
        if p.grad is None:
            continue
        assert torch.allclose(sd["state"][i]["exp_avg"], opt.state[p]["exp_avg"], rtol=1e-9, atol=1e-14)
        assert torch.allclose(sd["state"][i]["exp_avg_sq"], opt.state[p]["exp_avg_sq"], rtol=1e-9, atol=1e-18)
    # round trip: a fresh trainer that loads model + optimiser state takes the same second step
    m3 = _build(meta, d)
    m3.load_state_dict(m.state_dict())
    tr3 = trainer_mod.Trainer(m3, lr=1e-3, n_sed=42)
    tr3.load_optimizer_state_dict(sd)
    tr.step(xs[0], ts[0])
    tr3.step(xs[0], ts[0])
    for a, b in zip(m.parameters(), m3.parameters()):
        assert torch.allclose(a, b, rtol=1e-10, atol=1e-14)


def _sum_exchange_main(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    trainer_mod = importlib.import_module(PKG + ".trainer")
    trainer_mod._AR_SUM = True                   # SELDQ_AR_SUM=1: sum + scaling pass (gloo has no average either)
    t = torch.arange(10, dtype=torch.float32) * (rank + 1)
    trainer_mod._nccl_mean_(t[2:], None)         # a slice of the bucket, as the two-part exchange passes it
    torch.save(t, os.path.join(out_dir, "sum%d.pt" % rank))
    dist.destroy_process_group()


def test_sum_exchange_switch_averages_a_bucket_slice(tmp_path):
    """trainer._nccl_mean_ with SELDQ_AR_SUM=1 (ncclSum + scaling instead of ncclAvg): in place on a view of the bucket,
    mean over the ranks, the rest of the bucket untouched."""
    world = 2
    mp.spawn(_sum_exchange_main, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    base = torch.arange(10, dtype=torch.float32)
    for r in range(world):
        t = torch.load(os.path.join(str(tmp_path), "sum%d.pt" % r))
        assert torch.equal(t[:2], base[:2] * (r + 1))
        assert torch.allclose(t[2:], base[2:] * 1.5)
