"""Pins the CPU port used for the cpu_baseline / --impl reference timings (oracle/cpu_model.py:
the product's model assembly + the reference's torch.cat / F.conv arithmetic) against the
whole-model fixtures minted from the real reference.  float64 on CPU, so tolerances are tight."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import algebra as A
from oracle import cpu_model


@pytest.mark.parametrize("name", ["model_dq_tiny", "model_q_tiny", "model_dq_2branch_tiny"])
def test_cpu_port_reproduces_reference_model(name):
    meta, d = load_golden(name)
    cfg = dict(meta["cfg"])
    m = cpu_model.build_model(time_dim=meta["time_dim"], spatial_dropout_rate=0, dropout_perc=0, **cfg)
    m.load_state_dict({k[6:]: torch.from_numpy(np.asarray(v)) for k, v in d.items() if k.startswith("param/")})
    m = m.double().train()
    x = torch.from_numpy(d["x"].astype(np.float64))
    target = torch.from_numpy(d["target"].astype(np.float64))
    sed, doa = m(x)
    loss = cpu_model.seld_loss(sed, doa, target)
    loss.backward()
    assert A.rel_err(sed.detach().numpy(), d["sed"]) < 1e-9
    assert A.rel_err(doa.detach().numpy(), d["doa"]) < 1e-9
    assert abs(loss.item() - float(d["loss"])) < 1e-10
    n = 0
    for k, p in m.named_parameters():
        if ("grad/" + k) in d:
            assert A.rel_err(p.grad.numpy(), d["grad/" + k]) < 1e-5, k   # fixtures store grads as float32
            n += 1
        else:
            assert p.grad is None
    assert n == meta["n_grads"]
