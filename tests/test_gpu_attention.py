"""Fused multi-head attention kernels (csrc/attention.cu) against the reference's arithmetic
(model.py:40-48: einsum nqhd,nkhd->nhqk, softmax(energy / sqrt(d)), einsum nhql,nlhd->nqhd) evaluated in float64.
bf16 operands, fp32 accumulation: rel 2e-2 (max-abs normalised), the tensor-core mode's tolerance."""
import numpy as np
import pytest
import torch

from oracle import algebra as A

pytestmark = pytest.mark.gpu


def _reference(q, k, v, heads, gout):
    """q, k, v (N, E, S) float64 leaf tensors -> out (N, S, E) and the gradients, the reference's three steps."""
    n, e, s = q.shape
    d = e // heads
    sp = lambda t: t.view(n, e, s).permute(0, 2, 1).reshape(n, s, heads, d)           # model.py:34-36
    energy = torch.einsum("nqhd,nkhd->nhqk", sp(q), sp(k))
    att = torch.softmax(energy / (d ** 0.5), dim=3)
    out = torch.einsum("nhql,nlhd->nqhd", att, sp(v)).reshape(n, s, e)
    out.backward(gout)
    return out.detach(), q.grad, k.grad, v.grad


@pytest.mark.parametrize("n,heads,d,s", [(2, 2, 48, 200), (1, 3, 16, 136), (1, 2, 32, 264), (2, 1, 32, 128),
                                         (1, 8, 48, 2400)])
def test_attention_matches_float64_reference(seldq, n, heads, d, s):
    g = torch.Generator().manual_seed(100 + s)
    e = heads * d
    q, k, v = (torch.randn(n, e, s, generator=g, dtype=torch.float64) * sc for sc in (1.5, 1.5, 1.0))
    gout = torch.randn(n, s, e, generator=g, dtype=torch.float64)
    ref = _reference(q.clone().requires_grad_(True), k.clone().requires_grad_(True), v.clone().requires_grad_(True), heads,
                     gout)
    qc, kc, vc = (t.float().cuda().requires_grad_(True) for t in (q, k, v))
    assert seldq.functional.attention_supported(qc, heads)
    out = seldq.attention(qc, kc, vc, heads)
    out.backward(gout.float().cuda())
    torch.cuda.synchronize()
    errs = dict(out=A.rel_err(out.detach().cpu().numpy(), ref[0].numpy()), dq=A.rel_err(qc.grad.cpu().numpy(), ref[1].numpy()),
                dk=A.rel_err(kc.grad.cpu().numpy(), ref[2].numpy()), dv=A.rel_err(vc.grad.cpu().numpy(), ref[3].numpy()))
    print((n, heads, d, s), errs)
    assert all(np.isfinite(v_) and v_ < 2e-2 for v_ in errs.values()), errs


def test_attention_declines_what_it_does_not_serve(seldq):
    q = torch.randn(1, 128, 64, device="cuda")
    assert not seldq.functional.attention_supported(q, 2)            # head_dim 64
    assert not seldq.functional.attention_supported(torch.randn(1, 96, 100, device="cuda"), 2)      # S % 8 != 0
    with seldq.precision("fp32"):
        assert not seldq.functional.attention_supported(torch.randn(1, 96, 64, device="cuda"), 2)   # fp32 parity mode


def test_attention_module_uses_the_fused_kernels_and_matches_the_pytorch_path(seldq):
    """MultiHeadAttention (model.py:12-51) of the mirror model: 'own' kernels in bf16 mode against the PyTorch ops of
    fp32 mode, same weights."""
    sm = __import__("importlib").import_module(seldq.__name__ + ".seld_model")
    torch.manual_seed(4)
    mha = sm.MultiHeadAttention(96, 2).cuda()
    x = torch.randn(2, 168, 96, device="cuda", requires_grad=True)
    gy = torch.randn(2, 168, 96, device="cuda")
    res = {}
    for prec in ("fp32", "bf16"):
        mha.zero_grad(set_to_none=True)
        x.grad = None
        with seldq.precision(prec):
            seldq.functional.profile_reset(True)
            y = mha(x, x, x)
            y.backward(gy)
            prof = seldq.functional.profile_collect()
            seldq.functional.profile_reset(False)
        assert ("attn_kernel" in prof["kernels"]) == (prec == "bf16")
        res[prec] = (y.detach().cpu().numpy(), x.grad.cpu().numpy(),
                     {k: p.grad.cpu().numpy() for k, p in mha.named_parameters()})
    assert A.rel_err(res["bf16"][0], res["fp32"][0]) < 2e-2
    assert A.rel_err(res["bf16"][1], res["fp32"][1]) < 2e-2
    for k in res["fp32"][2]:
        assert A.rel_err(res["bf16"][2][k], res["fp32"][2][k]) < 2e-2, k
