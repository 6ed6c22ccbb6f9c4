"""Pins the oracle (oracle/algebra.py): against every committed golden fixture minted from the
reference (always), and against the live reference where /root/reference exists."""
import numpy as np
import pytest

from conftest import golden_names, load_golden
from oracle import algebra as A
from oracle import ref_import

NW = {"Q": 4, "DQ": 8}


def _tol(*arrays):
    # fixtures above 20k elements are stored as float32 (oracle/make_golden.py:_save)
    return 1e-12 if all(a.dtype == np.float64 for a in arrays) else 2e-6


@pytest.mark.parametrize("name", golden_names("conv"))
def test_conv_oracle_matches_golden(name):
    meta, d = load_golden(name)
    alg = meta["algebra"]
    ws = [d["w%d" % i].astype(np.float64) for i in range(NW[alg])]
    b = d["b"].astype(np.float64) if meta["bias"] else None
    x, gy = d["x"].astype(np.float64), d["gy"].astype(np.float64)
    y = A.qconv(x, ws, b, meta["stride"], meta["padding"], meta["dilation"], alg)
    assert A.rel_err(y, d["y"]) < _tol(d["y"])
    gx, gws, gb = A.qconv_backward(x, ws, gy, meta["stride"], meta["padding"], meta["dilation"], alg)
    assert A.rel_err(gx, d["gx"]) < _tol(d["gx"], d["gy"])
    for i in range(NW[alg]):
        assert A.rel_err(gws[i], d["gw%d" % i]) < _tol(d["gy"], d["x"])
    if meta["bias"]:
        assert A.rel_err(gb, d["gb"]) < _tol(d["gy"])


@pytest.mark.parametrize("name", golden_names("linear"))
def test_linear_oracle_matches_golden(name):
    meta, d = load_golden(name)
    alg = meta["algebra"]
    oalg = "Q" if alg == "Q" else "DQ_LINEAR"
    ws = [d["w%d" % i].astype(np.float64) for i in range(NW[alg])]
    b = d["b"].astype(np.float64) if meta["bias"] else None
    y = A.qlinear(d["x"], ws, b, oalg)
    assert A.rel_err(y, d["y"]) < 1e-12
    gx, gws, gb = A.qlinear_backward(d["x"], ws, d["gy"], oalg)
    assert A.rel_err(gx, d["gx"]) < 1e-12
    for i in range(NW[alg]):
        assert A.rel_err(gws[i], d["gw%d" % i]) < 1e-12
    if meta["bias"]:
        assert A.rel_err(gb, d["gb"]) < 1e-12


@pytest.mark.parametrize("name", golden_names("stft"))
def test_stft_oracle_matches_golden(name):
    meta, d = load_golden(name)
    out = A.spectrum_fast(d["x"].astype(np.float64), nperseg=meta["nperseg"], noverlap=meta["noverlap"],
                          output_phase=meta["output_phase"])
    assert out.shape == d["out"].shape
    C = d["x"].shape[0]
    assert A.rel_err(out[:C], d["out"][:C]) < 2e-6
    if meta["output_phase"]:
        dphi = np.angle(np.exp(1j * (out[C:] - d["out"][C:])))
        assert np.abs(dphi).max() < 1e-4


def test_block_tables_structure():
    # comp = a XOR b; DQ expanded weight = [[Q(w),0],[Q(w2),Q(w)]] (25 % zeros)
    widx, sign = A.block_table("Q")
    assert (widx == (np.arange(4)[:, None] ^ np.arange(4)[None, :])).all()
    assert sign.tolist() == [[1, -1, -1, -1], [1, 1, -1, 1], [1, 1, 1, -1], [1, -1, 1, 1]]
    widx, sign = A.block_table("DQ")
    assert (widx[:4, 4:] == -1).all() and (widx[:4, :4] == widx[4:, 4:]).all()
    assert (widx[4:, :4] == widx[:4, :4] + 4).all()
    wl, sl = A.block_table("DQ_LINEAR")
    assert (wl == widx.T).all() and (sl == sign.T).all()


def test_compact_grads_is_adjoint_of_expand():
    rng = np.random.default_rng(0)
    for alg, lin in (("Q", False), ("DQ", False), ("Q", True), ("DQ_LINEAR", True)):
        nw = 4 if alg == "Q" else 8
        shape = (3, 5) if lin else (5, 3, 2)
        ws = [rng.standard_normal(shape) for _ in range(nw)]
        W = A.expand_weight(ws, alg, linear=lin)
        G = rng.standard_normal(W.shape)
        O, I = (shape[1], shape[0]) if lin else (shape[0], shape[1])
        gs = A.compact_grads(G, alg, O, I, linear=lin)
        lhs = float((W * G).sum())
        rhs = float(sum((w * g).sum() for w, g in zip(ws, gs)))
        assert abs(lhs - rhs) < 1e-9 * max(1.0, abs(lhs))


@pytest.mark.reference
@pytest.mark.skipif(not ref_import.available(), reason="reference tree only exists in the build container")
def test_oracle_matches_live_reference():
    import torch
    ns = ref_import.load()
    rng = np.random.default_rng(5)
    for alg, nw in (("Q", 4), ("DQ", 8)):
        O, I = 3, 2
        ws = [rng.standard_normal((O, I, 3)) for _ in range(nw)]
        x = rng.standard_normal((2, nw * I, 19))
        b = rng.standard_normal(nw * O)
        fn = ns.q_ops.quaternion_conv if alg == "Q" else ns.dq_ops.dual_quaternion_conv
        yr = fn(torch.tensor(x), *[torch.tensor(w) for w in ws], torch.tensor(b), 1, 3, 1, 3).numpy()
        assert A.rel_err(A.qconv(x, ws, b, 1, 3, 3, alg), yr) < 1e-13
        wl = [rng.standard_normal((I, O)) for _ in range(nw)]
        xl = rng.standard_normal((5, nw * I))
        fl = ns.q_ops.quaternion_linear if alg == "Q" else ns.dq_ops.dual_quaternion_linear
        yr = fl(torch.tensor(xl), *[torch.tensor(w) for w in wl], None).numpy()
        assert A.rel_err(A.qlinear(xl, wl, None, "Q" if alg == "Q" else "DQ_LINEAR"), yr) < 1e-13
    x = 0.1 * rng.standard_normal((8, 32000))
    r = ns.uf.spectrum_fast(x, nperseg=512, noverlap=112, output_phase=True)
    o = A.spectrum_fast(x, nperseg=512, noverlap=112, output_phase=True)
    assert r.shape == o.shape == (16, 256, 80)
    assert np.abs(r[:8] - o[:8]).max() < 1e-14
