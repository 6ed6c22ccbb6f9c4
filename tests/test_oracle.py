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


def _random_seld_case(rng, frames, p_on, jitter):
    """Targets with up to three simultaneous events per class; predictions = targets with flips and jittered locations."""
    n_sed = 42
    t_sed = (rng.random((frames, n_sed)) < p_on).astype(np.float64)
    t_doa = (2 * rng.random((frames, n_sed * 3)) - 1) * np.repeat(t_sed, 3, axis=-1)
    sed = np.clip(t_sed * 0.9 + 0.05 + 0.6 * (rng.random((frames, n_sed)) < 0.03) * (1 - 2 * t_sed), 0, 1)
    doa = np.clip(t_doa + jitter * rng.standard_normal(t_doa.shape), -1, 1)
    return sed, doa, np.concatenate([t_sed, t_doa], axis=-1)


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference is only present in the build container")
@pytest.mark.parametrize("seed,frames,p_on,jitter", [(0, 600, 0.05, 0.05), (1, 600, 0.3, 0.5), (2, 95, 0.1, 0.2), (3, 40, 0.6, 1.0),
                                                     (4, 20, 0.0, 0.1)])
def test_seld_metric_restatement_matches_live_reference(seed, frames, p_on, jitter):
    """oracle/seld_metrics.py against the imported utility_functions.gen_submission_list_task2 +
    Dcase21_metrics.segment_labels + SELDMetrics (the loop of train.py:84-130): the four scores bit-identical."""
    from oracle import seld_metrics as M
    ns = ref_import.load()
    rng = np.random.default_rng(seed)
    ref = ns.dcase.SELDMetrics(nb_classes=14, doa_threshold=20)
    mine = M.SeldScores(20, 14)
    for clip in range(3):
        sed, doa, target = _random_seld_case(rng, frames, p_on, jitter)
        _, pd = ns.uf.gen_submission_list_task2(sed, doa, max_overlaps=3, max_loc_value=2.0)
        _, td = ns.uf.gen_submission_list_task2(target[:, :42], target[:, 42:], max_overlaps=3, max_loc_value=2.0)
        assert pd == M.events_per_frame(sed, doa) and td == M.events_per_frame(target[:, :42], target[:, 42:])
        pl, tl = ns.dcase.segment_labels(pd, frames), ns.dcase.segment_labels(td, frames)
        assert pl == M.blocks(pd, frames) and tl == M.blocks(td, frames)
        ref.update_seld_scores(pl, tl)
        mine.update(M.blocks(M.events_per_frame(sed, doa), frames), M.blocks(td, frames))
    assert tuple(float(v) for v in ref.compute_seld_scores()) == tuple(float(v) for v in mine.scores())


@pytest.mark.parametrize("name", golden_names("model"))
def test_seld_metric_restatement_matches_golden_scores(name):
    """The scores the REFERENCE's code gave for the reference's own outputs of each model fixture (oracle/make_golden.py)."""
    from oracle import seld_metrics as M
    meta, d = load_golden(name)
    got = M.seld_scores(d["sed"], d["doa"], d["target"], num_frames=d["sed"].shape[1])
    assert got == tuple(float(v) for v in d["seld_scores"])


@pytest.mark.parametrize("name", golden_names("convT"))
def test_oracle_transposed_conv_matches_reference_fixture(name):
    """quaternion_transpose_conv (quaternion_ops.py:149-172) = the convolution's input-gradient pass."""
    meta, d = load_golden(name)
    ws = [d["w%d" % i].astype(np.float64) for i in range(4)]
    y = A.qconv_transpose(d["x"], ws, d["b"] if meta["bias"] else None, meta["padding"], meta["dilation"], meta["stride"],
                          meta.get("output_padding", 0))
    assert A.rel_err(y, d["y"]) < 1e-12


# ---- rotation variants, hamilton_product, q_normalize, quaternion_exp --------------------------------
@pytest.mark.parametrize("name", golden_names("rot_conv") + golden_names("rot_convT") + golden_names("rot_linear"))
def test_oracle_rotation_variants_match_reference_fixture(name):
    """quaternion_conv_rotation / quaternion_transpose_conv_rotation / quaternion_linear_rotation
    (quaternion_ops.py:174-232, :235-295, :330-388): outputs, input gradients and -- through the hand-derived gradient of
    the rotation weight -- the compact weight gradients of the reference's autograd; the float32 weight to the last bit
    or the one before it (torch's vectorised CPU square root is not correctly rounded, numpy's is)."""
    meta, d = load_golden(name)
    ws = [d["w%d" % i].astype(np.float64) for i in range(4)]
    qf, kind = meta["quaternion_format"], meta["kind"]
    b = d["b"] if meta["bias"] else None
    W32 = A.rotation_weight(ws, qf, np.float32)
    assert W32.dtype == np.float32
    assert np.abs(W32.reshape(d["W32"].shape).astype(np.float64) - d["W32"]).max() <= 1.5e-7 * np.abs(d["W32"]).max()
    if kind == "rot_conv":
        y = A.qconv_rotation(d["x"], ws, b, meta["stride"], meta["padding"], meta["dilation"], qf)
        gx, gws, gb = A.qconv_rotation_backward(d["x"], ws, d["gy"], meta["stride"], meta["padding"], meta["dilation"], qf)
        assert A.rel_err(gx, d["gx"]) < 1e-12
        for i in range(4):
            assert A.rel_err(gws[i], d["gw%d" % i]) < 1e-12, i
        if b is not None:
            assert A.rel_err(gb, d["gb"]) < 1e-12
    elif kind == "rot_convT":
        y = A.qconv_transpose_rotation(d["x"], ws, b, meta["padding"], meta["dilation"], qf, meta["stride"],
                                       meta.get("output_padding", 0))
    else:
        y = A.qlinear_rotation(d["x"], ws, b, qf)
        W = A.rotation_weight(ws, qf)
        x2, g2 = d["x"].reshape(-1, W.shape[0]), d["gy"].reshape(-1, W.shape[1])
        assert A.rel_err((g2 @ W.T).reshape(d["x"].shape), d["gx"]) < 1e-12
        gws = A.rotation_weight_backward(ws, x2.T @ g2, qf)
        for i in range(4):
            assert A.rel_err(gws[i], d["gw%d" % i]) < 1e-12, i
    assert A.rel_err(y, d["y"]) < 1e-12


@pytest.mark.parametrize("name", golden_names("qpointwise"))
def test_oracle_quaternion_pointwise_operators_match_reference_fixture(name):
    meta, d = load_golden(name)
    assert A.rel_err(A.hamilton_product(d["a"], d["b"]), d["ham"]) < 1e-13
    assert A.rel_err(A.q_normalize(d["a"]), d["norm"]) < 1e-13
    assert A.rel_err(A.quaternion_exp(d["a"]), d["exp"]) < 1e-13
    # the gradients of a Hamilton product are Hamilton products with a conjugate
    conj = lambda q: np.concatenate([c if k == 0 else -c for k, c in enumerate(A._components(q))], axis=1)
    assert A.rel_err(A.hamilton_product(d["g"], conj(d["b"])), d["ham_ga"]) < 1e-13
    assert A.rel_err(A.hamilton_product(conj(d["a"]), d["g"]), d["ham_gb"]) < 1e-13


@pytest.mark.reference
@pytest.mark.skipif(not ref_import.available(), reason="reference tree only exists in the build container")
def test_oracle_rotation_variants_match_live_reference():
    import torch
    ns = ref_import.load()
    rng = np.random.default_rng(8)
    for qf in (False, True):
        nc = 4 if qf else 3
        ws = [0.5 * rng.standard_normal((3, 2, 3)) for _ in range(4)]
        x = rng.standard_normal((2, nc * 2, 17))
        b = rng.standard_normal(nc * 3)
        tw = [torch.tensor(w) for w in ws]
        yr = ns.q_ops.quaternion_conv_rotation(torch.tensor(x), *tw, torch.tensor(b), 1, 2, 1, 2, qf).numpy()
        assert A.rel_err(A.qconv_rotation(x, ws, b, 1, 2, 2, qf), yr) < 1e-13
        xt = rng.standard_normal((2, nc * 3, 17))
        yr = ns.q_ops.quaternion_transpose_conv_rotation(torch.tensor(xt), *tw, None, 1, 1, 0, 1, 1, qf).numpy()
        assert A.rel_err(A.qconv_transpose_rotation(xt, ws, None, 1, 1, qf), yr) < 1e-13
        wl = [0.5 * rng.standard_normal((4, 3)) for _ in range(4)]
        xl = rng.standard_normal((5, nc * 4))
        yr = ns.q_ops.quaternion_linear_rotation(torch.tensor(xl), *[torch.tensor(w) for w in wl], None, qf).numpy()
        assert A.rel_err(A.qlinear_rotation(xl, wl, None, qf), yr) < 1e-13
