"""TEST INFRASTRUCTURE ONLY.  Mints tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python oracle/make_golden.py

Every fixture is computed by the reference's own functions in float64 on CPU
(quaternion_conv / dual_quaternion_conv / quaternion_linear / QuaternionLinearFunction /
dual_quaternion_linear / spectrum_fast / SELD_Model) with gradients from torch autograd.
Inputs are stored as float32-representable values so that the CUDA path (fp32 storage) sees
bit-identical inputs.  Seeds are recorded in each file's ``meta`` JSON.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def f32(a):
    return np.asarray(a, np.float32).astype(np.float64)


def _save(name, meta, d):
    """Large arrays are stored as float32 (6e-8 relative: far below every parity tolerance);
    small ones keep the reference's float64 so the oracle can be pinned to 1e-12."""
    d = {k: (np.asarray(v, np.float32) if np.asarray(v).size > 20000 else np.asarray(v)) for k, v in d.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(meta), **d)


def conv_case(ns, name, algebra, ndim, N, I, O, spatial, k, stride, padding, dilation, bias, seed):
    rng = np.random.default_rng(seed)
    nc = {"Q": 4, "DQ": 8}[algebra]
    kshape = (k,) * ndim
    ws = [f32(rng.standard_normal((O, I) + kshape) * 0.2) for _ in range(nc)]
    x = f32(rng.standard_normal((N, nc * I) + tuple(spatial)))
    b = f32(rng.standard_normal(nc * O)) if bias else None
    tx = torch.tensor(x, requires_grad=True)
    tw = [torch.tensor(w, requires_grad=True) for w in ws]
    tb = torch.tensor(b, requires_grad=True) if bias else None
    fn = ns.q_ops.quaternion_conv if algebra == "Q" else ns.dq_ops.dual_quaternion_conv
    y = fn(tx, *tw, tb, stride, padding, 1, dilation)
    gy = f32(rng.standard_normal(tuple(y.shape)))
    y.backward(torch.tensor(gy))
    d = dict(x=x, gy=gy, y=y.detach().numpy(), gx=tx.grad.numpy())
    for i, (w, t) in enumerate(zip(ws, tw)):
        d["w%d" % i] = w
        d["gw%d" % i] = t.grad.numpy()
    if bias:
        d["b"] = b
        d["gb"] = tb.grad.numpy()
    meta = dict(kind="conv", algebra=algebra, ndim=ndim, stride=stride, padding=padding,
                dilation=dilation, bias=bool(bias), seed=seed,
                source="quaternion_ops.py:125-147" if algebra == "Q" else "dual_quaternion_ops.py:111-153")
    _save(name, meta, d)
    print(name, tuple(x.shape), "->", tuple(y.shape))


def convT_case(ns, name, ndim, N, I, O, spatial, k, padding, dilation, bias, seed, stride=1, output_padding=0):
    """quaternion_transpose_conv (quaternion_ops.py:149-172): weights are (in / 4, out / 4, k...)."""
    rng = np.random.default_rng(seed)
    kshape = (k,) * ndim
    ws = [f32(rng.standard_normal((I, O) + kshape) * 0.2) for _ in range(4)]
    x = f32(rng.standard_normal((N, 4 * I) + tuple(spatial)))
    b = f32(rng.standard_normal(4 * O)) if bias else None
    tx = torch.tensor(x, requires_grad=True)
    tw = [torch.tensor(w, requires_grad=True) for w in ws]
    tb = torch.tensor(b, requires_grad=True) if bias else None
    y = ns.q_ops.quaternion_transpose_conv(tx, *tw, tb, stride, padding, output_padding, 1, dilation)
    gy = f32(rng.standard_normal(tuple(y.shape)))
    y.backward(torch.tensor(gy))
    d = dict(x=x, gy=gy, y=y.detach().numpy(), gx=tx.grad.numpy())
    for i, (w, t) in enumerate(zip(ws, tw)):
        d["w%d" % i] = w
        d["gw%d" % i] = t.grad.numpy()
    if bias:
        d["b"] = b
        d["gb"] = tb.grad.numpy()
    meta = dict(kind="convT", algebra="Q", ndim=ndim, stride=stride, padding=padding, dilation=dilation, bias=bool(bias),
                output_padding=output_padding, seed=seed, source="quaternion_ops.py:149-172")
    _save(name, meta, d)
    print(name, tuple(x.shape), "->", tuple(y.shape))


def linear_case(ns, name, algebra, rows, I, O, bias, seed, use_function=False):
    rng = np.random.default_rng(seed)
    nc = {"Q": 4, "DQ": 8}[algebra]
    ws = [f32(rng.standard_normal((I, O)) * 0.2) for _ in range(nc)]
    x = f32(rng.standard_normal((rows, nc * I)))
    b = f32(rng.standard_normal(nc * O)) if bias else None
    tx = torch.tensor(x, requires_grad=True)
    tw = [torch.tensor(w, requires_grad=True) for w in ws]
    tb = torch.tensor(b, requires_grad=True) if bias else None
    if algebra == "Q":
        y = (ns.q_ops.QuaternionLinearFunction.apply(tx, *tw, tb) if use_function
             else ns.q_ops.quaternion_linear(tx, *tw, tb))
        src = "quaternion_ops.py:392-464" if use_function else "quaternion_ops.py:299-327"
    else:
        y = ns.dq_ops.dual_quaternion_linear(tx, *tw, tb)
        src = "dual_quaternion_ops.py:156-203"
    gy = f32(rng.standard_normal(tuple(y.shape)))
    y.backward(torch.tensor(gy))
    d = dict(x=x, gy=gy, y=y.detach().numpy(), gx=tx.grad.numpy())
    for i, (w, t) in enumerate(zip(ws, tw)):
        d["w%d" % i] = w
        d["gw%d" % i] = t.grad.numpy()
    if bias:
        d["b"] = b
        d["gb"] = tb.grad.numpy()
    meta = dict(kind="linear", algebra=algebra, bias=bool(bias), seed=seed, source=src)
    _save(name, meta, d)
    print(name, tuple(x.shape), "->", tuple(y.shape))


def _rotation_weight32(ns, ws, qf):
    """The float32 weight the reference's rotation variants build, read back exactly: quaternion_linear_rotation of an
    identity input returns its global_rot_kernel (every product is 1 * w or 0 * w), applied to the compact tensors
    flattened to (d0, d1 * taps) -- the construction is element-wise."""
    w32 = [torch.tensor(np.asarray(w, np.float32)) for w in ws]
    d0, d1 = w32[0].shape[:2]
    taps = int(np.prod(w32[0].shape[2:])) if w32[0].dim() > 2 else 1
    nc = 4 if qf else 3
    flat = [w.reshape(d0, d1 * taps) for w in w32]
    G = ns.q_ops.quaternion_linear_rotation(torch.eye(nc * d0, dtype=torch.float32), *flat, None, qf)
    return G.reshape(nc * d0, nc, d1, taps).reshape(nc * d0, nc * d1, taps).numpy()


def rotation_case(ns, name, kind, ndim, N, I, O, spatial, k, padding, dilation, bias, qf, seed, stride=1, output_padding=0):
    """quaternion_conv_rotation (quaternion_ops.py:174-232), quaternion_transpose_conv_rotation (:235-295) and
    quaternion_linear_rotation (:330-388).  kind: 'conv' | 'convT' | 'linear'; I, O = compact sizes; for 'linear'
    `spatial` holds the leading dimensions of the input."""
    rng = np.random.default_rng(seed)
    nc = 4 if qf else 3
    if kind == "linear":
        wshape, xshape = (I, O), tuple(spatial) + (nc * I,)
    else:
        kshape = (k,) * ndim
        wshape = ((O, I) if kind == "conv" else (I, O)) + kshape
        xshape = (N, nc * I) + tuple(spatial)
    ws = [f32(rng.standard_normal(wshape) * 0.4) for _ in range(4)]
    x = f32(rng.standard_normal(xshape))
    b = f32(rng.standard_normal(nc * O)) if bias else None
    tx = torch.tensor(x, requires_grad=True)
    tw = [torch.tensor(w, requires_grad=True) for w in ws]
    tb = torch.tensor(b, requires_grad=True) if bias else None
    if kind == "conv":
        y = ns.q_ops.quaternion_conv_rotation(tx, *tw, tb, stride, padding, 1, dilation, qf)
        src = "quaternion_ops.py:174-232"
    elif kind == "convT":
        y = ns.q_ops.quaternion_transpose_conv_rotation(tx, *tw, tb, stride, padding, output_padding, 1, dilation, qf)
        src = "quaternion_ops.py:235-295"
    else:
        y = ns.q_ops.quaternion_linear_rotation(tx, *tw, tb, qf)
        src = "quaternion_ops.py:330-388"
    gy = f32(rng.standard_normal(tuple(y.shape)))
    y.backward(torch.tensor(gy))
    d = dict(x=x, gy=gy, y=y.detach().numpy(), gx=tx.grad.numpy(), W32=_rotation_weight32(ns, ws, qf))
    for i, (w, t) in enumerate(zip(ws, tw)):
        d["w%d" % i] = w
        d["gw%d" % i] = t.grad.numpy()
    if bias:
        d["b"] = b
        d["gb"] = tb.grad.numpy()
    meta = dict(kind="rot_" + kind, ndim=ndim, stride=stride, padding=padding, dilation=dilation, bias=bool(bias),
                output_padding=output_padding, quaternion_format=bool(qf), seed=seed, source=src)
    _save(name, meta, d)
    print(name, tuple(x.shape), "->", tuple(y.shape))


def qpointwise_case(ns, name, shape, seed):
    """hamilton_product (dual_quaternion_ops.py:374-414; quaternion_ops.py:467-507 is the same arithmetic on 2-d
    inputs), q_normalize and quaternion_exp (dual_quaternion_ops.py:206-246), each with its autograd gradients."""
    rng = np.random.default_rng(seed)
    a, b = f32(rng.standard_normal(shape)), f32(rng.standard_normal(shape))
    g = f32(rng.standard_normal(shape))
    d = dict(a=a, b=b, g=g)
    ta, tb_ = torch.tensor(a, requires_grad=True), torch.tensor(b, requires_grad=True)
    y = ns.dq_ops.hamilton_product(ta, tb_)
    y.backward(torch.tensor(g))
    d.update(ham=y.detach().numpy(), ham_ga=ta.grad.numpy(), ham_gb=tb_.grad.numpy())
    if len(shape) == 2:
        tq0, tq1 = torch.tensor(a), torch.tensor(b)
        assert np.array_equal(ns.q_ops.hamilton_product(tq0, tq1).numpy(), d["ham"])
    for key, fn in (("norm", ns.dq_ops.q_normalize), ("exp", ns.dq_ops.quaternion_exp)):
        ta = torch.tensor(a, requires_grad=True)
        y = fn(ta)
        y.backward(torch.tensor(g))
        d[key] = y.detach().numpy()
        d[key + "_g"] = ta.grad.numpy()
    _save(name, dict(kind="qpointwise", seed=seed, source="dual_quaternion_ops.py:206-246, :374-414"), d)
    print(name, shape)


def strided_transpose_cases(ns):
    """Transposed convolutions with stride > 1 and output_padding (fp32 kernels)."""
    convT_case(ns, "convT1d_q_s2", 1, 2, 8, 4, (31,), 3, 1, 1, True, 60, stride=2, output_padding=1)
    convT_case(ns, "convT2d_q_s2", 2, 1, 4, 8, (5, 19), 3, 1, 1, False, 61, stride=2, output_padding=0)
    rotation_case(ns, "rot_convT1d_s3", "convT", 1, 2, 4, 3, (17,), 3, 0, 2, True, True, 62, stride=3, output_padding=2)


def n4_cases(ns):
    """Round 2, SURVEY.md 8f N4: the operators of quaternion_ops.py / dual_quaternion_ops.py the SELD models never
    call."""
    rotation_case(ns, "rot_conv1d_qf", "conv", 1, 2, 8, 8, (70,), 3, 2, 2, True, True, 51)
    rotation_case(ns, "rot_conv2d_3c", "conv", 2, 1, 4, 8, (6, 40), 3, 1, 1, False, False, 52)
    rotation_case(ns, "rot_conv1d_s2", "conv", 1, 2, 3, 5, (41,), 3, 1, 1, True, True, 53, stride=2)
    rotation_case(ns, "rot_convT1d_qf", "convT", 1, 2, 8, 4, (50,), 3, 1, 2, True, True, 54)
    rotation_case(ns, "rot_convT2d_3c", "convT", 2, 1, 3, 5, (5, 33), 3, 1, 1, False, False, 55)
    rotation_case(ns, "rot_linear_qf", "linear", 0, 0, 6, 5, (9,), 0, 0, 0, True, True, 56)
    rotation_case(ns, "rot_linear_3c", "linear", 0, 0, 8, 8, (4, 7), 0, 0, 0, False, False, 57)
    if "--strided-only" in sys.argv:
        strided_transpose_cases(ns)
        return
    qpointwise_case(ns, "qpointwise_2d", (7, 20), 58)
    qpointwise_case(ns, "qpointwise_4d", (2, 12, 5, 9), 59)
    strided_transpose_cases(ns)


def stft_case(ns, name, C, n, nperseg, noverlap, phase, seed):
    rng = np.random.default_rng(seed)
    x = f32(0.1 * rng.standard_normal((C, n)))
    out = ns.uf.spectrum_fast(x, nperseg=nperseg, noverlap=noverlap, output_phase=phase)
    meta = dict(kind="stft", nperseg=nperseg, noverlap=noverlap, output_phase=phase, seed=seed,
                source="utility_functions.py:129-155")
    _save(name, meta, dict(x=x.astype(np.float32), out=out))
    print(name, x.shape, "->", out.shape)


def seld_loss(sed, doa, target, n_sed):
    """train.py:186-204 with sed_loss_weight=1, doa_loss_weight=5 (train.py:785-786)."""
    t_sed = torch.flatten(target[:, :, :n_sed], start_dim=1)
    t_doa = torch.flatten(target[:, :, n_sed:], start_dim=1)
    sed = torch.flatten(sed, start_dim=1)
    doa = torch.flatten(doa, start_dim=1)
    return torch.nn.BCELoss()(sed, t_sed) * 1.0 + torch.nn.MSELoss()(doa, t_doa) * 5.0


def model_case(name, cfg, time_dim, B, seed):
    """Whole-model forward + backward of model.SELD_Model (model.py:324-480), dropout off."""
    m = ref_import.build_reference_model(cfg, time_dim=time_dim, spatial_dropout_rate=0,
                                         dropout_perc=0, seed=seed)
    m.train()
    sd32 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.double()
    rng = np.random.default_rng(seed + 100)
    x = f32(rng.standard_normal((B, cfg["input_channels"], cfg["freq_dim"], time_dim)))
    n_out = time_dim // 8
    n_sed = 14 * 3
    sed_t = (rng.random((B, n_out, n_sed)) < 0.05).astype(np.float64)
    doa_t = (2 * rng.random((B, n_out, n_sed * 3)) - 1) * np.repeat(sed_t, 3, axis=-1)
    target = f32(np.concatenate([sed_t, doa_t], axis=-1))
    sed, doa = m(torch.tensor(x))
    loss = seld_loss(sed, doa, torch.tensor(target), n_sed)
    loss.backward()
    d = dict(x=x.astype(np.float32), target=target.astype(np.float32),
             sed=sed.detach().numpy(), doa=doa.detach().numpy(), loss=np.float64(loss.item()))
    for k, v in sd32.items():
        d["param/" + k] = v.numpy()
    ngrad = 0
    for k, p in m.named_parameters():
        if p.grad is not None:
            d["grad/" + k] = p.grad.numpy().astype(np.float32)
            ngrad += 1
    # yardstick 1: the reference's own float32 run against its float64 run, per gradient tensor
    m32 = ref_import.build_reference_model(cfg, time_dim=time_dim, spatial_dropout_rate=0, dropout_perc=0, seed=seed)
    m32.load_state_dict(sd32)
    m32.train()
    s32, d32 = m32(torch.tensor(x, dtype=torch.float32))
    seld_loss(s32, d32, torch.tensor(target, dtype=torch.float32), n_sed).backward()
    from oracle import algebra as A
    for k, p in m32.named_parameters():
        if p.grad is not None:
            d["ref32err/" + k] = np.float64(A.rel_err(p.grad.numpy(), d["grad/" + k]))
    # yardstick 2: an ideal bf16-operand implementation (oracle/bf16_emulation.py) of the same model
    # (key bf16emu*), and the same with the 2-d conv outputs stored in fp16 as the fused CNN path does (bf16emu16*)
    from oracle import bf16_emulation, cpu_model
    for tag, store in (("bf16emu", False), ("bf16emu16", True)):
        me = cpu_model.build_model(time_dim=time_dim, spatial_dropout_rate=0, dropout_perc=0, **cfg)
        me.load_state_dict(sd32)
        me = me.double().train()
        with bf16_emulation.bf16_operand_convs(store_conv2d_f16=store):
            se, de = me(torch.tensor(x))
            cpu_model.seld_loss(se, de, torch.tensor(target)).backward()
        d[tag + "/sed"], d[tag + "/doa"] = se.detach().numpy(), de.detach().numpy()
        for k, p in me.named_parameters():
            if p.grad is not None:
                d[tag + "_grad/" + k] = p.grad.numpy().astype(np.float32)
    # the reference's own DCASE21 SELD scores of its own outputs (train.py:84-130: gen_submission_list_task2 ->
    # segment_labels -> SELDMetrics), clip by clip; num_frames = the fixture's output frames
    ns_ = ref_import.load()
    mt = ns_.dcase.SELDMetrics(nb_classes=14, doa_threshold=20)
    for i in range(B):
        _, pd_ = ns_.uf.gen_submission_list_task2(d["sed"][i], d["doa"][i], max_overlaps=3, max_loc_value=2.0)
        _, td_ = ns_.uf.gen_submission_list_task2(d["target"][i][:, :n_sed], d["target"][i][:, n_sed:], max_overlaps=3,
                                                  max_loc_value=2.0)
        mt.update_seld_scores(ns_.dcase.segment_labels(pd_, n_out), ns_.dcase.segment_labels(td_, n_out))
    d["seld_scores"] = np.array([float(v) for v in mt.compute_seld_scores()], np.float64)
    # how close the SED outputs come to the 0.5 threshold (metric identity needs the error to stay below this)
    d["sed_margin"] = np.float64(np.abs(d["sed"] - 0.5).min())
    meta = dict(kind="model", cfg=cfg, time_dim=time_dim, B=B, seed=seed, model_name=m.model_name,
                n_params=int(sum(p.numel() for p in m.parameters())), n_grads=ngrad,
                source="model.py:324-480 + train.py:186-204")
    _save(name, meta, d)
    print(name, m.model_name, meta["n_params"], "params;", "loss", loss.item())


def sample_indices(numel, i, n=4096):
    """The fixed subset of a flattened tensor that full-size fixtures keep (tensor number i in named_parameters order)."""
    if numel <= n:
        return np.arange(numel)
    return np.sort(np.random.default_rng(1000 + i).choice(numel, n, replace=False))


def full_size_input(cfg, time_dim, B, seed):
    """Seeded input / target of a full-size fixture: regenerated by the tests, only check sums are stored."""
    rng = np.random.default_rng(seed + 100)
    x = rng.standard_normal((B, cfg["input_channels"], cfg["freq_dim"], time_dim)).astype(np.float32)
    n_out, n_sed = time_dim // 8, 14 * 3
    sed_t = (rng.random((B, n_out, n_sed)) < 0.05).astype(np.float64)
    doa_t = (2 * rng.random((B, n_out, n_sed * 3)) - 1) * np.repeat(sed_t, 3, axis=-1)
    return x, np.concatenate([sed_t, doa_t], axis=-1).astype(np.float32)


def full_size_case(name, cfg_name, time_dim, B, seed):
    """BASELINE-size forward + backward of the reference's model.SELD_Model in float64 (model.py:324-480 +
    train.py:186-204), dropout off.  The fixture keeps the outputs, the loss and, per parameter, the gradient's L2
    norm and a fixed 4096-element sample (sample_indices); input, target and initial weights regenerate from the
    seeds (full_size_input; np.random.seed / torch.manual_seed as train.py:214-221) and are pinned by check sums.
    Yard-sticks as in model_case: the reference's own float32 run, and the ideal bf16-operand emulation with fp16
    storage of the 2-d conv outputs (what the fused path does)."""
    from oracle import algebra as A
    cfg = dict(ref_import.COMMON)
    cfg.update(ref_import.CONFIGS[cfg_name])
    m = ref_import.build_reference_model(cfg, time_dim=time_dim, spatial_dropout_rate=0, dropout_perc=0, seed=seed)
    m.train()
    sd32 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x, target = full_size_input(cfg, time_dim, B, seed)
    n_sed = 14 * 3
    d = dict(x_sum=np.float64(x.astype(np.float64).sum()), x_abs=np.float64(np.abs(x.astype(np.float64)).sum()),
             target_sum=np.float64(target.astype(np.float64).sum()))
    names = [k for k, _ in m.named_parameters()]
    for k, v in sd32.items():
        if v.dtype.is_floating_point:
            d["psum/" + k] = np.float64(v.double().abs().sum().item())
    m = m.double()
    sed, doa = m(torch.tensor(x, dtype=torch.float64))
    loss = seld_loss(sed, doa, torch.tensor(target, dtype=torch.float64), n_sed)
    loss.backward()
    d.update(sed=sed.detach().numpy().astype(np.float32), doa=doa.detach().numpy().astype(np.float32),
             loss=np.float64(loss.item()))
    grads64, ngrad = {}, 0
    for i, (k, p) in enumerate(m.named_parameters()):
        if p.grad is None:
            continue
        g = p.grad.numpy().ravel()
        grads64[k] = g
        d["gnorm/" + k] = np.float64(np.linalg.norm(g))
        d["gmax/" + k] = np.float64(np.abs(g).max())
        d["gsample/" + k] = g[sample_indices(g.size, i)].astype(np.float32)
        ngrad += 1
    del m
    print(name, "float64 run done, loss", loss.item(), flush=True)
    m32 = ref_import.build_reference_model(cfg, time_dim=time_dim, spatial_dropout_rate=0, dropout_perc=0, seed=seed)
    m32.load_state_dict(sd32)
    m32.train()
    s32, d32 = m32(torch.tensor(x))
    seld_loss(s32, d32, torch.tensor(target), n_sed).backward()
    for k, p in m32.named_parameters():
        if p.grad is not None:
            d["ref32err/" + k] = np.float64(np.abs(p.grad.numpy().ravel().astype(np.float64) - grads64[k]).max()
                                            / max(d["gmax/" + k], 1e-300))
    d["ref32/sed_err"] = np.float64(A.rel_err(s32.detach().numpy(), d["sed"]))
    d["ref32/doa_err"] = np.float64(A.rel_err(d32.detach().numpy(), d["doa"]))
    del m32
    print(name, "float32 run done", flush=True)
    from oracle import bf16_emulation, cpu_model
    me = cpu_model.build_model(time_dim=time_dim, spatial_dropout_rate=0, dropout_perc=0, **cfg)
    me.load_state_dict(sd32)
    me = me.double().train()
    with bf16_emulation.bf16_operand_convs(store_conv2d_f16=True):
        se, de = me(torch.tensor(x, dtype=torch.float64))
        cpu_model.seld_loss(se, de, torch.tensor(target, dtype=torch.float64)).backward()
    d["bf16emu16/sed"], d["bf16emu16/doa"] = se.detach().numpy().astype(np.float32), de.detach().numpy().astype(np.float32)
    for k, p in me.named_parameters():
        if p.grad is not None:
            d["bf16emu16_err/" + k] = np.float64(np.abs(p.grad.numpy().ravel() - grads64[k]).max()
                                                 / max(d["gmax/" + k], 1e-300))
    meta = dict(kind="model_full", cfg=cfg, cfg_name=cfg_name, time_dim=time_dim, B=B, seed=seed, n_grads=ngrad,
                param_names=names, source="model.py:324-480 + train.py:186-204",
                input="oracle.make_golden.full_size_input(cfg, time_dim, B, seed)")
    _save(name, meta, d)
    print(name, "saved;", ngrad, "gradient tensors")


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = ref_import.load()
    if "--full-size" in sys.argv:
        full_size_case("model_dq_8ch_full", "DQ_8ch", 4800, 1, 1)
        return
    if "--round2" in sys.argv:
        round2_cases(ns)
        return
    if "--n4" in sys.argv:
        n4_cases(ns)
        return
    only_models = "--models-only" in sys.argv
    if not only_models:
        per_op_cases(ns)
    model_cases()


def round2_cases(ns):
    """Pins added in round 2: the 16-channel first layer (C4: compact 8 x (24, 2, 3, 3)), a reduced 16chMagPhase
    model whose first block has the real 16 -> 192 widths, and a reduced real-valued model (config
    SERVER_SELD-TCN-S1-PHI_8ch.txt: nn.Conv* layers)."""
    conv_case(ns, "conv2d_dq_first16", "DQ", 2, 1, 2, 24, (16, 136), 3, 1, 1, 1, False, 22)
    if "--convT" in sys.argv or True:
        convT_case(ns, "convT1d_q_k3_d2", 1, 2, 16, 8, (72,), 3, 1, 2, True, 23)
        convT_case(ns, "convT2d_q_3x3", 2, 1, 8, 16, (6, 40), 3, 1, 1, False, 24)
        convT_case(ns, "convT1d_q_small", 1, 2, 3, 5, (31,), 3, 0, 1, True, 25)
    if "--convT" in sys.argv:
        return
    tiny = dict(ref_import.COMMON)
    tiny.update(input_channels=8, freq_dim=128, domain="DQ", domain_classifier="DQ",
                cnn_filters=[16, 16, 16], G=16, U=16, V=[16, 16], fc_layers=[16],
                parallel_ConvTC_block="False", parallel_magphase=False, extra_name="_tiny")
    c4 = dict(tiny)
    c4.update(input_channels=16, freq_dim=256, cnn_filters=[192, 64, 64], G=128, U=128, V=[128, 128], fc_layers=[128],
              extra_name="_16chMagPhase_mid")
    model_case("model_dq_16ch_mid", c4, 160, 2, 5)
    r8 = dict(tiny)
    r8.update(domain="R", domain_classifier="R", freq_dim=256, cnn_filters=[32, 32, 32], G=64, U=64, V=[64, 64],
              fc_layers=[64], extra_name="_r_mid")
    model_case("model_r_mid", r8, 160, 2, 6)


def per_op_cases(ns):
    # per-op fixtures (SURVEY.md 8c suggested shapes, plus stride / bias / odd sizes)
    conv_case(ns, "conv1d_q_k3_d5", "Q", 1, 2, 16, 16, (97,), 3, 1, 5, 5, True, 11)
    conv_case(ns, "conv1d_dq_k3_d5", "DQ", 1, 2, 8, 8, (97,), 3, 1, 5, 5, True, 12)
    conv_case(ns, "conv1d_dq_k1", "DQ", 1, 3, 6, 5, (50,), 1, 1, 0, 1, False, 13)
    conv_case(ns, "conv1d_q_k3_s2", "Q", 1, 2, 3, 5, (41,), 3, 2, 1, 2, True, 14)
    conv_case(ns, "conv2d_q_3x3", "Q", 2, 1, 4, 8, (9, 23), 3, 1, 1, 1, False, 15)
    conv_case(ns, "conv2d_dq_3x3", "DQ", 2, 2, 2, 4, (9, 23), 3, 1, 1, 1, True, 16)
    conv_case(ns, "conv2d_dq_first", "DQ", 2, 1, 1, 3, (16, 40), 3, 1, 1, 1, False, 17)
    # tensor-core friendly shapes (channels per component a multiple of 8/16)
    conv_case(ns, "conv1d_dq_c48_d3", "DQ", 1, 1, 48, 48, (168,), 3, 1, 3, 3, False, 18)
    conv_case(ns, "conv1d_q_c32_d2", "Q", 1, 2, 32, 32, (150,), 3, 1, 2, 2, True, 19)
    conv_case(ns, "conv2d_dq_c24", "DQ", 2, 1, 24, 24, (6, 140), 3, 1, 1, 1, False, 20)
    conv_case(ns, "conv2d_q_c16", "Q", 2, 1, 16, 16, (5, 130), 3, 1, 1, 1, False, 21)
    linear_case(ns, "linear_q", "Q", 7, 6, 5, True, 31)
    linear_case(ns, "linear_q_fn", "Q", 9, 4, 8, True, 32, use_function=True)
    linear_case(ns, "linear_dq", "DQ", 7, 6, 5, True, 33)
    linear_case(ns, "linear_dq_c48", "DQ", 40, 48, 48, True, 34)
    stft_case(ns, "stft_mag", 8, 6400, 512, 112, False, 41)
    stft_case(ns, "stft_magphase", 8, 6400, 512, 112, True, 42)
    stft_case(ns, "stft_magphase_default", 3, 5000, 512, 128, True, 43)


def model_cases():
    # whole-model fixtures
    tiny = dict(ref_import.COMMON)
    tiny.update(input_channels=8, freq_dim=128, domain="DQ", domain_classifier="DQ",
                cnn_filters=[16, 16, 16], G=16, U=16, V=[16, 16], fc_layers=[16],
                parallel_ConvTC_block="False", parallel_magphase=False, extra_name="_tiny")
    model_case("model_dq_tiny", tiny, 64, 2, 1)
    tq = dict(tiny)
    tq.update(domain="Q", domain_classifier="Q", extra_name="_tinyq")
    model_case("model_q_tiny", tq, 64, 2, 2)
    mid = dict(tiny)
    mid.update(freq_dim=256, cnn_filters=[64, 64, 64], G=128, U=128, V=[128, 128], fc_layers=[128],
               extra_name="_mid")
    model_case("model_dq_mid", mid, 160, 2, 3)
    two = dict(tiny)
    two.update(input_channels=16, domain_classifier="R", parallel_ConvTC_block="2Parallel",
               parallel_magphase=True, extra_name="_two")
    model_case("model_dq_2branch_tiny", two, 64, 2, 4)


if __name__ == "__main__":
    main()
