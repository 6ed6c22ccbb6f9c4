"""GPU parity of the operators the SELD models never call: the rotation variants of the quaternion
convolution / transposed convolution / linear layer (quaternion_ops.py:174-295, :330-388), hamilton_product, q_normalize
and quaternion_exp (quaternion_ops.py:467-507, dual_quaternion_ops.py:206-246, :374-414) -- through the op / layer API,
i.e. through the C ABI, against fixtures minted from the reference (oracle/make_golden.py --n4) and the oracle.
Tolerances as in test_gpu_parity.py: rel 1e-4 in fp32 mode, 2e-2 on the tensor-core path."""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden
from oracle import algebra as A

pytestmark = pytest.mark.gpu

ROT_FIXTURES = golden_names("rot_conv") + golden_names("rot_convT") + golden_names("rot_linear")
# stride 1 and at least 8 channels on both sides: the layers the tensor-core path's dense mode takes
ROT_BF16 = ["rot_conv1d_qf", "rot_conv2d_3c", "rot_convT1d_qf", "rot_linear_3c"]


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()


@pytest.mark.parametrize("name", ROT_FIXTURES)
def test_rotation_weight_kernel_is_bit_exact(seldq, name):
    """seldq_rotation_weight: bit-identical to the oracle's float32 restatement in both layouts, and within one unit in the
    last place of the float32 weight the reference built on the CPU (its vectorised square root is not correctly rounded)."""
    meta, d = load_golden(name)
    ws = [cuda(d["w%d" % i]) for i in range(4)]
    qf = meta["quaternion_format"]
    want = A.rotation_weight([d["w%d" % i] for i in range(4)], qf, np.float32)
    got = seldq.functional.rotation_weight(ws, qf).cpu().numpy()
    got_t = seldq.functional.rotation_weight(ws, qf, transpose_out=True).cpu().numpy()
    assert np.array_equal(got, want)
    assert np.array_equal(got_t, np.swapaxes(want, 0, 1))
    assert np.abs(got.reshape(d["W32"].shape).astype(np.float64) - d["W32"]).max() <= 1.5e-7 * np.abs(d["W32"]).max()


def _run_rotation(seldq, meta, d, prec):
    F = seldq.functional
    x = cuda(d["x"]).requires_grad_(True)
    ws = [cuda(d["w%d" % i]).requires_grad_(True) for i in range(4)]
    b = cuda(d["b"]).requires_grad_(True) if meta["bias"] else None
    qf, kind = meta["quaternion_format"], meta["kind"]
    if kind == "rot_conv":
        y = F.quaternion_conv_rotation(x, ws, b, meta["stride"], meta["padding"], 1, meta["dilation"], qf, prec=prec)
    elif kind == "rot_convT":
        y = F.quaternion_transpose_conv_rotation(x, ws, b, meta["stride"], meta["padding"], meta.get("output_padding", 0), 1,
                                                 meta["dilation"], qf, prec=prec)
    else:
        y = F.quaternion_linear_rotation(x, ws, b, qf, prec=prec)
    y.backward(cuda(d["gy"]))
    torch.cuda.synchronize()
    errs = {"y": A.rel_err(y.detach().cpu().numpy(), d["y"]), "gx": A.rel_err(x.grad.cpu().numpy(), d["gx"])}
    for i in range(4):
        errs["gw%d" % i] = A.rel_err(ws[i].grad.cpu().numpy(), d["gw%d" % i])
    if meta["bias"]:
        errs["gb"] = A.rel_err(b.grad.cpu().numpy(), d["gb"])
    return errs


@pytest.mark.parametrize("name", ROT_FIXTURES)
def test_rotation_variants_match_reference_fixture_fp32(seldq, name):
    meta, d = load_golden(name)
    errs = _run_rotation(seldq, meta, d, seldq._lib.PREC_FP32)
    assert all(v < 1e-4 for v in errs.values()), (name, errs)


@pytest.mark.parametrize("name", ROT_BF16)
def test_rotation_variants_on_the_tensor_core_path(seldq, name, monkeypatch):
    """The real algebra on the tcgen05 kernels (dense mode: the expanded tile is built from the fp32 weight in the
    kernel's prologue): what the rotation variants run in bf16 mode unless SELDQ_ROTATION_BF16=0."""
    monkeypatch.setattr(seldq.functional, "_ROTATION_BF16", True)
    meta, d = load_golden(name)
    errs = _run_rotation(seldq, meta, d, seldq._lib.PREC_BF16)
    assert all(v < 2e-2 for v in errs.values()), (name, errs)


def test_rotation_layers_follow_the_reference_constructors(seldq):
    """QuaternionConv / QuaternionTransposeConv / QuaternionLinearAutograd with rotation=True (quaternion_layers.py:75-80,
    :151-155, :212-214) against the oracle on the layers' own seeded parameters."""
    L = seldq.layers
    rng = np.random.default_rng(3)

    def params(layer):
        return [getattr(layer, n).detach().cpu().numpy().astype(np.float64) for n in ("r_weight", "i_weight", "j_weight", "k_weight")]

    with seldq.precision("fp32"):
        conv = L.QuaternionConv(16, 32, 3, 1, padding=1, bias=True, operation="convolution1d", rotation=True,
                                quaternion_format=True, seed=5).cuda()
        x = rng.standard_normal((2, 16, 37)).astype(np.float32)
        y = conv(cuda(x))
        want = A.qconv_rotation(x, params(conv), conv.bias.detach().cpu().numpy(), 1, 1, 1, True)
        assert A.rel_err(y.detach().cpu().numpy(), want) < 1e-4
        convt = L.QuaternionTransposeConv(16, 8, 3, 1, padding=1, bias=False, operation="convolution1d", rotation=True,
                                          quaternion_format=True, seed=6).cuda()
        yt = convt(cuda(x))
        want = A.qconv_transpose_rotation(x, params(convt), None, 1, 1, True)
        assert A.rel_err(yt.detach().cpu().numpy(), want) < 1e-4
        lin = L.QuaternionLinearAutograd(16, 24, bias=True, rotation=True, quaternion_format=True, seed=7).cuda()
        xl = rng.standard_normal((5, 3, 16)).astype(np.float32)
        yl = lin(cuda(xl))
        want = A.qlinear_rotation(xl, params(lin), lin.bias.detach().cpu().numpy(), True)
        assert yl.shape == (5, 3, 24) and A.rel_err(yl.detach().cpu().numpy(), want) < 1e-4


@pytest.mark.parametrize("name", golden_names("qpointwise"))
def test_quaternion_pointwise_operators_match_reference_fixture(seldq, name):
    meta, d = load_golden(name)
    F = seldq.functional
    g = cuda(d["g"])
    a, b = cuda(d["a"]).requires_grad_(True), cuda(d["b"]).requires_grad_(True)
    y = F.hamilton_product(a, b)
    y.backward(g)
    assert np.array_equal(y.detach().cpu().numpy(), A.hamilton_product(d["a"], d["b"], np.float32))
    assert A.rel_err(y.detach().cpu().numpy(), d["ham"]) < 1e-6
    assert A.rel_err(a.grad.cpu().numpy(), d["ham_ga"]) < 1e-6 and A.rel_err(b.grad.cpu().numpy(), d["ham_gb"]) < 1e-6
    for key, fn, exact in (("norm", F.q_normalize, True), ("exp", F.quaternion_exp, False)):
        a = cuda(d["a"]).requires_grad_(True)
        y = fn(a)
        y.backward(g)
        if exact:
            assert np.array_equal(y.detach().cpu().numpy(), A.q_normalize(d["a"], np.float32))
        assert A.rel_err(y.detach().cpu().numpy(), d[key]) < 1e-6, key
        assert A.rel_err(a.grad.cpu().numpy(), d[key + "_g"]) < 1e-5, key


def test_dropin_modules_expose_the_n4_operators(seldq):
    """The reference's own module names (dropin/): rotation variants, hamilton_product, q_normalize, quaternion_exp."""
    import importlib.util
    import os

    def load(rel, name):            # by path: other tests of the session may have imported the reference's own modules
        spec = importlib.util.spec_from_file_location(name, os.path.join(seldq.DROPIN_DIR, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    q_ops = load("quaternion/quaternion_ops.py", "_seldq_dropin_q_ops")
    dq_ops = load("dual_quaternion/dual_quaternion_ops.py", "_seldq_dropin_dq_ops")
    meta, d = load_golden("rot_linear_qf")
    ws = [cuda(d["w%d" % i]) for i in range(4)]
    with seldq.precision("fp32"):
        y = q_ops.quaternion_linear_rotation(cuda(d["x"]), *ws, cuda(d["b"]), True)
    assert A.rel_err(y.cpu().numpy(), d["y"]) < 1e-4
    meta, d = load_golden("qpointwise_2d")
    assert A.rel_err(q_ops.hamilton_product(cuda(d["a"]), cuda(d["b"])).cpu().numpy(), d["ham"]) < 1e-6
    assert A.rel_err(dq_ops.hamilton_product(cuda(d["a"]), cuda(d["b"])).cpu().numpy(), d["ham"]) < 1e-6
    assert A.rel_err(dq_ops.q_normalize(cuda(d["a"])).cpu().numpy(), d["norm"]) < 1e-6
    assert A.rel_err(dq_ops.quaternion_exp(cuda(d["a"])).cpu().numpy(), d["exp"]) < 1e-6
    with pytest.raises(NotImplementedError):
        dq_ops.q_normalize(torch.zeros(2, 3, 8, device="cuda"))
