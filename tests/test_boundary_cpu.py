"""CPU-side checks of the drop-in boundary: C-ABI symbols, state_dict compatibility, init parity,
error behaviour without a GPU.  No compute call succeeds here (there is no CPU fallback)."""
import ctypes
import importlib
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import PKG, ROOT, load_golden
from oracle import ref_import


def _cfg_kwargs(meta):
    cfg = dict(meta["cfg"])
    return dict(time_dim=meta["time_dim"], spatial_dropout_rate=0, dropout_perc=0, **cfg)


def test_library_exports_every_declared_symbol(seldq):
    seldq.build()
    header = open(os.path.join(ROOT, "include", "seldq.h")).read()
    declared = set(re.findall(r"\b(seldq_[a-z0-9_]+)\s*\(", header))
    declared -= {"seldq_status_t"}
    out = subprocess.check_output(["nm", "-D", "--defined-only", seldq._lib.LIB_PATH]).decode()
    exported = set(re.findall(r" T (seldq_[a-z0-9_]+)", out))
    assert declared, "no declarations parsed"
    assert declared <= exported, sorted(declared - exported)
    lib = seldq._lib.lib()
    assert lib.seldq_abi_version() == seldq._lib.ABI_VERSION == int(re.search(r"#define SELDQ_ABI_VERSION (\d+)", header).group(1))
    assert set(seldq._lib.exported_symbols()) <= exported


def test_every_declared_symbol_has_a_ctypes_prototype_and_an_integration_entry(seldq):
    """include/seldq.h is the boundary: each entry point must be bound by _lib.py (argument types checked by ctypes) and
    appear in INTEGRATION.md's table of the reference call sites it replaces."""
    header = open(os.path.join(ROOT, "include", "seldq.h")).read()
    declared = set(re.findall(r"\b(seldq_[a-z0-9_]+)\s*\(", header)) - {"seldq_status_t"}
    bound = set(seldq._lib.exported_symbols()) | {"seldq_abi_version", "seldq_last_error", "seldq_device_count"}
    assert declared <= bound, sorted(declared - bound)
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert not [f for f in sorted(declared) if f not in doc]


def test_descriptor_validation_runs_without_gpu(seldq):
    L = seldq._lib
    lib = L.lib()
    d = L.ConvDesc(L.ALG_DQ, L.PREC_FP32, 1, 2, 384, 384, 1, 4800, 1, 3, 1, 1, 0, 5, 1, 5)
    oh, ow = ctypes.c_int32(), ctypes.c_int32()
    assert lib.seldq_conv_out_shape(ctypes.byref(d), ctypes.byref(oh), ctypes.byref(ow)) == 0
    assert (oh.value, ow.value) == (1, 4800)
    bad = L.ConvDesc(L.ALG_DQ, L.PREC_FP32, 1, 2, 12, 384, 1, 4800, 1, 3, 1, 1, 0, 5, 1, 5)
    assert lib.seldq_conv_out_shape(ctypes.byref(bad), None, None) == L.ERR_INVALID
    assert b"divisible by 8" in lib.seldq_last_error()
    nb, nf = ctypes.c_int32(), ctypes.c_int32()
    assert lib.seldq_stft_shape(1920000, 512, 112, 1, 1, ctypes.byref(nb), ctypes.byref(nf)) == 0
    assert (nb.value, nf.value) == (256, 4800)
    d16 = L.ConvDesc(L.ALG_DQ, L.PREC_BF16, 1, 1, 384, 384, 1, 4800, 1, 3, 1, 1, 0, 1, 1, 1)
    # bf16 path: channels-last operands, every component padded to a multiple of 16 channels
    cp, dense = ctypes.c_int32(), ctypes.c_int32()
    clb, t16b = ctypes.c_size_t(), ctypes.c_size_t()
    assert lib.seldq_conv_operand_info(ctypes.byref(d16), 0, ctypes.byref(cp), ctypes.byref(dense), ctypes.byref(clb),
                                       ctypes.byref(t16b)) == 0
    assert (cp.value, dense.value) == (384, 0)
    assert clb.value >= 4800 * 384 * 2 and t16b.value >= 384 * 4800 * 2
    # CNN layers: 24 channels per component -> 32, first layer (1 channel per component) -> dense, 16 channels
    cnn1 = L.ConvDesc(L.ALG_DQ, L.PREC_BF16, 2, 1, 192, 192, 32, 4800, 3, 3, 1, 1, 1, 1, 1, 1)
    assert lib.seldq_conv_operand_info(ctypes.byref(cnn1), 0, ctypes.byref(cp), ctypes.byref(dense), None, None) == 0
    assert (cp.value, dense.value) == (256, 0)
    cnn0 = L.ConvDesc(L.ALG_DQ, L.PREC_BF16, 2, 1, 8, 192, 256, 4800, 3, 3, 1, 1, 1, 1, 1, 1)
    assert lib.seldq_conv_operand_info(ctypes.byref(cnn0), 0, ctypes.byref(cp), ctypes.byref(dense), None, None) == 0
    assert (cp.value, dense.value) == (16, 1)
    assert lib.seldq_conv_packed_bytes(ctypes.byref(cnn0), L.PASS_FWD) == 0
    # packed compact weights: 8 images x 3 taps x 3 slabs x (48 x 16) bf16 -- never the expanded weight
    assert lib.seldq_conv_packed_bytes(ctypes.byref(d16), L.PASS_FWD) == 8 * 3 * 3 * 48 * 16 * 2
    assert lib.seldq_conv_packed_bytes(ctypes.byref(d16), L.PASS_DGRAD) == 8 * 3 * 3 * 48 * 16 * 2
    assert lib.seldq_conv_workspace_bytes(ctypes.byref(d16), L.PASS_FWD) == clb.value * 0 + 4800 * 384 * 2 + 8 * 3 * 3 * 48 * 32


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(seldq):
    conv = seldq.DualQuaternionConv(16, 16, kernel_size=3, stride=1, padding=1, operation='convolution1d')
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        conv(torch.randn(1, 16, 32))
    lin = seldq.QuaternionLinear(8, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lin(torch.randn(3, 8))
    with pytest.raises(RuntimeError):
        seldq.spectrum_fast(np.zeros((2, 4000), np.float32))


@pytest.mark.parametrize("name", ["model_dq_tiny", "model_q_tiny", "model_dq_2branch_tiny", "model_dq_mid"])
def test_state_dict_matches_reference_checkpoint_layout(seldq, name):
    meta, d = load_golden(name)
    m = seldq.SELD_Model(**_cfg_kwargs(meta))
    sd = m.state_dict()
    ref_keys = [k[len("param/"):] for k in d if k.startswith("param/")]
    assert list(sd.keys()) == ref_keys or set(sd.keys()) == set(ref_keys)
    for k in ref_keys:
        assert tuple(sd[k].shape) == tuple(d["param/" + k].shape), k
    m.load_state_dict({k: torch.from_numpy(np.asarray(d["param/" + k])) for k in ref_keys}, strict=True)
    assert m.model_name == meta["model_name"]


@pytest.mark.parametrize("name", ["model_dq_tiny", "model_q_tiny", "model_dq_2branch_tiny"])
def test_init_is_bit_identical_to_reference_fixture(seldq, name):
    """Same numpy / torch seeds -> same initial weights as the reference (its init is a pure
    function of the global RNG streams)."""
    meta, d = load_golden(name)
    np.random.seed(meta["seed"])
    torch.manual_seed(meta["seed"])
    m = seldq.SELD_Model(**_cfg_kwargs(meta))
    for k, v in m.state_dict().items():
        ref = np.asarray(d["param/" + k])
        assert np.array_equal(v.numpy(), ref), k


@pytest.mark.reference
@pytest.mark.skipif(not ref_import.available(), reason="reference tree only exists in the build container")
def test_reference_model_py_builds_on_dropin_layers():
    """model.py, unmodified, star-imports the drop-in modules and assembles the DQ model with the
    same parameter names / shapes / values as with its own layers."""
    code = r'''
import sys, types, importlib, importlib.machinery
import numpy as np, torch
sys.path.insert(0, %r)
pkg = importlib.import_module(%r)
def stub(name, **a):
    m = types.ModuleType(name); m.__spec__ = importlib.machinery.ModuleSpec(name, None); m.__dict__.update(a); sys.modules[name] = m
stub("torchinfo", summary=lambda *a, **k: None); stub("librosa")
sys.path.insert(0, "/root/reference")
pkg.install_dropin()
import model
assert "dropin" in sys.modules["quaternion.quaternion_layers"].__file__
assert model.DualQuaternionConv is pkg.DualQuaternionConv and model.QuaternionLinear is pkg.QuaternionLinear
np.random.seed(1); torch.manual_seed(1)
m = model.SELD_Model(time_dim=64, freq_dim=128, input_channels=8, output_classes=14, domain="DQ",
    domain_classifier="DQ", cnn_filters=[16,16,16], G=16, U=16, V=[16,16], fc_layers=[16], fc_dropout="Last",
    fc_activations="linear", pool_time="TCN", batch_norm="BN", use_bias_conv=0, use_bias_linear=1, class_overlaps=3,
    spatial_dropout_rate=0, dropout_perc=0, extra_name="_tiny")
z = np.load(%r)
for k, v in m.state_dict().items():
    assert np.array_equal(v.numpy(), z["param/" + k]), k
print("OK", m.model_name)
''' % (ROOT, PKG, os.path.join(ROOT, "tests", "golden", "model_dq_tiny.npz"))
    out = subprocess.check_output([sys.executable, "-c", code], cwd="/tmp").decode()
    assert "OK DualQSELD-TCN-PHI-S1_BN_RF287_10RB_tiny" in out


def test_fused_paths_decline_what_they_do_not_serve(seldq):
    """fused.cnn_stack_supported / tcn_stack_supported gate the fused kernels: CPU tensors, eval mode and fp32
    precision fall through to the layer-by-layer modules (which raise on CPU: there is no CPU path)."""
    import importlib
    import torch
    model_mod = importlib.import_module(seldq.__name__ + ".seld_model")
    blk = model_mod.TC_Block(in_channels=128, domain="DQ", G=128, U=128, V=[128, 128], D=[2], spatial_dropout_rate=0.5,
                             use_bias_conv=False, batch_norm="BN")
    x = torch.zeros(1, 128, 64)
    assert not seldq.fused.tcn_stack_supported(blk.ResBlocks, x, True)            # CPU tensor
    assert not seldq.fused.tcn_stack_supported(blk.ResBlocks, x, False)           # eval mode
    nobn = model_mod.TC_Block(in_channels=128, domain="DQ", G=128, U=128, V=[128, 128], D=[1], use_bias_conv=False,
                              batch_norm="noBN")
    assert not seldq.fused.tcn_stack_supported(nobn.ResBlocks, x, True)


def test_trainer_keeps_parameters_and_gradients_in_flat_buffers(seldq):
    """trainer.FlatGradBucket: every parameter / gradient is a view into ONE tensor (single all-reduce, single-tensor
    Adam); an optimiser step through the flat parameter moves the module's parameters."""
    import importlib
    import torch
    trainer_mod = importlib.import_module(seldq.__name__ + ".trainer")
    m = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    before = [p.detach().clone() for p in m.parameters()]
    tr = trainer_mod.Trainer(m, lr=1e-2, n_sed=1)
    base = tr.bucket.flat_param.data.untyped_storage().data_ptr()
    for p, b in zip(m.parameters(), before):
        assert torch.equal(p.detach(), b)
        assert p.data.untyped_storage().data_ptr() == base
        assert p.grad.untyped_storage().data_ptr() == tr.bucket.flat.untyped_storage().data_ptr()
    tr.bucket.zero()
    m(torch.ones(5, 4)).sum().backward()
    tr.optimizer.step()
    assert all(not torch.equal(p.detach(), b) for p, b in zip(m.parameters(), before))


@pytest.mark.skipif(not os.path.isdir("/root/reference/quaternion"), reason="reference tree only exists in the build container")
def test_dropin_modules_define_every_public_name_of_the_reference_modules():
    """A user of the reference who switches sys.path to the drop-in finds every function and class the four hot-path
    modules define (quaternion_ops / quaternion_layers / dual_quaternion_ops / dual_quaternion_layers)."""
    import ast

    def defs(path):
        return {n.name for n in ast.parse(open(path).read()).body if isinstance(n, (ast.FunctionDef, ast.ClassDef))}

    def names(path):
        out = set()
        for n in ast.parse(open(path).read()).body:
            if isinstance(n, (ast.FunctionDef, ast.ClassDef)):
                out.add(n.name)
            elif isinstance(n, ast.Assign):
                out |= {t.id for t in n.targets if isinstance(t, ast.Name)}
        return out

    dropin = os.path.join(ROOT, PKG, "dropin")
    for rel in ("quaternion/quaternion_ops.py", "quaternion/quaternion_layers.py", "dual_quaternion/dual_quaternion_ops.py",
                "dual_quaternion/dual_quaternion_layers.py"):
        have = names(os.path.join(dropin, rel))
        if rel.endswith("_layers.py"):          # the layer modules star-import their ops module, as the reference's do
            have |= names(os.path.join(dropin, rel.replace("_layers.py", "_ops.py")))
        missing = defs(os.path.join("/root/reference", rel)) - have
        assert not missing, (rel, sorted(missing))


@pytest.mark.skipif(not os.path.isdir("/root/reference/quaternion"), reason="reference tree only exists in the build container")
def test_dropin_host_helpers_match_the_live_reference():
    """The host-side helpers that are plain tensor code in the reference too (component getters, get_modulus,
    get_normalized, create_dropout_mask, the depthwise-separable blocks), drop-in against reference, in a subprocess per
    side (both trees use the same module names)."""
    code = r'''
import sys, json, numpy as np, torch
from numpy.random import RandomState
sys.path.insert(0, %r)
if %r:
    sys.path.insert(0, %r); import importlib; importlib.import_module(%r).install_dropin()
from quaternion import quaternion_ops as q
from dual_quaternion import dual_quaternion_ops as dq
from dual_quaternion import dual_quaternion_layers as dql
torch.manual_seed(0)
x2, x3, x4 = torch.randn(5, 12), torch.randn(3, 5, 8), torch.randn(2, 8, 3, 4)
out = {}
out["q_mod"] = q.get_modulus(x2).tolist(); out["q_modv"] = q.get_modulus(x3, True).tolist()
out["q_norm"] = q.get_normalized(x2).tolist(); out["q_norm3"] = q.get_normalized(x3).tolist()
out["dq_mod"] = dq.get_modulus(x2).tolist(); out["dq_mod4"] = dq.get_modulus(x4, True).tolist()
out["dq_norm"] = dq.get_normalized(x3).tolist()
out["dq_k4"] = dq.get_k(x4).tolist(); out["q_j"] = q.get_j(x3).tolist()
out["mask"] = q.create_dropout_mask(0.3, (4, 6), RandomState(3), torch.FloatTensor).tolist()
out["dmask"] = np.asarray(dq.create_dropout_mask(0.5, (3, 3), RandomState(4), torch.FloatTensor)).tolist()
out["shape"] = [list(map(str, dq.get_kernel_and_weight_shape_dual("convolution2d", 3, 5, 3))),
                list(map(str, dq.get_kernel_and_weight_shape_dual("convolution1d", 3, 5, 2)))]
torch.manual_seed(1)
m = dql.DepthwiseSeparableConv1D(6, 10, 3, 1, 1)
out["ds_keys"] = sorted(m.state_dict()); out["ds"] = m(torch.randn(2, 6, 9)).tolist()
torch.manual_seed(2)
m2 = dql.DepthwiseSeparableConv2D(4, 6, 3, 2, 1)
out["ds2"] = m2(torch.randn(1, 4, 7, 8)).tolist()
print("RESULT" + json.dumps(out))
'''
    import json

    def run(dropin):
        root = os.path.join(ROOT, PKG, "dropin") if dropin else "/root/reference"
        src = code % (root, bool(dropin), ROOT, PKG)
        out = subprocess.check_output([sys.executable, "-c", src], stderr=subprocess.STDOUT).decode()
        return json.loads(out[out.index("RESULT") + 6:])

    ours, ref = run(True), run(False)
    assert sorted(ours) == sorted(ref)
    for k in ref:
        if isinstance(ref[k], list) and ref[k] and not isinstance(ref[k][0], str) and k not in ("shape", "ds_keys"):
            assert np.allclose(np.asarray(ours[k], np.float64), np.asarray(ref[k], np.float64), rtol=1e-6, atol=1e-7), k
        else:
            assert ours[k] == ref[k], k


def test_real_algebra_precision_choice_stays_inside_what_the_tensor_core_kernels_serve(seldq):
    """functional._real_prec (rotation variants): bf16 only for stride 1, 8 ... 256 channels on both sides, and channel
    counts that are both multiples of 8 or both <= 64 (csrc/wgrad_umma.cu); everything else runs the fp32 kernels."""
    F, L = seldq.functional, seldq._lib
    bf16 = [(32, 32), (24, 12), (16, 32), (24, 24), (72, 72), (60, 20), (256, 8)]
    fp32 = [(100, 40), (300, 8), (4, 4), (12, 4), (264, 264), (72, 12)]
    for ch in bf16:
        assert F._real_prec(L.PREC_BF16, 1, ch) == L.PREC_BF16, ch
        assert F._real_prec(L.PREC_BF16, 2, ch) == L.PREC_FP32, ch
        assert F._real_prec(L.PREC_FP32, 1, ch) == L.PREC_FP32, ch
    for ch in fp32:
        assert F._real_prec(L.PREC_BF16, 1, ch) == L.PREC_FP32, ch
