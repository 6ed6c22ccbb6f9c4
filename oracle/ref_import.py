"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Imports the *unmodified* reference (AuroraEchos/Sound-Event-Localization-and-Detection)
from /root/reference so that the oracle restatement in ``oracle/algebra.py`` can be pinned
against it and golden vectors can be minted (``oracle/make_golden.py``).

/root/reference exists only in the build container, not on the GPU box.  What travels to the box is
``oracle/_ref`` -- a verbatim, git-ignored copy of the reference's Python files made by the committed
recipe ``oracle/fetch_ref.sh`` (the reference is pure Python, so copying is its whole build).  The root
is resolved in this order: $SELDQ_REFERENCE_ROOT, /root/reference, oracle/_ref; everything here is
guarded by ``available()``.  Only tests/, smoke() and bench.py's reference arm may import this module.

Stub recipe: SURVEY.md Appendix A (model.py:3 needs torchinfo, utility_functions.py:9 needs
librosa; metrics.py:6-10 needs jiwer/pystoi/transformers).
"""
import importlib
import importlib.machinery
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
SHIPPED_ROOT = os.path.join(_HERE, "_ref")


def _resolve_root():
    env = os.environ.get("SELDQ_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", SHIPPED_ROOT):
        if os.path.isfile(os.path.join(cand, "model.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _resolve_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__dict__.update(attrs)
    sys.modules[name] = m


_REF_MODULE_NAMES = (
    "model", "utility_functions", "Dcase21_metrics",
    "quaternion_ops", "dual_quaternion_ops",
    "quaternion", "quaternion.quaternion_layers",
    "dual_quaternion", "dual_quaternion.dual_quaternion_layers",
)

_NS = None


def _is_reference_module(mod):
    f = getattr(mod, "__file__", None)
    if f is not None:
        return os.path.abspath(f).startswith(REFERENCE_ROOT)
    return any(os.path.abspath(p).startswith(REFERENCE_ROOT) for p in getattr(mod, "__path__", []))


def load():
    """Namespace with the reference modules (q_ops, q_layers, dq_ops, dq_layers, model, uf, dcase).
    Same-named modules imported from elsewhere (the product's drop-in directory) are evicted
    first so that the oracle is always the real reference."""
    global _NS
    if _NS is not None:
        return _NS
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _stub("torchinfo", summary=lambda *a, **k: None)
    _stub("librosa")
    for name in _REF_MODULE_NAMES:
        mod = sys.modules.get(name)
        if mod is not None and not _is_reference_module(mod):
            del sys.modules[name]
    # model.py:7-8 imports the two namespace packages; the layer files import their ops module
    # as a top-level name after appending their own directory (quaternion_layers.py:12-16)
    for p in (os.path.join(REFERENCE_ROOT, "dual_quaternion"), os.path.join(REFERENCE_ROOT, "quaternion"),
              REFERENCE_ROOT):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    ns = types.SimpleNamespace()
    ns.q_layers = importlib.import_module("quaternion.quaternion_layers")
    ns.dq_layers = importlib.import_module("dual_quaternion.dual_quaternion_layers")
    ns.q_ops = importlib.import_module("quaternion_ops")
    ns.dq_ops = importlib.import_module("dual_quaternion_ops")
    ns.model = importlib.import_module("model")
    ns.uf = importlib.import_module("utility_functions")
    ns.dcase = importlib.import_module("Dcase21_metrics")
    for m in (ns.q_layers, ns.dq_layers, ns.q_ops, ns.dq_ops, ns.model, ns.uf, ns.dcase):
        assert _is_reference_module(m), m
    _NS = ns
    return ns


# hyper-parameters train.py would pass after parsing config/SERVER_*.txt
CONFIGS = {
    "DQ_8ch": dict(input_channels=8, domain="DQ", domain_classifier="DQ", cnn_filters=[192, 192, 192],
                   G=384, U=384, V=[384, 384], fc_layers=[384], parallel_ConvTC_block="False",
                   parallel_magphase=False, extra_name="_8ch"),
    "DQ_16ch": dict(input_channels=16, domain="DQ", domain_classifier="DQ", cnn_filters=[192, 192, 192],
                    G=384, U=384, V=[384, 384], fc_layers=[384], parallel_ConvTC_block="False",
                    parallel_magphase=False, extra_name="_16chMagPhase"),
    "DQ_2branch": dict(input_channels=16, domain="DQ", domain_classifier="R", cnn_filters=[192, 192, 192],
                       G=384, U=384, V=[384, 384], fc_layers=[128], parallel_ConvTC_block="2Parallel",
                       parallel_magphase=True, extra_name="_micAMagPhaseParallelmicBMagPhase"),
    "Q_8ch": dict(input_channels=8, domain="Q", domain_classifier="R", cnn_filters=[64, 64, 64],
                  G=128, U=128, V=[128, 128], fc_layers=[128], parallel_ConvTC_block="1",
                  parallel_magphase=False, extra_name="_parallel_8ch"),
    "R_8ch": dict(input_channels=8, domain="R", domain_classifier="R", cnn_filters=[64, 64, 64],
                  G=128, U=128, V=[128, 128], fc_layers=[128], parallel_ConvTC_block="False",
                  parallel_magphase=False, extra_name="_8ch"),
}

COMMON = dict(freq_dim=256, output_classes=14, kernel_size_cnn_blocks=3,
              pool_size=[[8, 2], [8, 2], [2, 2]], pool_time="TCN", D=[10], dilation_mode="fibonacci",
              kernel_size_dilated_conv=3, V_kernel_size=3, fc_activations="linear", fc_dropout="Last",
              class_overlaps=3, use_bias_conv=0, use_bias_linear=1, batch_norm="BN")


def build_reference_model(cfg="DQ_8ch", time_dim=4800, spatial_dropout_rate=0.5, dropout_perc=0.3,
                          seed=1, **overrides):
    import numpy as np
    import torch
    ns = load()
    np.random.seed(seed)
    torch.manual_seed(seed)
    kw = dict(COMMON)
    kw.update(CONFIGS[cfg] if isinstance(cfg, str) else cfg)
    kw.update(overrides)
    m = ns.model.SELD_Model(time_dim=time_dim, spatial_dropout_rate=spatial_dropout_rate,
                            dropout_perc=dropout_perc, **kw)
    return m


def load_model_on_dropin(pkg, fuse_model=True):
    """The reference's model.py, UNMODIFIED, imported on top of the product's drop-in layer modules
    (pkg.install_dropin(): model.py:7-8 then star-imports the sm_100a layers).  Returns the module object; the
    interpreter's sys.modules / sys.path are restored afterwards, so the real reference (load()) and the drop-in
    build can live in one test process."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _stub("torchinfo", summary=lambda *a, **k: None)
    _stub("librosa")
    saved_mods = {n: sys.modules.pop(n) for n in _REF_MODULE_NAMES if n in sys.modules}
    saved_path = list(sys.path)
    try:
        sys.path[:] = [p for p in sys.path if not os.path.abspath(p).startswith(REFERENCE_ROOT)]
        pkg.install_dropin(fuse_model=fuse_model)
        sys.path.append(REFERENCE_ROOT)                    # model.py and utility_functions.py only
        mod = importlib.import_module("model")
        assert _is_reference_module(mod), mod.__file__
        assert mod.DualQuaternionConv is pkg.DualQuaternionConv, "model.py did not pick up the drop-in layers"
        if fuse_model:
            pkg._maybe_patch_reference_model()
        return mod
    finally:
        # keep the drop-in build reachable only through the returned module object
        for n in _REF_MODULE_NAMES:
            sys.modules.pop(n, None)
        sys.modules.update(saved_mods)
        sys.path[:] = saved_path
