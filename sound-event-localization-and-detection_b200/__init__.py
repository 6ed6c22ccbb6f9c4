"""seldq-b200: B200 (sm_100a) implementation of the quaternion / dual-quaternion convolution and
linear stack and the STFT front end of AuroraEchos/Sound-Event-Localization-and-Detection, behind
the reference's own layer / op signatures.  See DESIGN.md and INTEGRATION.md at the repo root.

The directory name is not a Python identifier; import it with
    importlib.import_module("sound-event-localization-and-detection_b200")
or put `<this dir>/dropin` on sys.path and import the reference's module names
(quaternion.quaternion_layers, dual_quaternion.dual_quaternion_layers, quaternion_ops, ...).
"""
import os as _os

from . import _lib
from . import fused
from .functional import (attention, block_conv, block_linear, get_precision, precision, set_precision,
                         stft_magphase)
from .features import gen_submission_list_task2, spectrum_fast
from .layers import (DualQuaternionConv, DualQuaternionLinear, QuaternionConv, QuaternionLinear,
                     QuaternionLinearAutograd, QuaternionTransposeConv)
from .seld_model import ConvTC_Block, MultiHeadAttention, ResBlock, SELD_Model, TC_Block

DROPIN_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "dropin")


def build(force=False):
    """Compile libseldq.so in-tree for sm_100a."""
    return _lib.build(force=force)


_DROPIN = {"fuse_model": False}


def install_dropin(patch_utility_functions=True, fuse_model=True):
    """Make `from quaternion.quaternion_layers import *` / `from dual_quaternion.dual_quaternion_layers
    import *` (model.py:7-8) resolve to this implementation: puts the drop-in directory first on
    sys.path and, if the reference's utility_functions is (or gets) imported, rebinds its
    spectrum_fast.

    fuse_model=True: the glue BETWEEN the convolutions (model.py:109-132, :210-231, :276-283) also runs in this
    repository's kernels, without editing model.py: the first Q / DQ layer that model.py constructs (model.py is
    fully imported by then) rebinds the forward methods of model.TC_Block / ConvTC_Block / MultiHeadAttention to
    seld_model.tc_block_forward / convtc_block_forward / mha_forward (seld_model.patch_reference_model).  Every
    case the fused kernels do not serve (eval mode, fp32 mode, CPU, biases, other pooling ...) falls through to
    the reference's own layer-by-layer code path."""
    import sys
    for p in (DROPIN_DIR,):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    if patch_utility_functions and "utility_functions" in sys.modules:
        sys.modules["utility_functions"].spectrum_fast = spectrum_fast
        sys.modules["utility_functions"].gen_submission_list_task2 = gen_submission_list_task2
    _DROPIN["fuse_model"] = bool(fuse_model)
    if fuse_model:
        _maybe_patch_reference_model()


def _maybe_patch_reference_model():
    """Called by install_dropin and by the constructors of the Q / DQ layer modules (layers.py)."""
    if not _DROPIN["fuse_model"]:
        return
    import sys
    mod = sys.modules.get("model")
    if mod is not None and getattr(mod, "DualQuaternionConv", None) is DualQuaternionConv:
        from . import seld_model
        seld_model.patch_reference_model(mod)
