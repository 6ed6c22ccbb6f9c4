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
from .functional import (block_conv, block_linear, get_precision, precision, set_precision,
                         stft_magphase)
from .features import spectrum_fast
from .layers import (DualQuaternionConv, DualQuaternionLinear, QuaternionConv, QuaternionLinear,
                     QuaternionLinearAutograd, QuaternionTransposeConv)
from .seld_model import ConvTC_Block, MultiHeadAttention, ResBlock, SELD_Model, TC_Block

DROPIN_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "dropin")


def build(force=False):
    """Compile libseldq.so in-tree for sm_100a."""
    return _lib.build(force=force)


def install_dropin(patch_utility_functions=True):
    """Make `from quaternion.quaternion_layers import *` / `from dual_quaternion.dual_quaternion_layers
    import *` (model.py:7-8) resolve to this implementation: puts the drop-in directory first on
    sys.path and, if the reference's utility_functions is (or gets) imported, rebinds its
    spectrum_fast."""
    import sys
    for p in (DROPIN_DIR,):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    if patch_utility_functions and "utility_functions" in sys.modules:
        sys.modules["utility_functions"].spectrum_fast = spectrum_fast
