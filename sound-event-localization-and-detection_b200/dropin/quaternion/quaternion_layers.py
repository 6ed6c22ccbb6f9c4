"""Drop-in for the reference's quaternion/quaternion_layers.py (star-imported by model.py:8)."""
import os
import sys

sys.path.append(os.path.join(os.path.dirname(__file__), '..', 'quaternion'))   # as quaternion_layers.py:12-14
from quaternion_ops import *  # noqa: E402,F401,F403
from quaternion_ops import _pkg  # noqa: E402

QuaternionConv = _pkg.QuaternionConv
QuaternionLinear = _pkg.QuaternionLinear
QuaternionLinearAutograd = _pkg.QuaternionLinearAutograd
QuaternionTransposeConv = _pkg.QuaternionTransposeConv
