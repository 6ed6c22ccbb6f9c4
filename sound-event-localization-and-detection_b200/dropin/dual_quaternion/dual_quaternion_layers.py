"""Drop-in for the reference's dual_quaternion/dual_quaternion_layers.py (star-imported by model.py:7)."""
import os
import sys

sys.path.append(os.path.join(os.path.dirname(__file__), '..', 'dual_quaternion'))   # as dual_quaternion_layers.py:9-11
from dual_quaternion_ops import *  # noqa: E402,F401,F403
from dual_quaternion_ops import _pkg  # noqa: E402

DualQuaternionConv = _pkg.DualQuaternionConv
DualQuaternionLinear = _pkg.DualQuaternionLinear
