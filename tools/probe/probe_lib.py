"""ctypes handle of the tools-only probe library (tools/probe/libseldq_probe.so)."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "libseldq_probe.so")


def lib():
    if not os.path.exists(PATH):
        subprocess.check_call(["bash", os.path.join(_HERE, "build.sh")])
    return ctypes.CDLL(PATH)
