import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG = "sound-event-localization-and-detection_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a box without a CUDA device.  On a GPU box nothing is skipped: a
    missing libseldq.so must fail loudly there (there is no fallback path)."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    return meta, {k: z[k] for k in z.files if k != "meta"}


def golden_names(kind):
    out = []
    for f in sorted(os.listdir(GOLDEN)):
        if f.endswith(".npz"):
            z = np.load(os.path.join(GOLDEN, f), allow_pickle=False)
            if json.loads(str(z["meta"]))["kind"] == kind:
                out.append(f[:-4])
    return out


@pytest.fixture(scope="session")
def seldq():
    """The product package (its directory name is not a Python identifier)."""
    return importlib.import_module(PKG)
