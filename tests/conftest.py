import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["kwargs"] = json.loads(str(d["kwargs"]))
    d["method"] = str(d["method"])
    d["name"] = name
    return d


def golden_names(method=None):
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        name = os.path.splitext(os.path.basename(p))[0]
        if method is None or str(np.load(p)["method"]) == method:
            out.append(name)
    return out


def rel_err(a, b):
    """max |a-b| / |b| over entries where b != 0, plus exact-zero agreement elsewhere."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    nz = b != 0
    if not np.array_equal(a != 0, nz):
        return float("inf")
    if not nz.any():
        return 0.0
    return float(np.max(np.abs(a[nz] - b[nz]) / np.abs(b[nz])))
