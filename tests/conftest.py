import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_cases():
    """[(file stem, case name)] for every (code, decoder) record in tests/golden/*.npz."""
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        stem = os.path.splitext(os.path.basename(f))[0]
        if stem == "quantizer_kat" or stem.endswith("_next"):
            continue
        z = np.load(f)
        for c in sorted(set(k.split("/")[0] for k in z.files)):
            out.append((stem, c))
    return out


class Golden:
    def __init__(self, stem, case):
        self.z = np.load(os.path.join(GOLDEN_DIR, stem + ".npz"))
        self.case = case

    def __getitem__(self, key):
        return self.z[f"{self.case}/{key}"]

    def __contains__(self, key):
        return f"{self.case}/{key}" in self.z.files

    @property
    def kind(self):
        return str(self["kind"])


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree CUDA library (built on demand where nvcc exists; on the GPU box it is prebuilt)."""
    import __graft_entry__ as ge
    pkg_lib = os.path.join(ROOT, ge.PKG, "libldpc_b200.so")
    if not os.path.exists(pkg_lib):
        ge.build()
    import ldpc_b200
    return ldpc_b200
