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
        if stem == "quantizer_kat" or stem.endswith("_next") or stem.startswith(("fullsize_", "train_")):
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


@pytest.fixture(autouse=True)
def _per_iteration_kernels_by_default(request, monkeypatch):
    """The parity suites use small codes and batches to exercise the per-iteration kernels, compaction, graph
    replay ...; left to the policy, such batches (at most one wave of thread blocks) would decode CTA-resident
    instead.  tests/test_gpu_resident.py covers that path and the policy itself and sets the variable on its own."""
    if request.module.__name__.rsplit(".", 1)[-1] != "test_gpu_resident":
        monkeypatch.setenv("LDPC_RESIDENT", "0")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree CUDA library (built on demand where nvcc exists; on the GPU box it is prebuilt)."""
    import __graft_entry__ as ge
    pkg_lib = os.path.join(ROOT, "ldpc_b200", "libldpc_b200.so")
    if not os.path.exists(pkg_lib):
        ge.build()
    import ldpc_b200
    return ldpc_b200


FULLSIZE_CASES = ["n2d2_dvbs2", "rcq_dvbs2", "wrcq1_qc", "rcq_layered_dvbs2_s4", "rcq_layered_qc"]


def load_fullsize(case):
    """(record, code) of a full-size golden case (tests/golden/make_golden_fullsize.py), or a pytest skip when the
    file is absent.  The graph is rebuilt from the package's seeded generator and checked against the stored hash."""
    import hashlib
    path = os.path.join(GOLDEN_DIR, f"fullsize_{case}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    import ldpc_b200 as L
    z = np.load(path)
    T = int(z["T"])
    if case.endswith("dvbs2_s4"):
        code = L.codes.dvbs2_shaped(max_iterations=T, scale=4)
    else:
        code = L.codes.dvbs2_shaped(max_iterations=T) if case.endswith("dvbs2") else L.codes.qc_shaped(max_iterations=T)
    g = code.graph
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(g.check_ptr).tobytes())
    h.update(np.ascontiguousarray(g.check_var).tobytes())
    assert h.hexdigest() == str(z["graph_sha256"]), "code generator changed: regenerate the full-size golden vectors"
    return z, code


def fullsize_tables(z, dec):
    """Load the recorded reference weights into our module; returns per-edge beta / per-variable alpha for the oracle."""
    import torch
    state = {str(k): torch.tensor([float(v)]) for k, v in zip(z["weight_keys"], z["weight_vals"])}
    dec.load_reference_state_dict(state)
    beta = dec._beta_table.detach().numpy()[:, dec._beta_index] if dec._beta_table is not None else None
    alpha = dec._alpha_table.detach().numpy()[:, dec._alpha_index] if dec._alpha_table is not None else None
    return beta, alpha
