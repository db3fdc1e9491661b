"""pytest configuration: markers, import paths, shared helpers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "arnoldi-py_b200")
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def gpu():
    """GPU tests never skip silently: without the library or a device they FAIL."""
    from arnoldi_b200 import _lib
    lib = _lib.load()
    n = lib.ab200_device_count()
    if n < 1:
        pytest.fail("no CUDA device visible: " + lib.ab200_last_error().decode(), pytrace=False)
    return lib


def csr_from_golden(g, prefix):
    import scipy.sparse as sp
    shape = tuple(int(x) for x in g[f"{prefix}_shape"])
    return sp.csr_matrix((g[f"{prefix}_data"], g[f"{prefix}_indices"], g[f"{prefix}_indptr"]),
                         shape=shape)


def lap2d(N):
    """kron form used when the goldens were generated: CSR WITH explicit zeros."""
    import scipy.sparse as sp
    T = sp.diags_array([-np.ones(N - 1), 2 * np.ones(N), -np.ones(N - 1)], offsets=[-1, 0, 1])
    I = sp.eye_array(N)
    return (sp.kron(I, T) + sp.kron(T, I)).tocsr()
