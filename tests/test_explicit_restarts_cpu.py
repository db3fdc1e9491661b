"""explicit_restarts_with_deflation (explicit_restarts.py:80-168): the oracle's restatement and
the host driver (through the device test double) against records of the unmodified reference
(tests/golden/explicit.npz, oracle/make_golden_r2.py)."""
import numpy as np
import pytest

import oracle
from conftest import csr_from_golden
from fake_device import FakeDeviceSolver

CASES = [("mark10", dict(nev=3, max_dim=10, stopping_criterion=1e-8)),
         ("mark20", dict(nev=4, max_dim=20, stopping_criterion=1e-8, max_restarts=400)),
         ("rect12", dict(nev=4, max_dim=24, stopping_criterion=1e-9, max_restarts=400))]


def _match(vals, ref):
    """Hungarian-free matching is enough here: both are sorted by the same key."""
    a = vals[np.argsort(-vals.real)]
    b = ref[np.argsort(-ref.real)]
    return np.abs(a - b) / np.abs(b)


@pytest.mark.parametrize("tag,kw", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference_record(golden, tag, kw):
    g = golden("explicit")
    A = csr_from_golden(g, tag)
    kw = dict(kw)
    nev = kw.pop("nev")
    np.random.seed(0)
    vals, vecs, hist = oracle.explicit_restarts_with_deflation(
        A, nev, sort_function=oracle.arg_largest_real, **kw)
    np.testing.assert_array_equal(hist.restarts, g[f"{tag}_hist_restarts"])
    np.testing.assert_array_equal(hist.matvecs, g[f"{tag}_hist_matvecs"])
    np.testing.assert_allclose(vals, g[f"{tag}_vals"], rtol=1e-12, atol=1e-14)
    assert np.linalg.norm(A @ vecs - vals * vecs, axis=0).max() < 1e-7


@pytest.mark.parametrize("tag,kw", CASES, ids=[c[0] for c in CASES])
def test_driver_matches_reference_record(monkeypatch, golden, tag, kw):
    import arnoldi_b200.explicit_restarts as er
    monkeypatch.setattr(er, "DeviceSolver", FakeDeviceSolver)
    from arnoldi_b200.utils import arg_largest_real
    g = golden("explicit")
    A = csr_from_golden(g, tag)
    kw = dict(kw)
    nev = kw.pop("nev")
    np.random.seed(0)
    vals, vecs, hist = er.explicit_restarts_with_deflation(A, nev, sort_function=arg_largest_real, **kw)
    np.testing.assert_array_equal(hist.restarts, g[f"{tag}_hist_restarts"])
    np.testing.assert_array_equal(hist.matvecs, g[f"{tag}_hist_matvecs"])
    assert _match(vals, g[f"{tag}_vals"]).max() < 1e-10
    assert vecs.shape == (A.shape[0], nev)
    assert np.linalg.norm(A @ vecs - vals * vecs, axis=0).max() < 1e-7


def test_driver_reports_non_convergence(monkeypatch):
    """tests/test_explicit_restarts.py:142-158 of the reference."""
    import arnoldi_b200.explicit_restarts as er
    from arnoldi_b200.matrices import mark
    monkeypatch.setattr(er, "DeviceSolver", FakeDeviceSolver)
    with pytest.raises(ValueError, match="Could not converge for value 0"):
        er.explicit_restarts_with_deflation(mark(10), 3, max_dim=5, stopping_criterion=1e-16,
                                            max_restarts=10)
