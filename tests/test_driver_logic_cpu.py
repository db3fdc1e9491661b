"""The host driver (krylov_schur.py, decomposition.py) executed on CPU against the reference's
records, with the device calls answered by a test double (tests/fake_device.py).  This pins the
part of the drop-in that is NOT CUDA: defaults, H bookkeeping across restarts, the spike row,
the convergence test, history and the exceptions.  Parity of the kernels themselves is the
job of the -m gpu tests."""
import numpy as np
import pytest

from conftest import csr_from_golden, lap2d
from fake_device import FakeDeviceSolver


@pytest.fixture
def fake(monkeypatch):
    import arnoldi_b200.decomposition as dec
    import arnoldi_b200.krylov_schur as ks
    monkeypatch.setattr(ks, "DeviceSolver", FakeDeviceSolver)
    monkeypatch.setattr(dec, "DeviceSolver", FakeDeviceSolver)
    return ks


CASES = [
    ("mark50_s0", "mark50", 0, dict(nev=5, max_dim=20, stopping_criterion=1e-8, max_restarts=1000)),
    ("mark50_s42", "mark50", 42, dict(nev=5, max_dim=20, stopping_criterion=1e-8, max_restarts=1000)),
    ("mark10_s0", "mark10", 0, dict(nev=3, max_dim=5, max_restarts=1000)),
    ("lap2d32_s0", "lap2d32", 0, dict(nev=10, max_dim=40, stopping_criterion=1e-8, max_restarts=1000)),
    ("cplx400_s0", "cplx400", 0, dict(nev=4, max_dim=24, stopping_criterion=1e-8, max_restarts=2000)),
]


def _matrix(golden, name):
    if name.startswith("mark"):
        return csr_from_golden(golden("matrices"), name)
    if name.startswith("lap2d"):
        return lap2d(int(name[5:]))
    return csr_from_golden(golden("solves"), name)


@pytest.mark.parametrize("tag,mat,seed,kw", CASES, ids=[c[0] for c in CASES])
def test_driver_reproduces_reference_records(fake, golden, tag, mat, seed, kw):
    from arnoldi_b200.utils import arg_largest_real
    g = golden("solves")
    A = _matrix(golden, mat)
    kw = dict(kw)
    nev = kw.pop("nev")
    np.random.seed(seed)
    stats = {}
    Q, T, hist = fake.partial_schur(A, nev, sort_function=arg_largest_real, stats=stats, **kw)
    np.testing.assert_array_equal(hist.restarts, g[f"{tag}_hist_restarts"])
    np.testing.assert_array_equal(hist.matvecs, g[f"{tag}_hist_matvecs"])
    assert stats["true_matvecs"] == int(g[f"{tag}_true_matvecs"])
    np.testing.assert_allclose(np.diag(T), g[f"{tag}_diagT"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(T, g[f"{tag}_T"], rtol=1e-9, atol=1e-11)
    assert Q.shape == (A.shape[0], nev) and Q.dtype == np.complex128 and T.shape == (nev, nev)
    assert stats["restart_cycles"] == int(hist.restarts[0]) and stats["converged"] is True


def test_driver_defaults_and_mgs_plug(fake, golden):
    g = golden("solves")
    A = csr_from_golden(golden("matrices"), "mark20")
    np.random.seed(3)
    Q, T, hist = fake.partial_schur(A, 4)          # tol, sort, max_dim, p all defaulted
    np.testing.assert_array_equal(hist.restarts, g["mark20_default_restarts"])
    np.testing.assert_array_equal(hist.matvecs, g["mark20_default_matvecs"])
    np.testing.assert_allclose(np.diag(T), g["mark20_default_diagT"], rtol=1e-10, atol=1e-12)
    from arnoldi_b200.utils import arg_largest_real
    A = csr_from_golden(golden("matrices"), "mark50")
    np.random.seed(0)
    Q, T, hist = fake.partial_schur(A, 5, max_dim=20, stopping_criterion=1e-8, max_restarts=1000,
                                    sort_function=arg_largest_real, ortho="mgs")
    np.testing.assert_array_equal(hist.restarts, g["mark50_mgs_s0_hist_restarts"])
    np.testing.assert_allclose(np.diag(T), g["mark50_mgs_s0_diagT"], rtol=1e-10, atol=1e-12)


def test_driver_exceptions(fake):
    import scipy.sparse as sp
    A = lap2d(8)
    np.random.seed(0)
    with pytest.raises(ValueError, match="Has not converged !"):
        fake.partial_schur(A, 3, max_dim=8, max_restarts=1, stopping_criterion=1e-14)
    np.random.seed(0)
    Q, T, hist = fake.partial_schur(A, 3, max_dim=8, max_restarts=1, stopping_criterion=1e-14,
                                    raise_on_no_convergence=False)
    assert Q.shape == (64, 3) and not hist.restarts.any()
    D = sp.diags_array(np.arange(1.0, 31.0)).tocsr()
    v0 = np.zeros(30, np.complex128)
    v0[[3, 9]] = 1 / np.sqrt(2)
    with pytest.raises(ValueError, match="Happy breakdown not supported yet"):
        fake.partial_schur(D, 2, max_dim=10, v0=v0)


def test_arnoldi_decomposition_wrapper(fake, golden):
    """decomposition.py: in-place update of caller-owned V / H in either memory order,
    truncation, breakdown return shapes (tests/test_decomposition.py:92-139 of the reference)."""
    from arnoldi_b200.decomposition import arnoldi_decomposition
    g = golden("arnoldi")
    A = csr_from_golden(golden("matrices"), "mark10")
    n, m, md = A.shape[0], g["mark10_H"].shape[1], int(g["mark10_maxdim"])
    for order in ("F", "C"):
        V = np.zeros((n, m + 1), np.complex128, order=order)
        H = np.zeros((m + 1, m), np.complex128)
        V[:, 0] = g["mark10_v0"]
        Va, Ha, k = arnoldi_decomposition(A, V, H, 1e-8, max_dim=md)
        assert k == md and Va.shape == (n, md + 1) and Ha.shape == (md + 1, md)
        np.testing.assert_allclose(V, g["mark10_V"], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(H, g["mark10_H"], rtol=1e-12, atol=1e-13)
    Ad = g["brk_A"]
    n, m = Ad.shape[0], g["brk_H"].shape[1]
    V = np.zeros((n, m + 1), np.complex128, order="F")
    H = np.zeros((m + 1, m), np.complex128)
    V[:, 0] = g["brk_v0"]
    Va, Ha, k = arnoldi_decomposition(Ad, V, H, 1e-8)
    assert k == 1 and Va.shape == (n, 2) and Ha.shape == (2, 1) and H[1, 0] == 0
    np.testing.assert_allclose(H, g["brk_H"], rtol=1e-12, atol=1e-13)
