"""Pin the oracle (oracle/krylov.py) to the reference's own outputs.

The goldens were written by oracle/make_golden.py from the unmodified reference.
On the machine that generated them the oracle reproduces them bit for bit (same
BLAS); tolerances below are a few ulp so a different host BLAS still passes.
"""
import numpy as np
import pytest

import oracle
from conftest import csr_from_golden, lap2d

TIGHT = dict(rtol=1e-12, atol=1e-13)


def test_cgs_and_mgs_match_reference(golden):
    g = golden("ortho")
    for name in g["names"]:
        V = np.asfortranarray(g[f"{name.split('_')[0]}_V"])
        for tag, fn in (("gs", oracle.cgs_dgks), ("mgs", oracle.mgs_dgks)):
            w = g[f"{name}_w"].copy()
            h = np.zeros(V.shape[1], np.complex128)
            beta, brk = fn(w, V, h, 1e-8)
            assert bool(brk) == bool(g[f"{name}_{tag}_brk"]), (name, tag)
            np.testing.assert_allclose(h, g[f"{name}_{tag}_h"], **TIGHT)
            np.testing.assert_allclose(w, g[f"{name}_{tag}_w"], rtol=1e-9, atol=1e-13)
            np.testing.assert_allclose(beta, g[f"{name}_{tag}_beta"], rtol=1e-9, atol=1e-13)
    # the three scenarios do what their names say
    assert not g["mid_gs_brk"] and g["mid_brk_gs_brk"]


def test_arnoldi_expand_matches_reference(golden):
    g = golden("arnoldi")
    A = csr_from_golden(g, "cplx")
    n, m = A.shape[0], g["cplx_H"].shape[1]
    V = np.zeros((n, m + 1), np.complex128, order="F")
    H = np.zeros((m + 1, m), np.complex128)
    V[:, 0] = g["cplx_v0"]
    _, _, it = oracle.arnoldi_expand(A, V, H, 1e-8)
    assert it == int(g["cplx_niter"])
    np.testing.assert_allclose(V, g["cplx_V"], **TIGHT)
    np.testing.assert_allclose(H, g["cplx_H"], **TIGHT)

    A = csr_from_golden(golden("matrices"), "mark10")
    n, m, md = A.shape[0], g["mark10_H"].shape[1], int(g["mark10_maxdim"])
    V = np.zeros((n, m + 1), np.complex128, order="F")
    H = np.zeros((m + 1, m), np.complex128)
    V[:, 0] = g["mark10_v0"]
    Va, Ha, it = oracle.arnoldi_expand(A, V, H, 1e-8, max_dim=md)
    assert it == md and Va.shape == (n, md + 1) and Ha.shape == (md + 1, md)
    np.testing.assert_allclose(V, g["mark10_V"], **TIGHT)
    np.testing.assert_allclose(H, g["mark10_H"], **TIGHT)

    Ad = g["brk_A"]
    n, m = Ad.shape[0], g["brk_H"].shape[1]
    V = np.zeros((n, m + 1), np.complex128, order="F")
    H = np.zeros((m + 1, m), np.complex128)
    V[:, 0] = g["brk_v0"]
    Va, Ha, it = oracle.arnoldi_expand(Ad, V, H, 1e-8)
    assert it == int(g["brk_niter"]) == 1
    assert Va.shape == (n, 2) and Ha.shape == (2, 1)
    np.testing.assert_allclose(H, g["brk_H"], **TIGHT)


def test_sorted_schur_and_restart_match_reference(golden):
    from scipy.linalg import schur
    g = golden("restart")
    T1, Q1 = schur(g["Hm"], output="complex")
    T2, Q2 = oracle.sorted_schur(T1, oracle.arg_largest_real)
    np.testing.assert_allclose(T2, g["T2"], **TIGHT)
    np.testing.assert_allclose(Q2, g["Q2"], **TIGHT)
    d = np.diag(T2).real
    assert np.all(d[:-1] >= d[1:])
    m, p = int(g["m"]), int(g["p"])
    V = np.asfortranarray(g["V"].copy())
    oracle.restart_update(V, g["Q"], m, p)
    np.testing.assert_allclose(V, g["Vout"], **TIGHT)


SOLVES = [
    ("mark50_s0", "mark50", 0, dict(nev=5, max_dim=20, stopping_criterion=1e-8, max_restarts=1000)),
    ("mark50_s1", "mark50", 1, dict(nev=5, max_dim=20, stopping_criterion=1e-8, max_restarts=1000)),
    ("mark50_s42", "mark50", 42, dict(nev=5, max_dim=20, stopping_criterion=1e-8, max_restarts=1000)),
    ("mark10_s0", "mark10", 0, dict(nev=3, max_dim=5, max_restarts=1000)),
    ("lap2d32_s0", "lap2d32", 0, dict(nev=10, max_dim=40, stopping_criterion=1e-8, max_restarts=1000)),
    ("cplx400_s0", "cplx400", 0, dict(nev=4, max_dim=24, stopping_criterion=1e-8, max_restarts=2000)),
    ("mark100_s0", "mark100", 0, dict(nev=20, max_dim=60, stopping_criterion=1e-8, max_restarts=1000)),
    ("lap2d64_s0", "lap2d64", 0, dict(nev=10, max_dim=40, stopping_criterion=1e-8, max_restarts=1000)),
]


def _matrix(golden, name):
    if name.startswith("mark"):
        if f"{name}_shape" not in golden("matrices"):
            from arnoldi_b200.matrices import mark   # bit-identical to the reference's mark()
            return mark(int(name[4:]))
        return csr_from_golden(golden("matrices"), name)
    if name.startswith("lap2d"):
        return lap2d(int(name[5:]))
    return csr_from_golden(golden("solves"), name)


@pytest.mark.parametrize("tag,mat,seed,kw", SOLVES, ids=[s[0] for s in SOLVES])
def test_partial_schur_matches_reference(golden, tag, mat, seed, kw):
    g = golden("solves")
    A = _matrix(golden, mat)
    counters = {}
    np.random.seed(seed)
    kw = dict(kw)
    Q, T, hist = oracle.partial_schur(A, kw.pop("nev"), sort_function=oracle.arg_largest_real,
                                      counters=counters, **kw)
    np.testing.assert_array_equal(hist.restarts, g[f"{tag}_hist_restarts"])
    np.testing.assert_array_equal(hist.matvecs, g[f"{tag}_hist_matvecs"])
    assert counters["matvecs"] == int(g[f"{tag}_true_matvecs"])
    np.testing.assert_allclose(np.diag(T), g[f"{tag}_diagT"], rtol=1e-10, atol=1e-12)


def test_partial_schur_mgs_plug_matches_reference(golden):
    g = golden("solves")
    A = _matrix(golden, "mark50")
    np.random.seed(0)
    Q, T, hist = oracle.partial_schur(A, 5, max_dim=20, stopping_criterion=1e-8,
                                      max_restarts=1000, sort_function=oracle.arg_largest_real,
                                      ortho=oracle.mgs_dgks)
    np.testing.assert_array_equal(hist.restarts, g["mark50_mgs_s0_hist_restarts"])
    np.testing.assert_allclose(np.diag(T), g["mark50_mgs_s0_diagT"], rtol=1e-10, atol=1e-12)


def test_partial_schur_defaults_match_reference(golden):
    # defaults: tol=sqrt(eps), largest magnitude, max_dim=max(2k+1,20), p=k+5
    g = golden("solves")
    A = csr_from_golden(golden("matrices"), "mark20")
    np.random.seed(3)
    Q, T, hist = oracle.partial_schur(A, 4)
    np.testing.assert_array_equal(hist.restarts, g["mark20_default_restarts"])
    np.testing.assert_allclose(np.diag(T), g["mark20_default_diagT"], rtol=1e-10, atol=1e-12)


def test_errors_match_reference():
    A = lap2d(8)
    with pytest.raises(ValueError, match="Has not converged"):
        np.random.seed(0)
        oracle.partial_schur(A, 3, max_dim=8, max_restarts=1, stopping_criterion=1e-14)
    with pytest.raises(AssertionError):
        oracle.partial_schur(A, 3, max_dim=3)
    with pytest.raises(AssertionError):
        oracle.partial_schur(A, 3, max_restarts=0)
