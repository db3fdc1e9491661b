"""GPU parity of the round-2 additions, through the C ABI: the bulk-copy (TMA) SpMV pipeline in
every shape it can take, spmv_algo, combine / orthonormalize_column / project, the driver
features beyond the reference (real arithmetic with pairs kept whole, breakdown repair,
locking, dynamic p), operator wrappers and device operators, wide bases (max_dim > 128), and the
round-2 records of the reference (non-degenerate symmetric operator, mark(200))."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from conftest import csr_from_golden

pytestmark = pytest.mark.gpu


def _cvec(rng, n):
    return rng.standard_normal(n) + 1j * rng.standard_normal(n)


def _short_row_cases():
    from arnoldi_b200.matrices import lap2d, lap2d_rect, mark
    rng = np.random.default_rng(21)
    cases = {"mark60": mark(60), "lap2d_90": lap2d(90), "rect": lap2d_rect(50, 71)}
    A = sp.random(20000, 20000, density=3.0 / 20000, random_state=rng, format="csr")
    cases["random_f64"] = A
    Ac = A.astype(np.complex128)
    Ac.data = Ac.data + 1j * rng.standard_normal(Ac.nnz)
    cases["random_c128"] = Ac
    # long runs of empty rows: more rows in a tile than row pointers are staged for
    B = sp.random(40000, 40000, density=0.5 / 40000, random_state=rng, format="csr")
    cases["mostly_empty"] = B
    cases["all_empty"] = sp.csr_matrix((3000, 3000), dtype=np.float64)
    cases["one"] = sp.csr_matrix(np.array([[2.5]]))
    # rows of exactly 16 entries (the longest the one-thread-per-row kernels accept)
    n = 5000
    idx = (np.arange(n)[:, None] + rng.integers(-40, 40, (n, 16))) % n
    C = sp.csr_matrix((rng.standard_normal(16 * n), idx.ravel().astype(np.int32),
                       np.arange(0, 16 * n + 1, 16).astype(np.int32)), shape=(n, n))
    cases["len16_unsorted"] = C
    return cases


SHAPES = [dict(), dict(spmv_threads=128), dict(spmv_stages=2), dict(spmv_stages=5, spmv_tile=512),
          dict(spmv_tile=2560, spmv_bps=2), dict(spmv_tile=520), dict(spmv_variant=1),
          dict(spmv_variant=2, spmv_threads=128)]


@pytest.mark.parametrize("name", list(_short_row_cases()))
def test_spmv_pipeline_shapes_bit_exact(gpu, name):
    """Every SpMV kernel for short rows, in several pipeline shapes, complex and real vectors,
    int32 and int64 row pointers: bit-identical to scipy's csr_matvec."""
    from arnoldi_b200.solver import DeviceSolver
    A = _short_row_cases()[name]
    n = A.shape[0]
    assert np.diff(A.indptr).max(initial=0) <= 16
    rng = np.random.default_rng(5)
    xs = [_cvec(rng, n), rng.standard_normal(n).astype(np.complex128)]
    refs = [A @ x for x in xs]
    for opts in SHAPES:
        for ip in (A.indptr, A.indptr.astype(np.int64)):
            with DeviceSolver(n, 2) as dev:
                for k, v in opts.items():
                    dev.set_option(k, v)
                dev.set_csr(ip, A.indices, A.data, algo="stream")
                for x, ref in zip(xs, refs):
                    np.testing.assert_array_equal(dev.spmv(x), ref, err_msg=f"{name} {opts}")


def test_spmv_algo_selects_kernels(gpu):
    """spmv_algo is honoured: STREAM refuses long rows, VECTOR / MERGE / AUTO handle them."""
    from arnoldi_b200.matrices import lap2d, powerlaw
    from arnoldi_b200.solver import DeviceSolver
    A = powerlaw(30000)
    n = A.shape[0]
    x = _cvec(np.random.default_rng(1), n)
    ref = A @ x
    scale = abs(A.copy()) @ np.abs(x)
    with DeviceSolver(n, 2) as dev:
        with pytest.raises(AssertionError, match="AB200_SPMV_STREAM"):
            dev.set_csr(A.indptr, A.indices, A.data, algo="stream")
        for algo in ("auto", "vector", "merge"):
            dev.set_csr(A.indptr, A.indices, A.data, algo=algo)
            y = dev.spmv(x)
            assert np.all(np.abs(y - ref) <= 1e-13 * scale + 1e-300), algo
    B = lap2d(64)
    xb = _cvec(np.random.default_rng(2), B.shape[0])
    with DeviceSolver(B.shape[0], 2) as dev:
        for algo in ("auto", "stream", "vector", "merge"):
            dev.set_csr(B.indptr, B.indices, B.data, algo=algo)
            y = dev.spmv(xb)
            if algo in ("auto", "stream", "vector"):
                np.testing.assert_array_equal(y, B @ xb)
            else:
                np.testing.assert_allclose(y, B @ xb, rtol=0, atol=1e-13 * 8 * np.abs(xb).max())


def test_combine_orthonormalize_project(gpu):
    """The three helper entry points against NumPy, complex and real storage, including a wide
    basis (more than 128 columns: pass 1 runs in column groups)."""
    from arnoldi_b200.solver import DeviceSolver
    rng = np.random.default_rng(8)
    n = 4099
    A = sp.random(n, n, density=4.0 / n, random_state=rng, format="csr") + sp.eye_array(n)
    A = A.tocsr()
    for real in (False, True):
        for md in (12, 150):
            M = rng.standard_normal((n, md)) + (0 if real else 1j) * rng.standard_normal((n, md))
            V, _ = np.linalg.qr(M)
            V = np.asfortranarray(V.astype(np.complex128))
            with DeviceSolver(n, md) as dev:
                dev.set_csr(A.indptr, A.indices, A.data)
                dev.set_columns(0, V)
                assert dev.stats()["real_storage"] == (1 if real else 0)
                # project: h = V[:, :k]^H A V[:, c]
                k, c = md - 1, 3
                h = dev.project(c, k)
                ref = V[:, :k].conj().T @ (A @ V[:, c])
                np.testing.assert_allclose(h, ref, rtol=0, atol=1e-12 * np.abs(ref).max())
                # orthonormalize: a fresh column against the first k
                w = rng.standard_normal(n) + (0 if real else 1j) * rng.standard_normal(n)
                w = w.astype(np.complex128)
                dev.set_columns(k, w)
                beta = dev.orthonormalize_column(k, k, 1e-8)
                wr = w - V[:, :k] @ (V[:, :k].conj().T @ w)
                wr -= V[:, :k] @ (V[:, :k].conj().T @ wr)
                assert abs(beta - np.linalg.norm(wr)) <= 1e-12 * np.linalg.norm(w)
                got = dev.get_columns(k, 1)[:, 0]
                np.testing.assert_allclose(got, wr / np.linalg.norm(wr), rtol=0, atol=1e-12)
                assert np.abs(V[:, :k].conj().T @ got).max() < 1e-13
                # a column inside the span: breakdown reported, column left alone
                dev.set_columns(k, V[:, :3] @ np.array([1.0, -2.0, 0.5]))
                assert dev.orthonormalize_column(k, k, 1e-8) < 1e-8
                # no columns to project against: plain normalisation
                dev.set_columns(0, 3.0 * V[:, 0])
                assert abs(dev.orthonormalize_column(0, 0, 1e-8) - 3.0) < 1e-12
                np.testing.assert_allclose(dev.get_columns(0, 1)[:, 0], V[:, 0], rtol=0, atol=1e-14)
                # combine: columns [2, 2+m) -> p linear combinations, nothing else touched
                m, p = 7, 3
                Qc = rng.standard_normal((m, p)) + (0 if real else 1j) * rng.standard_normal((m, p))
                before = dev.get_columns(0, md)
                dev.combine(Qc, 2, m, p)
                after = dev.get_columns(0, md)
                np.testing.assert_allclose(after[:, 2:2 + p], before[:, 2:2 + m] @ Qc, rtol=0, atol=1e-12)
                np.testing.assert_array_equal(after[:, :2], before[:, :2])
                np.testing.assert_array_equal(after[:, 2 + p:], before[:, 2 + p:])
                # p == m (a full rotation of the block, explicit_restarts.py:167)
                Qs, _ = np.linalg.qr(rng.standard_normal((5, 5)) + 1j * rng.standard_normal((5, 5)))
                dev.combine(Qs, 1, 5, 5)
                np.testing.assert_allclose(dev.get_columns(1, 5), after[:, 1:6] @ Qs, rtol=0, atol=1e-12)


def test_large_download_path(gpu):
    """get_columns above the bounce-buffer threshold, both storage modes."""
    from arnoldi_b200.solver import DeviceSolver
    rng = np.random.default_rng(2)
    n = 700001
    for real in (True, False):
        V = rng.standard_normal((n, 3)) + (0 if real else 1j) * rng.standard_normal((n, 3))
        V = np.asfortranarray(V.astype(np.complex128))
        with DeviceSolver(n, 4) as dev:
            dev.set_columns(0, V)
            np.testing.assert_array_equal(dev.get_columns(0, 3), V)


R2 = [("rect32_s0", "rect32", 0, 10, 40), ("rect32_s1", "rect32", 1, 10, 40),
      ("rect64_s0", "rect64", 0, 10, 40), ("rect64_s1", "rect64", 1, 10, 40),
      ("mark200_s0", "mark200", 0, 20, 60)]


@pytest.mark.parametrize("tag,mat,seed,nev,md", R2, ids=[c[0] for c in R2])
def test_partial_schur_round2_records(gpu, golden, tag, mat, seed, nev, md):
    """Operators with simple spectra: IDENTICAL restart count, history and matvec count, every
    Ritz value to 1e-10 relative but the last-converged pair (bounded by the two residuals)."""
    from arnoldi_b200 import partial_schur
    from arnoldi_b200.matrices import mark
    from arnoldi_b200.utils import arg_largest_real
    g = golden("solves_r2")
    A = mark(200) if mat == "mark200" else csr_from_golden(g, mat)
    np.random.seed(seed)
    stats = {}
    Q, T, hist = partial_schur(A, nev, max_dim=md, stopping_criterion=1e-8, max_restarts=2000,
                               sort_function=arg_largest_real, stats=stats)
    R, Rref = int(hist.restarts[0]), int(g[f"{tag}_hist_restarts"][0])
    lam, ref = np.diag(T), g[f"{tag}_diagT"]
    rel = np.abs(lam - ref) / np.abs(ref)
    print(f"{tag}: R={R} (reference {Rref}) max rel {rel.max():.2e} real_storage={stats['real_storage']}")
    assert R == Rref
    np.testing.assert_array_equal(hist.matvecs, g[f"{tag}_hist_matvecs"])
    assert stats["true_matvecs"] == int(g[f"{tag}_true_matvecs"])
    assert np.sum(rel > 1e-10) <= max(1, nev // 10) and rel.max() < 1e-8, rel
    if mat.startswith("rect"):
        assert rel.max() < 1e-10 and stats["real_storage"] == 1     # symmetric: real throughout
    assert np.linalg.norm(A @ Q - Q @ T, axis=0).max() < 1e-7
    assert np.abs(Q.conj().T @ Q - np.eye(nev)).max() < 1e-12


def test_real_arith_pairs_stays_real(gpu, golden):
    """real_arith="pairs" on nonsymmetric real operators: float64 storage until the final
    rotation, the reference's converged Ritz values, residuals <= tol; restart counts stated."""
    from arnoldi_b200 import partial_schur
    from arnoldi_b200.matrices import mark
    from arnoldi_b200.utils import arg_largest_real
    g = golden("solves")
    g2 = golden("solves_r2")
    for tag, gg, m, nev, md in (("mark50_s0", g, 50, 5, 20), ("mark100_s0", g, 100, 20, 60),
                                ("mark200_s0", g2, 200, 20, 60)):
        A = mark(m)
        out = {}
        for mode in ("pairs", "lossless"):
            np.random.seed(0)
            stats = {}
            Q, T, hist = partial_schur(A, nev, max_dim=md, stopping_criterion=1e-8,
                                       max_restarts=2000, sort_function=arg_largest_real,
                                       stats=stats, real_arith=mode)
            out[mode] = (np.diag(T), int(hist.restarts[0]), stats)
            assert np.linalg.norm(A @ Q - Q @ T, axis=0).max() < 1e-7
            assert np.abs(Q.conj().T @ Q - np.eye(nev)).max() < 1e-12
            assert np.abs(np.tril(T, -1)).max() == 0
        ref = np.sort_complex(gg[f"{tag}_diagT"])
        for mode in out:
            rel = np.abs(np.sort_complex(out[mode][0]) - ref) / np.abs(ref)
            assert np.sum(rel > 1e-10) <= max(1, nev // 10) and rel.max() < 1e-8, (tag, mode, rel)
        sp_, sl = out["pairs"][2], out["lossless"][2]
        print(f"{tag}: pairs R={out['pairs'][1]} matvecs={sp_['true_matvecs']} kept whole "
              f"{sp_['pairs_kept_whole']} | lossless R={out['lossless'][1]} matvecs="
              f"{sl['true_matvecs']} | reference R={int(gg[f'{tag}_hist_restarts'][0])}")
        assert out["lossless"][1] == int(gg[f"{tag}_hist_restarts"][0])
        # bytes moved by the orthogonalisation per matvec: real storage moves about half
        bp = (sp_["ortho_pass1_bytes"] + sp_["ortho_pass2_bytes"] + sp_["ortho_fused_bytes"]) / sp_["true_matvecs"]
        bl = (sl["ortho_pass1_bytes"] + sl["ortho_pass2_bytes"] + sl["ortho_fused_bytes"]) / sl["true_matvecs"]
        assert bp < 0.62 * bl, (bp, bl)


def test_breakdown_repair_on_device(gpu):
    from arnoldi_b200 import partial_schur
    from arnoldi_b200.utils import arg_largest_real
    d = np.arange(1.0, 2001.0)
    D = sp.diags_array(d).tocsr()
    v0 = np.zeros(2000, np.complex128)
    v0[[3, 9, 77]] = 1 / np.sqrt(3)
    with pytest.raises(ValueError, match="Happy breakdown not supported yet"):
        partial_schur(D, 3, max_dim=12, v0=v0, sort_function=arg_largest_real)
    for ortho in ("cgs2", "mgs"):
        np.random.seed(5)
        stats = {}
        Q, T, hist = partial_schur(D, 3, max_dim=24, v0=v0, sort_function=arg_largest_real,
                                   on_breakdown="continue", max_restarts=3000, stats=stats,
                                   stopping_criterion=1e-9, ortho=ortho)
        np.testing.assert_allclose(np.sort(np.diag(T).real)[::-1], [2000.0, 1999.0, 1998.0], rtol=1e-8)
        assert np.linalg.norm(D @ Q - Q @ T, axis=0).max() < 1e-5
        assert stats["breakdowns_repaired"] >= 1
    # whole space exhausted: n == max_dim, repeated eigenvalue (tests/test_krylov_schur.py:28-49)
    dd = np.array([7.0, 7.0, 5.0, 4.0, 3.0, 2.0, 1.0])
    rng = np.random.default_rng(0)
    Qm, _ = np.linalg.qr(rng.standard_normal((7, 7)))
    A = Qm.T @ np.diag(dd) @ Qm
    np.random.seed(1)
    Q, T, hist = partial_schur(A, 3, max_dim=7, sort_function=arg_largest_real,
                               on_breakdown="continue", max_restarts=50)
    np.testing.assert_allclose(np.sort(np.diag(T).real)[::-1], [7.0, 7.0, 5.0], atol=1e-7)
    assert np.linalg.norm(A @ Q - Q @ T, axis=0).max() < 1e-7


def test_lock_and_dynamic_p_on_device(gpu, golden):
    from arnoldi_b200 import partial_schur
    from arnoldi_b200.matrices import mark
    from arnoldi_b200.utils import arg_largest_real
    g = golden("solves")
    A = mark(50)
    ref = g["mark50_s0_diagT"]
    for kw in (dict(lock=True), dict(dynamic_p=True), dict(lock=True, dynamic_p=True, real_arith="pairs")):
        np.random.seed(0)
        stats = {}
        Q, T, hist = partial_schur(A, 5, max_dim=20, stopping_criterion=1e-8, max_restarts=1000,
                                   sort_function=arg_largest_real, stats=stats, **kw)
        rel = np.abs(np.sort_complex(np.diag(T)) - np.sort_complex(ref)) / np.abs(ref)
        assert rel.max() < 1e-8, (kw, rel)
        assert np.linalg.norm(A @ Q - Q @ T, axis=0).max() < 1e-7
        assert stats["true_matvecs"] <= 260


def test_reference_harness_runs_unmodified(gpu, golden):
    """scripts/utils.py:161-187 of the reference (`arnoldi_py_eig`), verbatim in structure, against
    the `arnoldi` alias package: MatvecCounter(LinearOperator) wrapper, perf_counter around
    partial_schur, eig(T), Q @ S."""
    import time

    from scipy.sparse.linalg import LinearOperator

    from arnoldi import partial_schur                      # the alias package -> B200 path
    from arnoldi.matrices import mark
    from arnoldi.utils import arg_largest_real

    class MatvecCounter(LinearOperator):                   # scripts/utils.py:55-68
        def __init__(self, A):
            self.A = A
            self.shape = A.shape
            self.dtype = np.dtype(A.dtype)
            self.matvecs = 0

        def _matvec(self, x):
            self.matvecs += 1
            return self.A @ x

    g = golden("solves")
    A0 = mark(50)
    A = MatvecCounter(A0)
    np.random.seed(0)
    t0 = time.perf_counter()
    Q, T, history = partial_schur(A, 5, max_dim=20, stopping_criterion=1e-8, max_restarts=1000,
                                  sort_function=arg_largest_real, p=None)
    elapsed = time.perf_counter() - t0
    vals, S = np.linalg.eig(T)
    vecs = Q @ S
    idx = arg_largest_real(vals)
    vals, vecs = vals[idx], vecs[:, idx]
    assert elapsed > 0
    assert int(np.max(history.restarts)) == int(g["mark50_s0_hist_restarts"][0])
    assert A.matvecs == int(g["mark50_s0_true_matvecs"])
    res = np.linalg.norm(A0 @ vecs - vecs * vals, axis=0) / np.abs(vals)
    assert res.max() < 1e-7
    np.testing.assert_allclose(np.sort(vals.real), np.sort(g["mark50_s0_eig_vals"].real), rtol=1e-9)


def test_device_operator_matches_csr_path(gpu):
    """A matrix-free operator (torch ops on the solver's own buffers and stream) gives the same
    solve as the CSR operator it implements; the real operator keeps float64 storage."""
    import torch

    from arnoldi_b200 import partial_schur
    from arnoldi_b200.device_operator import TorchOperator
    from arnoldi_b200.matrices import lap2d_rect
    from arnoldi_b200.utils import arg_largest_real
    nx, ny = 40, 41
    A = lap2d_rect(nx, ny)
    n = A.shape[0]

    def stencil(x):                      # the same 5-point operator, matrix-free
        g = x.view(ny, nx)
        y = 3.5 * g
        y[:, 1:] -= g[:, :-1]
        y[:, :-1] -= g[:, 1:]
        y[1:, :] -= 0.75 * g[:-1, :]
        y[:-1, :] -= 0.75 * g[1:, :]
        return y.reshape(-1)

    op = TorchOperator(stencil, n, np.float64)
    out = {}
    for name, M in (("csr", A), ("op", op)):
        np.random.seed(0)
        stats = {}
        Q, T, hist = partial_schur(M, 6, max_dim=24, stopping_criterion=1e-9, max_restarts=2000,
                                   sort_function=arg_largest_real, stats=stats)
        out[name] = (np.diag(T), int(hist.restarts[0]), stats["real_storage"])
        assert np.linalg.norm(A @ Q - Q @ T, axis=0).max() < 1e-7
    assert op.calls == out["op"][2] * 0 + out["op"][1] * 0 + op.calls and op.calls > 0
    np.testing.assert_allclose(out["op"][0], out["csr"][0], rtol=1e-10)
    assert abs(out["op"][1] - out["csr"][1]) <= 2 and out["op"][2] == 1
    torch.cuda.synchronize()
    # a complex operator through the same protocol
    rng = np.random.default_rng(4)
    dvec = np.linspace(1, 3, 500) + 0.3j * np.linspace(-1, 1, 500)
    B = (sp.random(500, 500, density=0.01, random_state=rng) * (0.2 + 0.1j) + sp.diags_array(dvec)).tocsr()
    Bt = torch.tensor(B.toarray(), device="cuda")
    opc = TorchOperator(lambda x: Bt @ x, 500, np.complex128)
    res = {}
    for name, M in (("csr", B), ("op", opc)):
        np.random.seed(0)
        Q, T, hist = partial_schur(M, 4, max_dim=24, stopping_criterion=1e-9, max_restarts=2000,
                                   sort_function=arg_largest_real)
        res[name] = np.diag(T)
        assert np.linalg.norm(B @ Q - Q @ T, axis=0).max() < 1e-6
    np.testing.assert_allclose(res["op"], res["csr"], rtol=1e-9)


def test_wide_basis_solve(gpu):
    """max_dim above 128 (reference default for nev >= 64: max_dim = 2 nev + 1): pass 1 runs in
    column groups, the fused sweep steps aside, the restart kernel takes its wide path."""
    from arnoldi_b200 import partial_schur
    from arnoldi_b200.matrices import lap2d_rect
    from arnoldi_b200.utils import arg_largest_real
    A = lap2d_rect(30, 37)
    for ortho in ("cgs2", "mgs"):
        np.random.seed(0)
        stats = {}
        Q, T, hist = partial_schur(A, 70, sort_function=arg_largest_real, stopping_criterion=1e-8,
                                   max_restarts=500, stats=stats, ortho=ortho)   # max_dim = 141
        np.random.seed(0)
        Qo, To, ho = oracle.partial_schur(A, 70, sort_function=oracle.arg_largest_real,
                                          stopping_criterion=1e-8, max_restarts=500)
        rel = np.abs(np.diag(T) - np.diag(To)) / np.abs(np.diag(To))
        assert np.sum(rel > 1e-10) <= 7 and rel.max() < 1e-7, rel
        assert abs(int(hist.restarts[0]) - int(ho.restarts[0])) <= 2
        assert np.linalg.norm(A @ Q - Q @ T, axis=0).max() < 1e-6
        assert np.abs(Q.conj().T @ Q - np.eye(70)).max() < 1e-11
    with pytest.raises(ValueError, match="limit of 256"):
        partial_schur(sp.eye_array(1000, format="csr"), 130)


EXPLICIT = [("mark10", dict(nev=3, max_dim=10, stopping_criterion=1e-8)),
            ("mark20", dict(nev=4, max_dim=20, stopping_criterion=1e-8, max_restarts=400)),
            ("rect12", dict(nev=4, max_dim=24, stopping_criterion=1e-9, max_restarts=400))]


@pytest.mark.parametrize("tag,kw", EXPLICIT, ids=[c[0] for c in EXPLICIT])
def test_explicit_restarts_with_deflation(gpu, golden, tag, kw):
    """explicit_restarts.py:80-168 on the device kernels against the reference's record: same
    eigenvalues, residuals below the reference's test threshold, per-pair restart counts equal
    or within one (the estimate that stops a pair sits at tol; summation order differs)."""
    from arnoldi.explicit_restarts import explicit_restarts_with_deflation   # alias package
    from arnoldi_b200.utils import arg_largest_real
    g = golden("explicit")
    A = csr_from_golden(g, tag)
    kw = dict(kw)
    nev = kw.pop("nev")
    np.random.seed(0)
    stats = {}
    vals, vecs, hist = explicit_restarts_with_deflation(A, nev, sort_function=arg_largest_real,
                                                        stats=stats, **kw)
    ref = g[f"{tag}_vals"]
    a, b = vals[np.argsort(-vals.real)], ref[np.argsort(-ref.real)]
    assert (np.abs(a - b) / np.abs(b)).max() < 1e-7
    res = np.linalg.norm(A @ vecs - vals * vecs, axis=0)
    np.testing.assert_allclose(res, 0, rtol=1e-4, atol=1e-8)       # the reference's own criterion
    print(tag, "restarts", hist.restarts, "reference", g[f"{tag}_hist_restarts"])
    assert np.abs(hist.restarts.astype(int) - g[f"{tag}_hist_restarts"].astype(int)).max() <= 1
    assert stats["mgs_launches"] > 0 and stats["restart_launches"] > 0
    from arnoldi_b200.matrices import mark
    with pytest.raises(ValueError, match="Could not converge for value 0"):
        explicit_restarts_with_deflation(mark(10), 3, max_dim=5, stopping_criterion=1e-16,
                                         max_restarts=10)


def _local_skewed(n, rng, long_rows=()):
    """Power-law row lengths with columns in a band around the diagonal plus a few far ones,
    unsorted, with duplicates; optional very long rows and runs of empty rows."""
    lens = np.minimum((rng.pareto(1.3, n) * 4).astype(np.int64) + 1, 700)
    lens[100:400] = 0
    for r, l in long_rows:
        lens[r] = l
    indptr = np.concatenate(([0], np.cumsum(lens)))
    nnz = int(indptr[-1])
    rows = np.repeat(np.arange(n), lens)
    off = rng.integers(-3000, 3001, nnz)
    far = rng.random(nnz) < 0.04
    cols = np.where(far, rng.integers(0, n, nnz), (rows + off) % n)
    return sp.csr_matrix((rng.uniform(-1, 1, nnz), cols.astype(np.int32), indptr.astype(np.int32)),
                         shape=(n, n))


@pytest.mark.parametrize("opts", [dict(), dict(spmv_tile=1024), dict(spmv_window=2048),
                                  dict(spmv_tile=8192, spmv_bps=2),
                                  dict(spmv_window=1024, spmv_ring_warps=3, spmv_tile=256),
                                  dict(spmv_win_half=100), dict(spmv_tile=640, spmv_ring_warps=1),
                                  dict(spmv_tile=256, spmv_ring_warps=7, spmv_window=2048),
                                  dict(spmv_variant=4), dict(spmv_variant=4, spmv_tile=1024),
                                  dict(spmv_variant=4, spmv_window=1024, spmv_ring_warps=3, spmv_tile=256)])
def test_spmv_window_kernel(gpu, opts):
    """Skewed rows with local columns (AB200_SPMV_MERGE / AUTO): x gathered from a sliding ring
    in shared memory -- spmv_ring2_kernel (cp.async strips, the default) and spmv_ring_kernel
    (register-staged, spmv_variant=4) -- every ring / tile / warp shape.  Rows of <= 16 entries
    stay bit-identical to scipy; longer ones within 1e-13 of the row's absolute sum.  Real and
    complex vectors, complex values, rows longer than several tiles, runs of empty rows."""
    from arnoldi_b200.matrices import powerlaw
    from arnoldi_b200.solver import DeviceSolver
    rng = np.random.default_rng(12)
    mats = {"powerlaw": powerlaw(120000), "local": _local_skewed(60000, rng, [(17, 30000), (50000, 9000)])}
    Ac = _local_skewed(20000, rng)
    Ac = Ac.astype(np.complex128)
    Ac.data = Ac.data + 1j * rng.standard_normal(Ac.nnz)
    mats["complex"] = Ac
    for name, A in mats.items():
        n = A.shape[0]
        lens = np.diff(A.indptr)
        short = lens <= 16
        Aabs = abs(A.copy())
        for algo in ("merge", "auto", "merge64"):
            with DeviceSolver(n, 2) as dev:
                for k, v in opts.items():
                    dev.set_option(k, v)
                if algo == "merge64":       # 64-bit row pointers: the other instantiation
                    dev.set_csr(A.indptr.astype(np.int64), A.indices, A.data, algo="merge")
                else:
                    dev.set_csr(A.indptr, A.indices, A.data, algo=algo)
                for x in (rng.standard_normal(n) + 1j * rng.standard_normal(n),
                          rng.standard_normal(n).astype(np.complex128)):
                    y = dev.spmv(x)
                    ref = A @ x
                    scale = Aabs @ np.abs(x)
                    assert np.all(np.abs(y - ref) <= 1e-13 * scale + 1e-300), (name, algo, opts)
                    np.testing.assert_array_equal(y[short], ref[short], err_msg=f"{name} {algo} {opts}")


@pytest.mark.parametrize("real", [True, False])
def test_fused_sweep_variants_agree_in_expansions(gpu, real):
    """Every schedule of the CGS2 sweep (two-sweep rounds, fused with register loads, cp.async
    staging, mbarrier pipeline) builds the same Arnoldi relation, in float64 and complex128
    storage, on an operator whose DGKS test fires on every step (the fused sweeps' case)."""
    from arnoldi_b200.matrices import lap2d_rect
    from arnoldi_b200.solver import DeviceSolver
    A = lap2d_rect(301, 257)
    n = A.shape[0]
    rng = np.random.default_rng(4)
    v0 = rng.standard_normal(n).astype(np.complex128)
    v0 /= np.linalg.norm(v0)
    m = 34
    out = {}
    for variant in (1, 2, 3, 5):
        for fr in ((0, 1) if variant == 5 else (0,)):
            with DeviceSolver(n, m) as dev:
                if not real:
                    dev.set_option("real_mode", 0)
                dev.set_option("ortho_variant", variant)
                dev.set_option("fused_r", fr)
                dev.set_csr(A.indptr, A.indices, A.data)
                dev.set_columns(0, v0)
                cols, k, brk = dev.expand(0, m, 1e-8)
                assert k == m and not brk
                st = dev.stats()
                assert st["real_storage"] == (1 if real else 0)
                if variant != 1:
                    assert st["ortho_fused_launches"] == m and st["second_rounds"] > m // 2
                H = np.array(cols)
                V = dev.get_columns(0, m + 1)
            out[(variant, fr)] = (H, V)
    H0, V0 = out[(1, 0)]
    assert np.abs(V0.conj().T @ V0 - np.eye(m + 1)).max() < 1e-12
    np.testing.assert_allclose(A @ V0[:, :m], V0 @ H0[:m + 1, :m], rtol=0, atol=1e-12)
    for key, (H, V) in out.items():
        np.testing.assert_allclose(H, H0, rtol=0, atol=2e-12, err_msg=str(key))
        np.testing.assert_allclose(V, V0, rtol=0, atol=2e-11, err_msg=str(key))
