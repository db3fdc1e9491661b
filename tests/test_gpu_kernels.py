"""GPU parity: each n-length kernel, through the C ABI, against the CPU oracle and the
reference goldens.  Bit-exact where the arithmetic order is the reference's (SpMV rows
summed in stored order); otherwise within a few ulp-level tolerances stated in place."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from conftest import csr_from_golden, lap2d

pytestmark = pytest.mark.gpu


def _solver(n, max_dim, A=None, **opts):
    from arnoldi_b200.solver import DeviceSolver
    dev = DeviceSolver(n, max_dim)
    for k, v in opts.items():
        dev.set_option(k, v)
    if A is not None:
        dev.set_csr(A.indptr, A.indices, A.data)
    return dev


def _cvec(rng, n):
    return rng.standard_normal(n) + 1j * rng.standard_normal(n)


# ------------------------------------------------------------------------- SpMV
def _spmv_cases():
    from arnoldi_b200.matrices import lap2d as lap2d_direct, mark
    rng = np.random.default_rng(11)
    cases = {}
    cases["mark50"] = mark(50)
    cases["mark3"] = mark(3)
    cases["lap2d_40"] = lap2d_direct(40)
    cases["lap2d_kron_zeros"] = lap2d(17)           # explicit zeros stored
    A = sp.random(3000, 3000, density=0.004, random_state=rng, format="csr")
    cases["random_f64"] = A
    Ac = A.astype(np.complex128)
    Ac.data = Ac.data + 1j * rng.standard_normal(Ac.nnz)
    cases["random_c128"] = Ac
    # empty rows, including leading / trailing runs, and a 1 x 1 matrix
    B = sp.random(2000, 2000, density=0.002, random_state=rng, format="lil")
    B[:130, :] = 0
    B[700:1400, :] = 0
    B[1990:, :] = 0
    cases["empty_rows"] = B.tocsr()
    cases["all_empty"] = sp.csr_matrix((500, 500), dtype=np.float64)
    cases["one"] = sp.csr_matrix(np.array([[2.5]]))
    # unsorted column order inside rows + duplicates kept (scipy sums them in order)
    rows = rng.integers(0, 400, 3000)
    cols = rng.integers(0, 400, 3000)
    order = np.argsort(rows, kind="stable")
    indptr = np.concatenate(([0], np.cumsum(np.bincount(rows, minlength=400))))
    C = sp.csr_matrix((rng.standard_normal(3000), cols[order].astype(np.int32),
                       indptr.astype(np.int32)), shape=(400, 400))
    cases["unsorted_dups"] = C
    return cases


@pytest.mark.parametrize("name", list(_spmv_cases()))
def test_spmv_bit_exact_vs_scipy(gpu, name):
    """Rows of at most 16 entries (all of mark, the Laplacians, most of the random cases) are
    summed in stored order with separate multiply and add, exactly like scipy's csr_matvec:
    y must match bit for bit.  Longer rows are summed by a warp (fixed butterfly order):
    within 1e-13 of the row's absolute sum."""
    A = _spmv_cases()[name]
    n = A.shape[0]
    rng = np.random.default_rng(5)
    short = np.diff(A.indptr) <= 16
    Aabs = abs(A.copy())      # abs() de-duplicates IN PLACE: keep the operand itself untouched
    with _solver(n, 2, A) as dev:
        for x in (_cvec(rng, n), rng.standard_normal(n).astype(np.complex128)):
            y = dev.spmv(x)
            ref = A @ x
            np.testing.assert_array_equal(y[short], ref[short])
            scale = Aabs @ np.abs(x)
            assert np.all(np.abs(y - ref) <= 1e-13 * scale + 1e-300)
    if name in ("mark50", "lap2d_40", "lap2d_kron_zeros", "empty_rows"):
        assert short.all()


@pytest.mark.parametrize("tile", [512, 1024, 2048])
def test_spmv_skewed_rows_and_int64_indptr(gpu, tile):
    """Power-law row lengths: rows longer than a tile are carried across tile iterations,
    segments longer than 16 entries are reduced by a warp (not stored order => tolerance
    1e-13 relative to the row's absolute sum instead of bit-exact)."""
    rng = np.random.default_rng(3)
    n = 6000
    lens = np.minimum((rng.pareto(1.2, n) * 3).astype(np.int64) + 1, n)
    lens[17] = 5000      # longer than any tile
    lens[18] = 0
    lens[4000] = 2100
    indptr = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    nnz = int(indptr[-1])
    indices = np.concatenate([np.sort(rng.choice(n, l, replace=False)) for l in lens]).astype(np.int32)
    data = rng.uniform(-1, 1, nnz)
    A = sp.csr_matrix((data, indices, indptr), shape=(n, n))
    x = _cvec(rng, n)
    ref = A @ x
    scale = abs(A) @ np.abs(x)
    for ip in (indptr, indptr.astype(np.int32)):
        from arnoldi_b200.solver import DeviceSolver
        with DeviceSolver(n, 2) as dev:
            dev.set_option("spmv_tile", tile)
            dev.set_csr(ip, indices, data)
            y = dev.spmv(x)
        assert np.all(np.abs(y - ref) <= 1e-13 * scale + 1e-300)
        short = lens <= 16
        np.testing.assert_array_equal(y[short], ref[short])


# ------------------------------------------------------------------------- orthogonalisation
def test_ortho_matches_reference_goldens(gpu, golden):
    """dgks_gs / dgks_mgs in isolation on the reference's own outputs (tests/golden/ortho.npz):
    plain, DGKS-second-round and breakdown inputs, c = 1..33."""
    from arnoldi_b200.ortho import dgks_gs, dgks_mgs
    g = golden("ortho")
    for name in g["names"]:
        V = np.asfortranarray(g[f"{name.split('_')[0]}_V"])
        for tag, fn in (("gs", dgks_gs), ("mgs", dgks_mgs)):
            w = g[f"{name}_w"].copy()
            h = np.zeros(V.shape[1], np.complex128)
            beta, brk = fn(w, V, h, 1e-8)
            assert brk == bool(g[f"{name}_{tag}_brk"]), (name, tag)
            wn = np.linalg.norm(g[f"{name}_w"])
            np.testing.assert_allclose(h, g[f"{name}_{tag}_h"], rtol=0, atol=1e-13 * wn)
            np.testing.assert_allclose(w, g[f"{name}_{tag}_w"], rtol=0, atol=2e-13 * wn)
            if not brk:
                np.testing.assert_allclose(beta, g[f"{name}_{tag}_beta"], rtol=1e-10)
            else:
                assert beta < 1e-8


@pytest.mark.parametrize("n,c", [(1, 1), (31, 3), (257, 8), (4099, 17), (100003, 40), (20011, 100),
                                 (5000, 128)])
@pytest.mark.parametrize("kind", ["gs", "gs_fused_regs", "gs_fused_cpasync", "gs_fused_warp",
                                  "gs_fused_pipe", "mgs"])
def test_ortho_vs_oracle_shapes(gpu, n, c, kind):
    """Ragged sizes (n not a multiple of any tile, c = 1..128) against the oracle; every
    CGS2 schedule (two-sweep rounds, fused sweep with register loads / cp.async staging)."""
    import functools
    from arnoldi_b200 import ortho as _o
    dgks_mgs = _o.dgks_mgs
    variant = {"gs": 1, "gs_fused_regs": 2, "gs_fused_cpasync": 3, "gs_fused_warp": 4,
               "gs_fused_pipe": 5}.get(kind, 0)
    dgks_gs = functools.partial(_o.dgks_gs, options={"ortho_variant": variant})
    kind = "mgs" if kind == "mgs" else "gs"
    rng = np.random.default_rng(n * 131 + c)
    c = min(c, n)
    M = rng.standard_normal((n, c)) + 1j * rng.standard_normal((n, c))
    V = np.asfortranarray(np.linalg.qr(M)[0])
    for mix in (1.0, 1e-4):   # 1e-4: w nearly in span(V) -> the DGKS second round fires
        w0 = V @ _cvec(rng, c) + mix * _cvec(rng, n)
        w_ref, h_ref = w0.copy(), np.zeros(c, np.complex128)
        cnt = {}
        fo = oracle.cgs_dgks if kind == "gs" else oracle.mgs_dgks
        beta_ref, brk_ref = fo(w_ref, V, h_ref, 1e-8, counters=cnt)
        w, h = w0.copy(), np.zeros(c, np.complex128)
        beta, brk = (dgks_gs if kind == "gs" else dgks_mgs)(w, V, h, 1e-8)
        wn = np.linalg.norm(w0)
        assert brk == brk_ref
        np.testing.assert_allclose(h, h_ref, rtol=0, atol=1e-13 * wn)
        np.testing.assert_allclose(w, w_ref, rtol=0, atol=4e-13 * wn)
        if n > c:
            np.testing.assert_allclose(beta, beta_ref, rtol=1e-9)
            # the new vector is orthogonal to the basis to working precision
            assert np.abs(V.conj().T @ w).max() <= 1e-13 * wn


def test_ortho_round_counter_and_determinism(gpu):
    from arnoldi_b200 import _lib
    from arnoldi_b200.solver import DeviceSolver
    rng = np.random.default_rng(8)
    n, c = 50000, 24
    V = np.asfortranarray(np.linalg.qr(rng.standard_normal((n, c)) + 0j)[0])
    w_far = _cvec(rng, n)
    w_near = V @ _cvec(rng, c) + 1e-5 * _cvec(rng, n)
    with DeviceSolver(n, c) as dev:
        dev.set_columns(0, V)
        outs = []
        for w0, rounds in ((w_far, 1), (w_near, 2), (w_near, 2)):
            dev.reset_stats()
            w, h = w0.copy(), np.zeros(c, np.complex128)
            dev.ortho(c, w, h, 1e-8, np.sqrt(0.5), _lib.ORTHO_CGS2)
            st = dev.stats()
            assert st["ortho_rounds"] == rounds and st["second_rounds"] == rounds - 1
            outs.append((w, h))
        # same input twice -> identical bits (fixed-order reductions, no float atomics)
        np.testing.assert_array_equal(outs[1][0], outs[2][0])
        np.testing.assert_array_equal(outs[1][1], outs[2][1])


# ------------------------------------------------------------------------- restart
def test_restart_matches_reference_golden(gpu, golden):
    from arnoldi_b200.solver import DeviceSolver
    g = golden("restart")
    V, Q, m, p = g["V"], g["Q"], int(g["m"]), int(g["p"])
    with DeviceSolver(V.shape[0], m) as dev:
        dev.set_columns(0, V)
        dev.restart(Q, m, p)
        out = dev.get_columns(0, m + 1)
    np.testing.assert_allclose(out[:, : p + 1], g["Vout"][:, : p + 1], rtol=0, atol=1e-14)
    # columns beyond p are left as they were (the reference does not touch them either)
    np.testing.assert_array_equal(out[:, p + 1:], V[:, p + 1:])


@pytest.mark.parametrize("n,m,p", [(1, 2, 1), (33, 5, 3), (1000, 20, 10), (70001, 40, 15),
                                   (9001, 60, 25), (3000, 100, 85), (2000, 128, 127)])
def test_restart_vs_oracle_shapes(gpu, n, m, p):
    from arnoldi_b200.solver import DeviceSolver
    rng = np.random.default_rng(n + m)
    V = np.asfortranarray(rng.standard_normal((n, m + 1)) + 1j * rng.standard_normal((n, m + 1)))
    Q = np.linalg.qr(rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m)))[0]
    ref = V.copy(order="F")
    oracle.restart_update(ref, Q, m, p)
    with DeviceSolver(n, m) as dev:
        dev.set_columns(0, V)
        dev.restart(Q, m, p)
        out = dev.get_columns(0, m + 1)
    scale = np.abs(V[:, :m]) @ np.abs(Q[:, :p])
    assert np.all(np.abs(out[:, :p] - ref[:, :p]) <= 4e-16 * m * scale + 1e-300)
    np.testing.assert_array_equal(out[:, p], ref[:, p])
    np.testing.assert_array_equal(out[:, p + 1:], ref[:, p + 1:])


# ------------------------------------------------------------------------- Arnoldi expansion
def _check_invariants(A, V, H, k):
    """tests/test_decomposition.py:36-68 of the reference."""
    Vm, Hm = V[:, :k], H[:k, :k]
    np.testing.assert_allclose(V[:, : k + 1].conj().T @ V[:, : k + 1], np.eye(k + 1), atol=1e-12)
    lhs = A @ Vm
    rhs = Vm @ Hm + np.outer(V[:, k], H[k, :k])
    np.testing.assert_allclose(lhs, rhs, atol=1e-12 * max(1.0, np.abs(lhs).max()))
    np.testing.assert_allclose(Vm.conj().T @ (A @ Vm), Hm, rtol=1e-4, atol=1e-8)


@pytest.mark.parametrize("order", ["F", "C"])
@pytest.mark.parametrize("ortho", ["cgs2", "mgs"])
def test_arnoldi_decomposition_matches_reference(gpu, golden, order, ortho):
    """The reference's own expansion outputs (tests/golden/arnoldi.npz), with the caller's
    V in either memory order (the reference's tests pass C-order)."""
    from arnoldi_b200.decomposition import arnoldi_decomposition
    g = golden("arnoldi")
    tol = 1e-12 if ortho == "cgs2" else 1e-10   # MGS differs from the CGS2 golden by rounding
    # complex sparse operator, full expansion
    A = csr_from_golden(g, "cplx")
    n, m = A.shape[0], g["cplx_H"].shape[1]
    V = np.zeros((n, m + 1), np.complex128, order=order)
    H = np.zeros((m + 1, m), np.complex128)
    V[:, 0] = g["cplx_v0"]
    Va, Ha, k = arnoldi_decomposition(A, V, H, 1e-8, ortho=ortho)
    assert k == int(g["cplx_niter"]) and Va.shape == (n, k + 1) and Ha.shape == (k + 1, k)
    np.testing.assert_allclose(V, g["cplx_V"], rtol=0, atol=tol)
    np.testing.assert_allclose(H, g["cplx_H"], rtol=0, atol=tol * 10)
    _check_invariants(A, V, H, k)
    # real operator, max_dim < m: trailing columns stay zero
    A = csr_from_golden(golden("matrices"), "mark10")
    n, m, md = A.shape[0], g["mark10_H"].shape[1], int(g["mark10_maxdim"])
    V = np.zeros((n, m + 1), np.complex128, order=order)
    H = np.zeros((m + 1, m), np.complex128)
    V[:, 0] = g["mark10_v0"]
    Va, Ha, k = arnoldi_decomposition(A, V, H, 1e-8, max_dim=md, ortho=ortho)
    assert k == md and Va.shape == (n, md + 1) and Ha.shape == (md + 1, md)
    np.testing.assert_allclose(V, g["mark10_V"], rtol=0, atol=tol)
    np.testing.assert_allclose(H, g["mark10_H"], rtol=0, atol=tol * 10)
    assert not V[:, md + 1:].any() and not H[:, md:].any()


def test_arnoldi_breakdown_matches_reference(gpu, golden):
    """v0 an eigenvector of a dense diagonal A: the first step breaks down, n_iter == 1,
    H[1, 0] is not written (tests/test_decomposition.py:115-139)."""
    from arnoldi_b200.decomposition import arnoldi_decomposition
    g = golden("arnoldi")
    Ad = g["brk_A"]
    n, m = Ad.shape[0], g["brk_H"].shape[1]
    V = np.zeros((n, m + 1), np.complex128, order="F")
    H = np.zeros((m + 1, m), np.complex128)
    V[:, 0] = g["brk_v0"]
    Va, Ha, k = arnoldi_decomposition(Ad, V, H, 1e-8)
    assert k == int(g["brk_niter"]) == 1
    assert Va.shape == (n, 2) and Ha.shape == (2, 1)
    np.testing.assert_allclose(H, g["brk_H"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(V, g["brk_V"], rtol=0, atol=1e-14)
    assert H[1, 0] == 0


def test_arnoldi_restartable_and_argument_checks(gpu):
    """start_dim > 0 continues an existing relation; shape errors are AssertionErrors."""
    from arnoldi_b200.decomposition import arnoldi_decomposition
    from arnoldi_b200.matrices import mark
    A = mark(12)
    n, m = A.shape[0], 14
    rng = np.random.default_rng(2)
    v0 = _cvec(rng, n)
    v0 /= np.linalg.norm(v0)
    V1 = np.zeros((n, m + 1), np.complex128, order="F"); H1 = np.zeros((m + 1, m), np.complex128)
    V1[:, 0] = v0
    arnoldi_decomposition(A, V1, H1, 1e-10)
    V2 = np.zeros((n, m + 1), np.complex128, order="F"); H2 = np.zeros((m + 1, m), np.complex128)
    V2[:, 0] = v0
    arnoldi_decomposition(A, V2, H2, 1e-10, max_dim=6)
    arnoldi_decomposition(A, V2, H2, 1e-10, start_dim=6)
    np.testing.assert_allclose(V2, V1, rtol=0, atol=1e-13)
    np.testing.assert_allclose(H2, H1, rtol=0, atol=1e-13)
    _check_invariants(A, V1, H1, m)
    with pytest.raises(AssertionError):
        arnoldi_decomposition(A, V1[:-1], H1, 1e-10)
    with pytest.raises(AssertionError):
        arnoldi_decomposition(A, V1, H1[:, :-1], 1e-10)
    with pytest.raises(AssertionError):
        arnoldi_decomposition(A, V1, H1, 1e-10, max_dim=m + 1)
