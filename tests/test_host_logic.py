"""CPU-side checks: the C ABI library loads and exports what the header declares,
the host control logic matches the reference goldens, the fixtures are bit-exact."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT, csr_from_golden, lap2d


def test_library_exports_every_declared_symbol():
    from arnoldi_b200 import _lib
    header = open(os.path.join(ROOT, "include", "arnoldi_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(ab200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no prototypes found in the header"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    lib = _lib.load()
    assert lib.ab200_abi_version() == _lib.ABI_VERSION


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product path must fail loudly, not fall back."""
    from arnoldi_b200 import _lib, partial_schur
    from arnoldi_b200.matrices import mark
    lib = _lib.load()
    if lib.ab200_device_count() >= 1:
        pytest.skip("a GPU is visible here")
    with pytest.raises(_lib.DeviceError):
        partial_schur(mark(10), 3, max_dim=5)


def test_mark_is_bit_identical_to_reference(golden):
    from arnoldi_b200.matrices import mark
    g = golden("matrices")
    for m in (2, 3, 10, 17, 20, 50):
        A = mark(m)
        assert A.shape == tuple(g[f"mark{m}_shape"])
        np.testing.assert_array_equal(A.indptr, g[f"mark{m}_indptr"])
        np.testing.assert_array_equal(A.indices, g[f"mark{m}_indices"])
        np.testing.assert_array_equal(A.data, g[f"mark{m}_data"])  # bit-exact
    # tests/test_matrices.py:8-19 of the reference
    np.testing.assert_array_equal(mark(2).toarray(), [[0, 1, 1], [0.5, 0, 0], [0.5, 0, 0]])
    assert mark(50).shape == (1275, 1275) and mark(50).nnz == 4900


def test_laplace_fixtures(golden):
    from arnoldi_b200.matrices import laplace, laplace_eigen, lap2d as lap2d_direct
    g = golden("matrices")
    np.testing.assert_array_equal(laplace(5).toarray(), csr_from_golden(g, "laplace5").toarray())
    np.testing.assert_array_equal(laplace_eigen(5), g["laplace_eigen5"])
    for N in (2, 3, 8, 33):
        A, B = lap2d_direct(N), lap2d(N)
        B.eliminate_zeros()
        B.sort_indices()
        np.testing.assert_array_equal(A.indptr, B.indptr)
        np.testing.assert_array_equal(A.indices, B.indices)
        np.testing.assert_array_equal(A.data, B.data)
    A = lap2d_direct(64)
    assert A.nnz == 5 * 64 * 64 - 4 * 64 and A.indices.dtype == np.int32


def test_ordered_schur_matches_reference(golden):
    from scipy.linalg import schur
    from arnoldi_b200.utils import arg_largest_real, ordered_schur
    g = golden("restart")
    T1, Q1 = schur(g["Hm"], output="complex")
    T2, Q2 = ordered_schur(T1, output="complex", sort_function=arg_largest_real)
    np.testing.assert_allclose(T2, g["T2"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(Q2, g["Q2"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(Q1 @ Q2, g["Q"], rtol=1e-12, atol=1e-13)
    with pytest.raises(ValueError):
        ordered_schur(T1, output="real")
    # complex64 path (tests/test_utils.py:23-35 of the reference)
    rng = np.random.default_rng(0)
    a = (rng.standard_normal((6, 6)) + 1j * rng.standard_normal((6, 6))).astype(np.complex64)
    T, Z = ordered_schur(a, output="complex")
    d = np.abs(np.diag(T))
    assert np.all(d[:-1] >= d[1:] - 1e-6)
    np.testing.assert_allclose(Z @ T @ Z.conj().T, a, atol=1e-4)


def test_operator_conversion():
    from arnoldi_b200.operator import as_csr
    A = sp.random(20, 20, density=0.2, random_state=1, format="csr")
    ip, ix, d, shape = as_csr(A)
    assert ip is A.indptr and ix is A.indices and d is A.data and shape == (20, 20)
    ip, ix, d, _ = as_csr(A.tocoo())
    np.testing.assert_array_equal(d, A.tocoo().tocsr().data)
    D = A.toarray()
    ip, ix, d, _ = as_csr(D)
    np.testing.assert_array_equal(sp.csr_matrix((d, ix, ip), shape=(20, 20)).toarray(), D)
    ip, ix, d, _ = as_csr(A.astype(np.float32))
    assert d.dtype == np.float64
    # wrappers that expose the wrapped matrix as `.A` are unwrapped (scripts/utils.py:55-68 of
    # the reference wraps A in MatvecCounter(LinearOperator) before calling partial_schur)
    from scipy.sparse.linalg import LinearOperator, aslinearoperator
    from arnoldi_b200.operator import credit_matvecs, unwrap
    ip, ix, d, shape = as_csr(aslinearoperator(A))
    assert ip is A.indptr and d is A.data and shape == (20, 20)

    class MatvecCounter(LinearOperator):          # the reference's class, with a working base init
        def __init__(self, A):
            super().__init__(np.dtype(A.dtype), A.shape)
            self.A = A
            self.matvecs = 0

        def _matvec(self, x):
            self.matvecs += 1
            return self.A @ x

    class BareCounter:                            # scripts/utils.py:55-60 verbatim shape: no base init
        def __init__(self, A):
            self.A, self.shape, self.dtype, self.matvecs = A, A.shape, np.dtype(A.dtype), 0

    for W in (MatvecCounter, BareCounter):
        outer = W(W(A))
        M, wrappers = unwrap(outer)
        assert M is A and len(wrappers) == 2
        credit_matvecs(wrappers, 37)
        assert outer.matvecs == 37 and outer.A.matvecs == 37
        assert as_csr(outer)[2] is A.data
    # an opaque callable has no entries to upload: loud TypeError, never a host fallback
    opaque = LinearOperator((20, 20), matvec=lambda x: A @ x, dtype=np.float64)
    with pytest.raises(TypeError, match="LinearOperator"):
        as_csr(opaque)


def test_history_and_argument_checks():
    from arnoldi_b200 import History, partial_schur
    from arnoldi_b200.matrices import mark
    h = History.from_k(4)
    assert h.k == 4 and h.matvecs.dtype == np.int32 and h.total_matvecs == 0
    A = mark(10)
    # krylov_schur.py:24,27,36: argument errors are AssertionErrors, raised before any device work
    with pytest.raises(AssertionError):
        partial_schur(A, 3, max_restarts=0)
    with pytest.raises(AssertionError):
        partial_schur(A, 3, max_dim=3)
    with pytest.raises(AssertionError):
        partial_schur(A[:, :10], 3)
    with pytest.raises(AssertionError):
        partial_schur(A, 3, ortho="householder")


def test_powerlaw_generator_is_shardable():
    """Config-4 operator: any block of rows regenerates bit-identically, whatever the chunking."""
    from arnoldi_b200.matrices import POWERLAW_TOP, powerlaw, powerlaw_rows
    n = 30000
    A = powerlaw(n)
    lens = np.diff(A.indptr)
    assert lens.min() >= 4 and lens.max() <= 2048 and 8 < lens.mean() < 25
    assert lens.max() > 20 * np.median(lens)            # skewed rows
    for r0, r1, chunk in ((0, n, 1 << 20), (123, 20011, 4096), (29990, n, 7)):
        b = powerlaw_rows(n, r0, r1, chunk=chunk)
        S = A[r0:r1]
        np.testing.assert_array_equal(b.indptr, S.indptr)
        np.testing.assert_array_equal(b.indices, S.indices)
        np.testing.assert_array_equal(b.data, S.data)
    # nonsymmetric, diagonal first in every row, separated leading diagonal entries
    assert (A != A.T).nnz > 0
    np.testing.assert_array_equal(A.indices[A.indptr[:-1]], np.arange(n))
    d = np.sort(A.diagonal())[::-1]
    np.testing.assert_allclose(d[:POWERLAW_TOP], 3.0 + 0.25 * np.arange(POWERLAW_TOP)[::-1])
    assert d[POWERLAW_TOP] <= 2.0


def test_ordered_schur_triangular_shortcut_is_exact():
    """ordered_schur skips LAPACK's zgees when its input is already upper triangular (what
    partial_schur passes).  That must be bit-identical to calling zgees: zgees maps a triangular
    matrix to itself with Z = I."""
    from scipy.linalg import schur
    from scipy.linalg.lapack import ztrexc
    from arnoldi_b200.utils import arg_largest_magnitude, arg_largest_real, ordered_schur
    rng = np.random.default_rng(12)
    for trial in range(40):
        m = int(rng.integers(2, 64))
        if trial % 2:
            H = np.triu(rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m)), -1)
        else:
            H = np.triu(rng.standard_normal((m, m)), -1).astype(np.complex128)
        T1, Q1 = schur(H, output="complex")
        T0, Z0 = schur(T1, output="complex")
        np.testing.assert_array_equal(T0, T1)
        np.testing.assert_array_equal(Z0, np.eye(m))
        for sort in (arg_largest_real, arg_largest_magnitude):
            # the reference's sequence, spelled out with the zgees call it makes
            T, Z = T0.copy(), Z0.copy()
            slots = list(range(m))
            for dest, original in enumerate(sort(np.diag(T0))):
                here = slots.index(original)
                if here != dest:
                    T, Z, _ = ztrexc(T, Z, here + 1, dest + 1)
                    slots.insert(dest, slots.pop(here))
            T2, Z2 = ordered_schur(T1, output="complex", sort_function=sort)
            np.testing.assert_array_equal(T2, T)
            np.testing.assert_array_equal(Z2, Z)


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver times beside the GPU arm) must run
    without a GPU and print exactly one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, OMP_NUM_THREADS="1")      # what torchrun exports to its children
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "20", "--warmup", "5", "--grid", "192", "--no-converged"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["unit"] == "matvec/s" and rec["value"] > 0
    assert rec["higher_is_better"] is True and rec["vs_baseline"] is None
    ref_installed = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "arnoldi", "__init__.py"))
    assert rec["cpu_baseline"]["kind"] == ("reference" if ref_installed else "port")
    # the launcher's OMP_NUM_THREADS=1 must not shrink the baseline to one core
    assert rec["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    # what the driver checks: steps x ms_per_step is the region that was really timed
    assert rec["steps"] == 20 and rec["warmup"] == 5
    assert abs(rec["steps"] * rec["ms_per_step"] * 1e-3 - rec["timed_region_s"]) < 1e-9
    assert rec["timed_cycles"] == 2 and rec["timed_matvecs"] == 50
    assert rec["e2e"]["h2d_bytes_per_step"] == 0 and rec["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in rec["config"]


def test_reference_import_names_resolve_to_the_device_path():
    """User code written against the reference (README.md:20-31) imports `arnoldi`; with this
    repo's package directory on sys.path those names are the B200 implementations."""
    import arnoldi
    import arnoldi.krylov_schur
    import arnoldi_b200
    from arnoldi.decomposition import arnoldi_decomposition
    from arnoldi.explicit_restarts import History
    from arnoldi.matrices import mark
    from arnoldi.ortho import dgks_gs, dgks_mgs
    from arnoldi.utils import arg_largest_real, ordered_schur
    assert arnoldi.partial_schur is arnoldi_b200.partial_schur
    assert arnoldi.krylov_schur.partial_schur is arnoldi_b200.partial_schur
    assert History is arnoldi_b200.History and mark(3).shape == (6, 6)
    assert all(callable(f) for f in (arnoldi_decomposition, dgks_gs, dgks_mgs, arg_largest_real,
                                     ordered_schur))
    assert "arnoldi-py_b200" in arnoldi.__file__


def test_fast_real_rotate_paths():
    """fast_real_schur's host rotate: a real symmetric H_m (what a symmetric operator projects
    to, up to rounding) takes dsyevd + a column permutation, a real nonsymmetric one dgees + 2 x 2
    rotations, anything else zgees.  Every path returns an ordered Schur form of H: unitary Q,
    upper-triangular T with the diagonal in the order sort_function asks for, H Q = Q T; the
    eigenvalues agree with the reference's zgees path to rounding."""
    from arnoldi_b200.rotate import rotate
    from arnoldi_b200.utils import arg_largest_magnitude, arg_largest_real
    rng = np.random.default_rng(21)
    for trial in range(30):
        m = int(rng.integers(3, 64))
        p = int(rng.integers(1, m - 1))
        d, e, s = rng.standard_normal(m), rng.standard_normal(m - 1), 1e-3 * rng.standard_normal(p)
        H = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)      # Krylov-Schur shape: diagonal block,
        H[:p, :p] = np.diag(d[:p])                            # spike row / column, tridiagonal rest
        H[p, :p] = s
        H[:p, p] = s
        kind = trial % 3
        if kind == 0:      # symmetric up to rounding noise
            H = H + 1e-15 * np.triu(rng.standard_normal((m, m)), 1)
        elif kind == 1:    # real, clearly nonsymmetric
            H = H + np.triu(rng.standard_normal((m, m)), 1)
        else:              # complex
            H = H + 1j * np.triu(rng.standard_normal((m, m)), 0)
        H = H.astype(np.complex128)
        for sort in (arg_largest_real, arg_largest_magnitude):
            T0, Q0 = rotate(H, sort, fast_real=False)
            T, Q = rotate(H, sort, fast_real=True)
            scale = np.abs(H).max()
            assert np.abs(Q.conj().T @ Q - np.eye(m)).max() < 1e-13
            assert np.abs(np.tril(T, -1)).max() == 0.0
            assert np.abs(H @ Q - Q @ T).max() < 1e-12 * scale * m
            np.testing.assert_array_equal(np.diag(T)[sort(np.diag(T))], np.diag(T))
            if kind == 0:
                assert np.abs(np.triu(T, 1)).max() == 0.0 and not np.any(Q.imag)      # dsyevd path
                np.testing.assert_allclose(np.diag(T), np.diag(T0), rtol=0, atol=1e-12 * scale)
            if kind == 2:
                np.testing.assert_array_equal(T, T0)                                   # zgees both ways
