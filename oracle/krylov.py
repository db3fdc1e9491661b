"""NumPy/SciPy restatement of the reference's Krylov-Schur hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference (cournape/arnoldi-py) has no native code: its n-length
arithmetic is scipy's ``csr_matvec`` and OpenBLAS ``zgemv``/``dznrm2``/``zgemm``
and its m-by-m arithmetic is LAPACK ``zgees``/``ztrexc``.  This module states the
same algorithm, one function per n-length operation, calling the same library
routines in the same order so that (a) on one machine it reproduces the
reference bit for bit (pinned by ``tests/golden``), and (b) timing it is timing
the reference's CPU path.  Citations are ``file:line`` under
``/root/reference/src/arnoldi``.

Functions (n-length ops first, control logic after):

* ``spmv``            -- ``decomposition.py:57-58``
* ``cgs_dgks``        -- ``ortho.py:56-107``
* ``mgs_dgks``        -- ``ortho.py:9-53``
* ``restart_update``  -- ``krylov_schur.py:78,81``
* ``arnoldi_expand``  -- ``decomposition.py:13-68``
* ``sorted_schur``    -- ``utils.py:32-67``
* ``partial_schur``   -- ``krylov_schur.py:10-114``
* ``explicit_restarts_with_deflation`` -- ``explicit_restarts.py:63-168``
"""

from __future__ import annotations

import dataclasses

import numpy as np
from scipy.linalg import get_blas_funcs, schur
from scipy.linalg.lapack import ztrexc

_znrm2, _zgemv = get_blas_funcs(("nrm2", "gemv"), dtype=np.complex128)

SQRT_HALF = np.sqrt(0.5)


# --------------------------------------------------------------------------
# n-length operations
# --------------------------------------------------------------------------
def spmv(A, x, out):
    """out[:] = A @ x   (``decomposition.py:57-58``).

    ``A`` is whatever the caller passed (scipy CSR, ndarray, LinearOperator);
    for CSR this lands in ``scipy.sparse._sparsetools.csr_matvec`` which sums
    each row sequentially in stored order.
    """
    out[:] = A @ x


def cgs_dgks(w, V, h, tol=1e-8, eta=SQRT_HALF, counters=None):
    """Classical Gram-Schmidt with one DGKS-triggered repeat (``ortho.py:56-107``).

    Round: c = V^H w (zgemv, conjugate-transpose), w -= V c (zgemv), norm.
    The round is repeated once, accumulating into ``h``, when the norm dropped
    by more than the factor ``eta`` (strict ``<``, ``ortho.py:101``).
    Breakdown is the absolute test ``beta < tol`` (``ortho.py:107``).
    ``counters`` (optional dict) gets ``rounds`` incremented: test-side
    bookkeeping, not part of the reference.
    """
    ncols = V.shape[1]
    norm_in = _znrm2(w)  # ortho.py:92

    c = _zgemv(1.0, V, w, trans=2)  # ortho.py:94
    h[: ncols + 1] = c  # ortho.py:95 (slice clipped to len(h) == ncols)
    w -= _zgemv(1.0, V, c)  # ortho.py:96
    beta = _znrm2(w)  # ortho.py:98
    rounds = 1

    if beta < norm_in * eta:  # ortho.py:101
        c = _zgemv(1.0, V, w, trans=2)
        h[: ncols + 1] += c
        w -= _zgemv(1.0, V, c)
        beta = _znrm2(w)
        rounds = 2

    if counters is not None:
        counters["rounds"] = counters.get("rounds", 0) + rounds
        counters["calls"] = counters.get("calls", 0) + 1
    return beta, beta < tol


def mgs_dgks(w, V, h, tol=1e-8, eta=SQRT_HALF, counters=None):
    """Modified Gram-Schmidt with one DGKS-triggered repeat (``ortho.py:9-53``).

    Column by column: h_i = <V_i, w> (vdot conjugates its first argument),
    then w -= h_i V_i.  A second sweep accumulates into ``h`` under the same
    criterion as ``cgs_dgks``.
    """
    ncols = V.shape[1]
    norm_in = np.linalg.norm(w)  # ortho.py:36

    for i in range(ncols):  # ortho.py:39-41
        h[i] = np.vdot(V[:, i], w)
        w -= h[i] * V[:, i]

    norm_mid = np.linalg.norm(w)  # ortho.py:43
    rounds = 1

    if norm_mid < eta * norm_in:  # ortho.py:46
        for i in range(ncols):
            c = np.vdot(V[:, i], w)
            h[i] += c
            w -= c * V[:, i]
        rounds = 2

    beta = np.linalg.norm(w)  # ortho.py:52
    if counters is not None:
        counters["rounds"] = counters.get("rounds", 0) + rounds
        counters["calls"] = counters.get("calls", 0) + 1
    return beta, beta < tol


def restart_update(V, Q, m, p):
    """Krylov-Schur truncation of the basis (``krylov_schur.py:78,81``).

    V[:, :p] <- V[:, :m] Q[:, :p]  (zgemm into a temporary, then copied), and
    the residual direction V[:, m] moves to column p.
    """
    V[:, :p] = V[:, :m] @ Q[:, :p]
    V[:, p] = V[:, m]


# --------------------------------------------------------------------------
# Arnoldi expansion
# --------------------------------------------------------------------------
def arnoldi_expand(A, V, H, tol=None, *, start_dim=0, max_dim=None,
                   ortho=cgs_dgks, counters=None):
    """Grow the Arnoldi relation in place from ``start_dim`` to ``max_dim``.

    Follows ``decomposition.py:13-68``.  Returns ``(V_view, H_view, n_iter)``;
    on breakdown at step j the new vector is left un-normalised, H[j+1, j] is
    not written and ``n_iter == j + 1`` (``decomposition.py:61-63``).
    """
    if tol is None:
        tol = np.sqrt(np.finfo(A.dtype).eps)  # decomposition.py:41-42

    n = A.shape[0]
    m = V.shape[1] - 1
    assert A.shape[1] == n, "A is expected to be square matrix"
    assert V.shape == (n, m + 1), "V must have the same number of rows as A"
    assert H.shape == (m + 1, m), f"H must be {m + 1, m}, is {H.shape}"
    if max_dim is None:
        max_dim = m
    assert max_dim <= m, "max_dim > m violated"

    for j in range(start_dim, max_dim):
        w = V[:, j + 1]
        spmv(A, V[:, j], w)
        if counters is not None:
            counters["matvecs"] = counters.get("matvecs", 0) + 1
        if counters is not None:
            beta, broke = ortho(w, V[:, : j + 1], H[: j + 1, j], tol, counters=counters)
        else:
            beta, broke = ortho(w, V[:, : j + 1], H[: j + 1, j], tol)
        if broke:
            k = j + 1
            return V[:, : k + 1], H[: k + 1, :k], k
        H[j + 1, j] = beta
        w /= beta  # decomposition.py:66
    return V[:, : max_dim + 1], H[: max_dim + 1, :max_dim], max_dim


# --------------------------------------------------------------------------
# m-by-m control logic
# --------------------------------------------------------------------------
def rand_unit_vector(n, dtype=np.float64):
    """Start vector from the legacy global NumPy RNG (``utils.py:7-13``)."""
    v = np.random.randn(n).astype(dtype)
    v /= np.linalg.norm(v)
    return v


def arg_largest_magnitude(x):
    """``utils.py:16-17``"""
    return np.argsort(-np.abs(x))


def arg_largest_real(x):
    """``utils.py:20-21``"""
    return np.argsort(-np.real(x))


def sorted_schur(a, sort_function=None):
    """Complex Schur form with the diagonal in ``sort_function`` order.

    ``utils.py:32-67`` (complex output only; the reference raises for real
    output).  zgees, then one ztrexc move per target slot while a Python list
    tracks where each original diagonal entry currently sits.
    """
    if sort_function is None:
        sort_function = arg_largest_magnitude
    T, Z = schur(a, output="complex")
    order = sort_function(np.diag(T))
    where = list(range(T.shape[0]))
    for target, wanted in enumerate(order):
        source = where.index(wanted)
        if source != target:
            T, Z, _info = ztrexc(T, Z, source + 1, target + 1)  # 1-based
            where.insert(target, where.pop(source))
    return T, Z


@dataclasses.dataclass
class History:
    """``explicit_restarts.py:13-28``"""

    matvecs: np.ndarray
    restarts: np.ndarray

    @classmethod
    def from_k(cls, k):
        return cls(np.zeros(k, np.int32), np.zeros(k, np.int32))

    @property
    def k(self):
        return self.matvecs.shape[0]

    @property
    def total_matvecs(self):
        return self.matvecs.sum()


# --------------------------------------------------------------------------
# Krylov-Schur driver
# --------------------------------------------------------------------------
def partial_schur(A, nev, *, max_dim=None, stopping_criterion=None,
                  max_restarts=100, sort_function=None, p=None,
                  ortho=cgs_dgks, counters=None):
    """Partial Schur decomposition by Krylov-Schur (``krylov_schur.py:10-114``).

    Returns ``(Q, T, history)``.  ``ortho`` and ``counters`` are oracle-side
    extras (the reference hard-wires ``dgks_gs``, ``decomposition.py:60``).
    """
    tol = np.sqrt(np.finfo(A.dtype).eps) if stopping_criterion is None \
        else stopping_criterion
    if sort_function is None:
        sort_function = arg_largest_magnitude
    assert max_restarts > 0
    n = A.shape[0]
    assert A.shape[1] == n
    if max_dim is None:
        max_dim = min(max(2 * nev + 1, 20), n)
    if p is None:
        p = min(nev + 5, max_dim - 1)
    assert nev <= p < max_dim

    V = np.zeros((n, max_dim + 1), dtype=np.complex128, order="F")
    H = np.zeros((max_dim + 1, max_dim), dtype=np.complex128)
    V[:, 0] = rand_unit_vector(n, np.complex128)

    history = History.from_k(nev)
    done = False
    _, _, m = arnoldi_expand(A, V, H, tol, start_dim=0, max_dim=max_dim,
                             ortho=ortho, counters=counters)

    for restart in range(max_restarts):
        if m != max_dim:
            raise ValueError("Happy breakdown not supported yet")
        reported_matvecs = restart * (max_dim - nev) + (m - nev)  # :63

        # rotate: Schur form of H_m, reordered (zgees runs twice, as in the
        # reference: once here and once inside sorted_schur on T1)
        T1, Q1 = schur(H[:m, :m], output="complex")
        T2, Q2 = sorted_schur(T1, sort_function)
        Q = Q1 @ Q2
        Qp = Q[:, :p]

        coupling = H[m, :m].copy()
        beta_last = H[m, m - 1]

        # truncate
        restart_update(V, Q, m, p)
        H[:p, :p] = T2[:p, :p]
        H[p, :p] = coupling @ Qp
        H[p, p:] = 0

        # convergence (:91-101)
        est = np.abs(beta_last * Q[m - 1, :]) / np.abs(np.diag(T2))
        for k in range(nev):
            if est[k] <= tol:
                history.matvecs[k] = reported_matvecs
                history.restarts[k] = restart + 1
        if np.all(est[:nev] < tol):
            done = True
            break

        _, _, m = arnoldi_expand(A, V, H, tol, start_dim=p, max_dim=max_dim,
                                 ortho=ortho, counters=counters)

    if not done:
        raise ValueError("Has not converged !")
    return V[:, :nev], H[:nev, :nev], history


# --------------------------------------------------------------------------
# Explicit restarts with deflation
# --------------------------------------------------------------------------
def _mgs_sweep(basis, w, tol):
    """``explicit_restarts.py:63-77``: one modified Gram-Schmidt sweep, then normalise."""
    for j in range(basis.shape[1]):
        w -= np.vdot(basis[:, j], w) * basis[:, j]
    beta = np.linalg.norm(w)
    assert beta > tol, "MGS: Too small norm when orthornormalizing"
    w /= beta
    return w


def explicit_restarts_with_deflation(A, nev, *, max_dim=None, stopping_criterion=None,
                                     max_restarts=100, sort_function=None):
    """``explicit_restarts.py:80-168``: one eigenpair after the other; Arnoldi from column k,
    leading Ritz vector of the active block back into column k, MGS against the locked
    columns, until its residual estimate passes; H column of the locked vector by explicit
    projection; eigenpairs of the final nev x nev block."""
    tol = np.sqrt(np.finfo(A.dtype).eps) if stopping_criterion is None else stopping_criterion
    if sort_function is None:
        sort_function = arg_largest_magnitude
    assert max_restarts > 0
    n = A.shape[0]
    assert A.shape[1] == n
    if max_dim is None:
        max_dim = min(max(2 * nev + 1, 20), n)
    V = np.zeros((n, max_dim + 1), dtype=np.complex128)
    H = np.zeros((max_dim + 1, max_dim), dtype=np.complex128)
    history = History.from_k(nev)
    for k in range(nev):
        v0 = rand_unit_vector(n, np.complex128)
        _mgs_sweep(V[:, :k], v0, tol)
        V[:, k] = v0
        for restart in range(max_restarts):
            _, _, m = arnoldi_expand(A, V, H, tol, start_dim=k)
            assert m > k
            lucky = m != max_dim
            matvecs = restart * (max_dim - k) + (m - k)
            Hk = H[k:, k:]
            w, S = np.linalg.eig(Hk[: m - k, : m - k])         # decomposition.py:123-126
            ind = sort_function(w)[: m - k]
            S = S[:, ind]
            values = w[ind]
            resid = np.abs(Hk[m - k, m - k - 1] * S[-1])       # decomposition.py:129
            V[:, k] = V[:, k:m] @ S[:, 0]
            _mgs_sweep(V[:, :k], V[:, k], tol)
            if lucky or (resid / np.abs(values))[0] < tol:
                for i in range(k + 1):
                    H[i, k] = np.vdot(V[:, i], A @ V[:, k])
                H[k + 1:-1, k] = 0
                history.matvecs[k] = matvecs
                history.restarts[k] = restart + 1
                break
        else:
            raise ValueError(f"Could not converge for value {k}")
    vals, Y = np.linalg.eig(H[:nev, :nev])
    return vals, V[:, :nev] @ Y, history
