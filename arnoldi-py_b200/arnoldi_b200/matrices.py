"""Test / benchmark operators (reference: src/arnoldi/matrices.py), built vectorised.

``mark(m)`` in the reference is a pure-Python double loop (matrices.py:30-71), which at
the benchmark size m = 4000 means 8 M iterations; here the same CSR matrix is assembled
with array arithmetic.  The result is bit-identical to the reference's
``coo_matrix(...).tocsr()`` (sorted columns, duplicate entries summed), which
tests/test_matrices.py checks against the committed goldens.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def _index_dtype(nnz, n):
    return np.int32 if max(nnz, n) < 2**31 - 1 else np.int64


def mark(m):
    """Markov chain of a random walk on an m-row triangular grid (Saad, sec. 2.5.1).

    n = m (m + 1) / 2 states, at most 4 transitions per state (west, south, north, east
    in column order), nonsymmetric.  matrices.py:5-73.
    """
    n = m * (m + 1) // 2
    cst = 0.5 / (m - 1)
    i = np.repeat(np.arange(m), np.arange(m, 0, -1))          # grid row of each state
    start = np.concatenate(([0], np.cumsum(np.arange(m, 0, -1))[:-1]))
    j = np.arange(n) - start[i]                                # position inside the row
    width = m - i                                              # states in this grid row
    idx = np.arange(n)

    up = cst * (i + j + 1)                                     # probability of north / east
    down = 0.5 - cst * (i + j - 1)                             # probability of south / west
    inner = j < width - 1

    # candidate entries in ascending column order: west, south, north, east
    cols = np.stack([idx - width - 1, idx - 1, idx + 1, idx + width], axis=1)
    vals = np.stack([down, down,
                     np.where(i == 0, up + up, up),            # boundary moves are reflected:
                     np.where(j == 0, up + up, up)], axis=1)   # the COO duplicate is summed
    keep = np.stack([i > 0, j > 0, inner, inner], axis=1)

    counts = keep.sum(axis=1)
    indptr = np.concatenate(([0], np.cumsum(counts)))
    it = _index_dtype(int(indptr[-1]), n)
    return sp.csr_matrix((vals[keep], cols[keep].astype(it), indptr.astype(it)), shape=(n, n))


def laplace(n, dtype=None):
    """1-D Laplacian: -2 on the diagonal, 1 beside it (matrices.py:87-95)."""
    off = np.ones(n - 1, dtype=dtype)
    return sp.diags_array([-2 * np.ones(n, dtype=dtype), off, off], offsets=[0, -1, 1])


def laplace_eigen(n):
    """Eigenvalues of ``laplace(n)`` (matrices.py:76-84)."""
    return -2 + 2 * np.cos(np.arange(1, n + 1) * np.pi / (n + 1))


def lap2d(N):
    """2-D 5-point Laplacian on an N x N grid, ``kron(I, T) + kron(T, I)`` with
    T = tridiag(-1, 2, -1): n = N^2 rows, 5 N^2 - 4 N entries, float64, sorted CSR.
    Assembled directly (BASELINE config 2 is N = 4096, 83.9 M entries)."""
    n = N * N
    idx = np.arange(n)
    gi, gj = idx // N, idx % N
    keep = np.stack([gi > 0, gj > 0, np.ones(n, bool), gj < N - 1, gi < N - 1], axis=1)
    counts = keep.sum(axis=1)
    indptr = np.concatenate(([0], np.cumsum(counts)))
    it = _index_dtype(int(indptr[-1]), n)
    cols = np.stack([idx - N, idx - 1, idx, idx + 1, idx + N], axis=1).astype(it)[keep]
    vals = np.broadcast_to(np.array([-1.0, -1.0, 4.0, -1.0, -1.0]), (n, 5))[keep]
    return sp.csr_matrix((vals, cols, indptr.astype(it)), shape=(n, n))


def lap2d_rows(N, row0, row1):
    """Rows [row0, row1) of ``lap2d(N)`` as a ``RowBlock`` (global column ids), without building
    the rest: what a rank generates for itself when the whole operator would not fit one host
    (BASELINE config 5: n up to 2e8).  Bit-identical to slicing ``lap2d(N)``."""
    from .distributed import RowBlock
    n = N * N
    idx = np.arange(row0, row1, dtype=np.int64)
    gi, gj = idx // N, idx % N
    keep = np.stack([gi > 0, gj > 0, np.ones(idx.shape[0], bool), gj < N - 1, gi < N - 1], axis=1)
    counts = keep.sum(axis=1)
    indptr = np.concatenate(([0], np.cumsum(counts)))
    it = _index_dtype(5 * n, n)
    cols = np.stack([idx - N, idx - 1, idx, idx + 1, idx + N], axis=1).astype(it)[keep]
    vals = np.broadcast_to(np.array([-1.0, -1.0, 4.0, -1.0, -1.0]), (idx.shape[0], 5))[keep]
    return RowBlock(indptr.astype(it), cols, vals, int(row0), (n, n))


# ----------------------------------------------------------------------------------------
# BASELINE config 4: synthetic nonsymmetric CSR with power-law row lengths, generated
# shard by shard with a counter-based hash so that ANY block of rows can be regenerated
# bit-identically on any rank (or on the host at small n for partition checks).
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(x):
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def _unit(h):
    """uint64 hash -> float64 in [0, 1)."""
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


POWERLAW_TOP = 16          # rows carrying the separated leading eigenvalues


def powerlaw_rows(n, row0, row1, *, seed=0, band=4096, far_prob=1.0 / 32.0, lmin=4, lmax=2048,
                  alpha=1.35, chunk=1 << 20, top_base=3.0, top_step=0.25):
    """Rows [row0, row1) of the n x n synthetic operator, as a ``RowBlock`` (global columns).

    * row length  L_r = clip(floor(lmin * u^(-1/alpha)), lmin, lmax): Pareto, mean ~ 15
    * entry 0 is the diagonal; entry k > 0 sits at column (r + off) mod n with off uniform in
      [-band, band] (probability 1 - far_prob) or uniform over all columns (far_prob): a local
      band plus a sparse long-range tail.  Columns inside a row are NOT sorted and may repeat
      (scipy's csr_matvec and the device SpMV both sum entries in stored order).
    * off-diagonal values U(-1, 1) * 0.5 / L_r (row sums < 0.5); diagonal 1 + r/n except
      POWERLAW_TOP rows spread over the matrix whose diagonal is top_base + top_step i: with the
      defaults (3.0 + 0.25 i) the wanted largest-real-part eigenvalues are well separated and the
      solve converges in a few restarts; a small top_step clusters them (many restart cycles:
      the sustained-throughput variant of config 4).
    """
    from .distributed import RowBlock
    s0 = np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)
    ptr_parts, idx_parts, val_parts = [np.zeros(1, np.int64)], [], []
    total = 0
    top_rows = (np.arange(POWERLAW_TOP, dtype=np.int64) * (n // POWERLAW_TOP) + n // (2 * POWERLAW_TOP))
    for c0 in range(row0, row1, chunk):
        c1 = min(c0 + chunk, row1)
        r = np.arange(c0, c1, dtype=np.uint64)
        with np.errstate(over="ignore"):
            hr = _mix(r * np.uint64(0xD1342543DE82EF95) + s0)
        u = 1.0 - _unit(hr)                                   # (0, 1]
        L = np.clip(np.floor(lmin * u ** (-1.0 / alpha)), lmin, lmax).astype(np.int64)
        L = np.minimum(L, n)
        starts = np.concatenate(([0], np.cumsum(L)))
        nnz = int(starts[-1])
        row_of = np.repeat(np.arange(c1 - c0, dtype=np.int64), L)
        k = np.arange(nnz, dtype=np.int64) - starts[row_of]
        rr = r[row_of]
        with np.errstate(over="ignore"):
            h1 = _mix(hr[row_of] + k.astype(np.uint64) * np.uint64(0x2545F4914F6CDD1D))
            h2 = _mix(h1 + np.uint64(0x632BE59BD9B4E019))
            h3 = _mix(h2 + np.uint64(0x9E3779B97F4A7C15))
        far = _unit(h2) < far_prob
        off = (h1 % np.uint64(2 * band + 1)).astype(np.int64) - band
        col = np.where(far, (h1 % np.uint64(n)).astype(np.int64),
                       (rr.astype(np.int64) + off) % n)
        val = (2.0 * _unit(h3) - 1.0) * (0.5 / L[row_of])
        diag = k == 0
        col[diag] = rr[diag].astype(np.int64)
        dval = 1.0 + rr[diag].astype(np.float64) / n
        is_top = np.isin(rr[diag].astype(np.int64), top_rows)
        if is_top.any():
            which = np.searchsorted(top_rows, rr[diag].astype(np.int64)[is_top])
            dval[is_top] = top_base + top_step * which
        val[diag] = dval
        idx_parts.append(col.astype(np.int32 if n < 2**31 else np.int64))
        val_parts.append(val)
        ptr_parts.append(starts[1:] + total)
        total += nnz
    indptr = np.concatenate(ptr_parts)
    if total < 2**31 - 1:
        indptr = indptr.astype(np.int32)
    return RowBlock(indptr, np.concatenate(idx_parts), np.concatenate(val_parts), row0, (n, n))


def powerlaw(n, **kw):
    """The whole operator as a scipy CSR matrix (small n: tests, single-GPU runs)."""
    b = powerlaw_rows(n, 0, n, **kw)
    return sp.csr_matrix((b.data, b.indices, b.indptr), shape=(n, n))


def lap2d_rect(nx, ny, wx=1.0, wy=0.75):
    """Anisotropic 5-point Laplacian on an ny x nx grid (row-major, x fastest):
    ``wy * kron(T_ny, I_nx) + wx * kron(I_ny, T_nx)``, T = tridiag(-1, 2, -1).

    Same sparsity, symmetry and real storage as ``lap2d`` but its eigenvalues
    ``wx (2 - 2 cos(i pi / (nx+1))) + wy (2 - 2 cos(j pi / (ny+1)))`` are simple when
    nx != ny and wx != wy, so restart counts and Ritz values of a Krylov-Schur solve are not
    decided by rounding noise (the square isotropic grid has double eigenvalues).  Used to pin
    config-2-family parity; sorted CSR, float64."""
    n = nx * ny
    idx = np.arange(n)
    gi, gj = idx // nx, idx % nx
    keep = np.stack([gi > 0, gj > 0, np.ones(n, bool), gj < nx - 1, gi < ny - 1], axis=1)
    counts = keep.sum(axis=1)
    indptr = np.concatenate(([0], np.cumsum(counts)))
    it = _index_dtype(int(indptr[-1]), n)
    cols = np.stack([idx - nx, idx - 1, idx, idx + 1, idx + nx], axis=1).astype(it)[keep]
    d = 2.0 * wx + 2.0 * wy
    vals = np.broadcast_to(np.array([-wy, -wx, d, -wx, -wy]), (n, 5))[keep]
    return sp.csr_matrix((vals, cols, indptr.astype(it)), shape=(n, n))


def lap2d_rect_eigenvalues(nx, ny, wx=1.0, wy=0.75):
    """All eigenvalues of ``lap2d_rect`` (unsorted)."""
    cx = 2 - 2 * np.cos(np.arange(1, nx + 1) * np.pi / (nx + 1))
    cy = 2 - 2 * np.cos(np.arange(1, ny + 1) * np.pi / (ny + 1))
    return (wy * cy[:, None] + wx * cx[None, :]).ravel()
