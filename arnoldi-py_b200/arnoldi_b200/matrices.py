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
