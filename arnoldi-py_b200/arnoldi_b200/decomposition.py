"""Arnoldi expansion on host arrays: drop-in for ``arnoldi.decomposition.arnoldi_decomposition``.

Reference: src/arnoldi/decomposition.py:13-68.  The reference's tests drive this
function directly with caller-owned ``V`` and ``H`` (tests/test_decomposition.py:81-87),
so the boundary takes host arrays in either memory order, runs the expansion on the
device and writes the new columns back in place.  ``partial_schur`` does not go through
here: it keeps the basis resident on the device between expansions.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .krylov_schur import _ortho_kind
from .operator import as_csr
from .solver import DeviceSolver


def arnoldi_decomposition(A, V, H, invariant_tol=None, *, start_dim=0, max_dim=None,
                          ortho="cgs2", device=0):
    """Grow the Arnoldi relation ``A V_m = V_{m+1} H`` from ``start_dim`` to ``max_dim``.

    Returns ``(V[:, :k+1], H[:k+1, :k], k)`` with ``k == max_dim`` unless a step broke
    down (new vector norm below ``invariant_tol``), in which case ``k`` is that step + 1,
    ``H[k, k-1]`` is left untouched and ``V[:, k]`` is left un-normalised.
    """
    if invariant_tol is None:
        invariant_tol = np.sqrt(np.finfo(A.dtype).eps)

    n = A.shape[0]
    m = V.shape[1] - 1
    assert A.shape[1] == n, "A is expected to be square matrix"
    assert V.shape == (n, m + 1), "V must have the same number of rows as A"
    assert H.shape == (m + 1, m), f"H must be {m + 1, m}, is {H.shape}"
    if max_dim is None:
        max_dim = m
    assert max_dim <= m, "max_dim > m violated"
    assert 0 <= start_dim <= max_dim
    kind = _ortho_kind(ortho)

    indptr, indices, data, _ = as_csr(A)
    with DeviceSolver(n, m, device=device) as dev:
        dev.set_csr(indptr, indices, data)
        dev.set_columns(0, V[:, : start_dim + 1])
        cols, k, broke = dev.expand(start_dim, max_dim, invariant_tol, ortho=kind)
        for j in range(start_dim, k):
            rows = j + 1 if (broke and j == k - 1) else j + 2
            H[:rows, j] = cols[:rows, j]
        if k > start_dim:
            V[:, start_dim + 1 : k + 1] = dev.get_columns(start_dim + 1, k - start_dim)
    return V[:, : k + 1], H[: k + 1, :k], k
