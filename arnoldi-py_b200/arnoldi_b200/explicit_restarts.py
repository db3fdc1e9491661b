"""Explicitly restarted Arnoldi with deflation on the device kernels: drop-in for
``arnoldi.explicit_restarts.explicit_restarts_with_deflation``.

Reference: src/arnoldi/explicit_restarts.py:80-168.  The reference finds
the wanted eigenpairs one after the other: for pair k it runs Arnoldi expansions from column k
(so the new vectors are orthogonalised against the k locked vectors as well), replaces column
k by the leading Ritz vector of the active block, re-orthogonalises it against the locked
columns with a plain modified Gram-Schmidt sweep (``mgs``, :63-77) and repeats until the
residual estimate of that Ritz pair is below tol; the column of H for the locked vector is then
recomputed by explicit projection (:150-151).

Everything n-length runs in the kernels ``partial_schur`` uses: ``ab200_expand`` (SpMV +
CGS2/DGKS), ``ab200_combine`` (Ritz vector out of the active columns), ``ab200_orthonormalize_
column`` (the MGS deflation sweep + normalisation) and ``ab200_project`` (one SpMV + one
fused dot sweep instead of the reference's k + 1 separate ``A @ v`` products).  The
(m - k) x (m - k) eigenproblem and the bookkeeping stay NumPy on the host, as in the reference.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .history import History
from .operator import as_csr, credit_matvecs, is_device_operator, unwrap
from .solver import DeviceSolver
from .utils import arg_largest_magnitude, rand_normalized_vector


def explicit_restarts_with_deflation(
    A, nev, *, max_dim=None, stopping_criterion=None, max_restarts=100,
    sort_function=None,
    device=0, stats=None,
):
    """Returns ``(eigenvalues, eigenvectors, history)`` like the reference: ``eigenvalues``
    (nev,) complex128, ``eigenvectors`` (n, nev) complex128 from the final projected block
    (explicit_restarts.py:166-168)."""
    if stopping_criterion is None:
        tol = np.sqrt(np.finfo(A.dtype).eps)
    else:
        tol = stopping_criterion
    if sort_function is None:
        sort_function = arg_largest_magnitude
    assert max_restarts > 0
    n = A.shape[0]
    assert A.shape[1] == n
    if max_dim is None:
        max_dim = min(max(2 * nev + 1, 20), n)

    H = np.zeros((max_dim + 1, max_dim), dtype=np.complex128)
    history = History.from_k(nev)
    wrappers = []
    with DeviceSolver(n, max_dim, device=device) as dev:
        if stats is not None:
            dev.set_timing(True)
        if is_device_operator(A):
            dev.set_operator(A)
        else:
            _, wrappers = unwrap(A)
            dev.set_csr(*as_csr(A)[:3])

        def deflate(col):
            """``mgs(V[:, :col], V[:, col], tol)`` (explicit_restarts.py:63-77): one modified
            Gram-Schmidt sweep against the locked columns, then normalise."""
            beta = dev.orthonormalize_column(col, col, tol, eta=0.0, ortho=_lib.ORTHO_MGS)
            assert beta > tol, "MGS: Too small norm when orthornormalizing"

        for k in range(nev):
            dev.set_columns(k, rand_normalized_vector(n, np.complex128))    # :110-112
            deflate(k)
            for restart in range(max_restarts):
                cols, m, _ = dev.expand(k, max_dim, tol)                     # :115-117
                for j in range(k, m):
                    rows = j + 2 if (j + 1 < m or m == max_dim) else j + 1
                    H[:rows, j] = cols[:rows, j]
                assert m > k
                happy_breakdown = m != max_dim
                matvecs = restart * (max_dim - k) + (m - k)                  # :127

                # Ritz pairs of the active block (decomposition.py:115-129)
                Hk = H[k:, k:]
                w, S = np.linalg.eig(Hk[: m - k, : m - k])
                ind = sort_function(w)[: m - k]
                S = S[:, ind]
                values = w[ind]
                approximate_residuals = np.abs(Hk[m - k, m - k - 1] * S[-1])

                dev.combine(S[:, :1], k, m - k, 1)                           # v_k = V_k S[:, 0]
                deflate(k)                                                   # :140-141
                approximate_convergence = approximate_residuals / np.abs(values)
                if happy_breakdown or approximate_convergence[0] < tol:
                    H[: k + 1, k] = dev.project(k, k + 1)                    # :150-151
                    H[k + 1:-1, k] = 0
                    history.matvecs[k] = matvecs
                    history.restarts[k] = restart + 1
                    break
            else:
                raise ValueError(f"Could not converge for value {k}")

        eivals, Y = np.linalg.eig(H[:nev, :nev])                             # :166
        dev.combine(Y, 0, nev, nev)                                          # V[:, :nev] @ Y
        eivecs = dev.get_columns(0, nev)
        if stats is not None:
            stats.update(dev.stats())
            stats["true_matvecs"] = stats["arnoldi_steps"]
        credit_matvecs(wrappers, dev.true_matvecs())
    return eivals, eivecs, history
