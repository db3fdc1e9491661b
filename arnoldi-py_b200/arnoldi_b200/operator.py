"""Turning what callers pass as ``A`` into the CSR block the device SpMV consumes.

The reference only needs ``A.shape``, ``A.dtype`` and ``A @ x`` (decomposition.py:44,58)
and is called with scipy CSR matrices (README.md:29), dense ndarrays
(tests/test_krylov_schur.py:36,43) and LinearOperators (scripts/utils.py:162-172).
The device path needs the entries of A, so an opaque LinearOperator cannot run here.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def as_csr(A):
    """Return (indptr, indices, data, shape) in scipy's canonical CSR layout.

    A scipy CSR matrix is passed through untouched (same entry order, duplicates and
    explicit zeros included) so the device SpMV walks each row exactly as
    ``csr_matvec`` would.  Index arrays keep scipy's width (int32 below 2^31 entries).
    """
    if sp.issparse(A):
        M = A if A.format == "csr" else A.tocsr()
    elif isinstance(A, np.ndarray):
        assert A.ndim == 2, "A must be two-dimensional"
        M = sp.csr_matrix(A)
    else:
        raise TypeError(
            f"{type(A).__name__} has no stored entries: the device path needs a scipy sparse "
            "matrix or an ndarray (an opaque LinearOperator / callable cannot be applied on the "
            "GPU; see INTEGRATION.md)")
    if M.dtype.kind not in "fc":
        M = M.astype(np.float64)
    data = M.data
    if data.dtype not in (np.float64, np.complex128):
        data = data.astype(np.complex128 if data.dtype.kind == "c" else np.float64)
    return M.indptr, M.indices, data, M.shape
