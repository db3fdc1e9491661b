"""Turning what callers pass as ``A`` into the CSR block the device SpMV consumes.

The reference only needs ``A.shape``, ``A.dtype`` and ``A @ x`` (decomposition.py:44,58)
and is called with scipy CSR matrices (README.md:29), dense ndarrays
(tests/test_krylov_schur.py:36,43) and LinearOperators (scripts/utils.py:162-172: its own
timing harness wraps the matrix in ``MatvecCounter(LinearOperator)`` before calling
``partial_schur``).  The device path needs the entries of A, so

* a scipy sparse matrix / ndarray is used directly;
* a wrapper that exposes the wrapped matrix as ``.A`` (the reference's ``MatvecCounter``,
  scipy's ``aslinearoperator(M)``, and wrappers of wrappers) is unwrapped, and its ``matvecs``
  counter, when it has one, is credited with the true number of operator applications after
  the solve (``credit_matvecs``);
* an object with a ``device_apply`` method runs as a device operator (``DeviceOperator``
  protocol, ``arnoldi_b200.device_operator``);
* anything else (an opaque callable) raises ``TypeError``: there is no host path for n-length
  work.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

_MAX_UNWRAP = 8


def unwrap(A):
    """Follow ``.A`` attributes down to a scipy sparse matrix or ndarray.

    Returns ``(matrix, wrappers)``; ``wrappers`` lists the objects peeled off, outermost first.
    """
    wrappers = []
    M = A
    for _ in range(_MAX_UNWRAP):
        if sp.issparse(M) or isinstance(M, np.ndarray):
            return M, wrappers
        inner = getattr(M, "A", None)
        if inner is None or inner is M:
            break
        wrappers.append(M)
        M = inner
    if sp.issparse(M) or isinstance(M, np.ndarray):
        return M, wrappers
    raise TypeError(
        f"{type(A).__name__} has no stored entries: the device path needs a scipy sparse matrix, "
        "an ndarray, a wrapper exposing one as `.A` (e.g. the reference's MatvecCounter), or an "
        "object implementing the DeviceOperator protocol (`device_apply`); an opaque "
        "LinearOperator / callable cannot be applied on the GPU (see INTEGRATION.md)")


def credit_matvecs(wrappers, count):
    """Add the true operator-application count to every unwrapped counter
    (scripts/utils.py:55-68 increments ``matvecs`` once per ``A @ x``)."""
    for w in wrappers:
        cur = getattr(w, "matvecs", None)
        if isinstance(cur, (int, np.integer)) and not isinstance(cur, bool):
            try:
                w.matvecs = int(cur) + int(count)
            except AttributeError:
                pass


def is_device_operator(A):
    return callable(getattr(A, "device_apply", None))


def as_csr(A):
    """Return (indptr, indices, data, shape) in scipy's canonical CSR layout.

    A scipy CSR matrix is passed through untouched (same entry order, duplicates and
    explicit zeros included) so the device SpMV walks each row exactly as
    ``csr_matvec`` would.  Index arrays keep scipy's width (int32 below 2^31 entries).
    """
    A, _ = unwrap(A)
    if sp.issparse(A):
        M = A if A.format == "csr" else A.tocsr()
    else:
        assert A.ndim == 2, "A must be two-dimensional"
        M = sp.csr_matrix(A)
    if M.dtype.kind not in "fc":
        M = M.astype(np.float64)
    data = M.data
    if data.dtype not in (np.float64, np.complex128):
        data = data.astype(np.complex128 if data.dtype.kind == "c" else np.float64)
    return M.indptr, M.indices, data, M.shape
