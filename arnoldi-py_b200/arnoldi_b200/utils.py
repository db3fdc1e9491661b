"""m-by-m control logic kept on the host (reference: src/arnoldi/utils.py).

The projected problem is at most ~100 x 100, <0.1% of the solve time, and its
LAPACK call sequence decides which Schur vectors come out -- so it is kept as the
same LAPACK calls in the same order (zgees, then one ztrexc move per target slot)
rather than re-derived on the device.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import schur
from scipy.linalg.lapack import ctrexc, ztrexc


def rand_normalized_vector(n, dtype=np.float64):
    """Unit-norm start vector from NumPy's legacy global RNG (utils.py:7-13).

    Drawing from the same stream in the same way is what makes a seeded run of the
    drop-in start from the same v0 as a seeded run of the reference.
    """
    v = np.random.randn(n).astype(dtype)
    v /= np.linalg.norm(v)
    return v


def arg_largest_magnitude(x):
    """Permutation that sorts by decreasing modulus (utils.py:16-17)."""
    return np.argsort(-np.abs(x))


def arg_largest_real(x):
    """Permutation that sorts by decreasing real part (utils.py:20-21)."""
    return np.argsort(-np.real(x))


_SWAPPERS = {np.dtype(np.complex64): ctrexc, np.dtype(np.complex128): ztrexc}


def ordered_schur(a, output="real", *, sort_function=None):
    """Complex Schur form whose diagonal follows ``sort_function`` (utils.py:32-67).

    ``sort_function(diag(T))`` yields the wanted order as indices into the unsorted
    diagonal; entry by entry the wanted eigenvalue is moved up to its slot with one
    LAPACK ``?trexc`` call, while ``slots`` remembers where every original entry
    currently sits.  Only ``output="complex"`` is implemented, as in the reference.
    """
    if output != "complex":
        raise ValueError("output!='complex' not implemented yet")
    if sort_function is None:
        sort_function = arg_largest_magnitude
    a = np.asarray(a)
    if a.dtype.kind == "c" and not np.tril(a, -1).any():
        # The reference re-factors its input here (utils.py:45) although partial_schur passes
        # the already-triangular T1 (krylov_schur.py:70).  LAPACK's zgees returns an exactly
        # triangular input unchanged with Z = I (bit for bit; test_host_logic checks it), so
        # the call is skipped: 0.3-0.5 ms per restart that the GPUs would spend idle.
        T, Z = np.array(a, order="F", copy=True), np.eye(a.shape[0], dtype=a.dtype, order="F")
    else:
        T, Z = schur(a, output="complex")
        T, Z = np.asfortranarray(T), np.asfortranarray(Z)
    swap = _SWAPPERS[np.result_type(a.dtype, 1j)]
    slots = list(range(T.shape[0]))
    for dest, original in enumerate(sort_function(np.diag(T))):
        here = slots.index(original)
        if here == dest:
            continue
        # T and Z are this function's own Fortran-ordered arrays: let LAPACK work in place
        T, Z, _ = swap(T, Z, here + 1, dest + 1, overwrite_a=1, overwrite_q=1)  # 1-based
        slots.insert(dest, slots.pop(here))
    return T, Z
