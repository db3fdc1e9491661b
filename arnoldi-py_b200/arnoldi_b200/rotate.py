"""The m x m "rotate" step of a Krylov-Schur restart (krylov_schur.py:69-72 of the reference):

    T1, Q1 = schur(H_m, output="complex")                       zgees
    T2, Q2 = ordered_schur(T1, "complex", sort_function)        (zgees again) + ztrexc moves
    Q = Q1 @ Q2

The reference's call sequence is kept -- which Schur vectors come out depends on it -- but the
GPUs are idle while it runs (1.5 ms of a 10 ms restart cycle on 8 GPUs), so it is executed
with as little interpreter and wrapper overhead as possible: the LAPACK routines are called
through the function pointers scipy exports (``scipy.linalg.cython_lapack``), from C, inside
``libarnoldi_b200.so`` (``ab200_host_rotate``, csrc/hostschur.cpp); only ``sort_function`` --
a Python callable by contract (utils.py:50) -- comes back to the interpreter, once per restart.
When the pointers cannot be obtained the SciPy wrappers are used, call for call as the
reference does.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import schur

from .utils import ordered_schur


def rotate_scipy(Hm, sort_function):
    T1, Q1 = schur(Hm, output="complex")
    T2, Q2 = ordered_schur(T1, output="complex", sort_function=sort_function)
    return T2, Q1 @ Q2


_native = None


def _load_native():
    global _native
    if _native is None:
        try:
            from . import _hostlapack
            _native = _hostlapack.NativeRotate()
        except Exception:
            _native = False
    return _native


def rotate(Hm, sort_function, fast_real=False):
    """Ordered complex Schur form of ``Hm``: returns ``(T2, Q)`` with ``Hm = Q T2 Q^H`` and
    ``diag(T2)`` in the order ``sort_function`` asks for.

    ``fast_real``: when ``Hm`` has no imaginary part, factor it with dgees (a third of zgees's
    arithmetic) and rotate the 2 x 2 blocks to triangular form.  The result is a valid ordered
    Schur form but not bit-for-bit the reference's (signs / phases of the vectors, rounding),
    so the driver only asks for it where it has left the reference's arithmetic anyway."""
    nat = _load_native()
    if nat:
        return nat(Hm, sort_function, real_ok=fast_real)
    return rotate_scipy(Hm, sort_function)
