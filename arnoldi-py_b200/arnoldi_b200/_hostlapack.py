"""LAPACK entry points for the host "rotate" step, taken from SciPy and called from C.

``scipy.linalg.cython_lapack`` exports the addresses of the LAPACK routines SciPy itself
calls (``__pyx_capi__``: PyCapsules holding C function pointers with Fortran-style by-reference
arguments).  They are handed to ``ab200_host_schur`` / ``ab200_host_reorder``
(csrc/hostschur.cu), which issue the reference's call sequence -- zgees('V', 'N') with a
queried workspace exactly like ``scipy.linalg.schur(output="complex")``, then one ztrexc per
target slot like ``ordered_schur`` (utils.py:52-63) -- without the f2py wrappers, the NumPy
temporaries and the interpreter between the calls.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _capsule_pointer(name):
    from scipy.linalg import cython_lapack
    cap = cython_lapack.__pyx_capi__[name]
    api = C.pythonapi
    api.PyCapsule_GetName.restype = C.c_char_p
    api.PyCapsule_GetName.argtypes = [C.py_object]
    api.PyCapsule_GetPointer.restype = C.c_void_p
    api.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
    ptr = api.PyCapsule_GetPointer(cap, api.PyCapsule_GetName(cap))
    if not ptr:
        raise ImportError(f"no address for LAPACK routine {name}")
    return C.c_void_p(ptr)


# max |H - H^T| / max |H| below which a real H_m is treated as symmetric (fast_real_schur only)
SYM_TOL = 1e-12


class NativeRotate:
    def __init__(self):
        self.lib = _lib.load()
        self.zgees = _capsule_pointer("zgees")
        self.ztrexc = _capsule_pointer("ztrexc")
        self.dgees = _capsule_pointer("dgees")
        try:
            self.dsyevd = _capsule_pointer("dsyevd")
        except Exception:
            self.dsyevd = None

    def __call__(self, Hm, sort_function, real_ok=False):
        m = Hm.shape[0]
        T = np.array(Hm, dtype=np.complex128, order="F", copy=True)
        Q = np.empty((m, m), np.complex128, order="F")
        if real_ok and not np.any(T.imag):
            if self.dsyevd is not None:
                # symmetric H_m (symmetric operator): eigendecomposition, ordered by a permutation
                info = self.lib.ab200_host_eigh_real(self.dsyevd, m, T.ctypes.data_as(C.c_void_p),
                                                     Q.ctypes.data_as(C.c_void_p), SYM_TOL)
                if info == 0:
                    perm = np.ascontiguousarray(sort_function(np.diag(T)), dtype=np.int64)
                    assert perm.shape == (m,), "sort_function must return a permutation of all m indices"
                    return (np.asfortranarray(np.diag(np.diag(T)[perm])),
                            np.asfortranarray(Q[:, perm]))
                if info != 1:
                    raise np.linalg.LinAlgError(f"dsyevd failed with info = {info}")
            info = self.lib.ab200_host_schur_real(self.dgees, m, T.ctypes.data_as(C.c_void_p),
                                                  Q.ctypes.data_as(C.c_void_p))
            if info != 0:
                raise np.linalg.LinAlgError(f"dgees failed with info = {info}")
        else:
            info = self.lib.ab200_host_schur(self.zgees, m, T.ctypes.data_as(C.c_void_p),
                                             Q.ctypes.data_as(C.c_void_p), None)
            if info != 0:
                raise np.linalg.LinAlgError(f"zgees failed with info = {info}")
        perm = np.ascontiguousarray(sort_function(np.diag(T)), dtype=np.int64)
        assert perm.shape == (m,), "sort_function must return a permutation of all m indices"
        Q2 = np.empty((m, m), np.complex128, order="F")
        info = self.lib.ab200_host_reorder(self.ztrexc, m, T.ctypes.data_as(C.c_void_p),
                                           Q2.ctypes.data_as(C.c_void_p),
                                           perm.ctypes.data_as(C.c_void_p), None)
        if info != 0:
            raise np.linalg.LinAlgError(f"ztrexc failed with info = {info}")
        return T, Q @ Q2          # krylov_schur.py:72
