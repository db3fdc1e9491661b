"""Orthonormalisation plugs on host arrays: drop-ins for ``arnoldi.ortho.dgks_gs`` / ``dgks_mgs``.

Reference: src/arnoldi/ortho.py:56-107 and :9-53.  Same contract
``f(w, V, h, tol=1e-8, eta=sqrt(1/2)) -> (beta, breakdown)`` with ``w`` and ``h`` updated
in place.  These wrappers upload ``V`` for every call -- they exist so the plugs can be
exercised and compared in isolation; inside ``partial_schur`` the same kernels run on
the resident basis and the choice is made with ``ortho="cgs2" | "mgs"``.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .solver import DeviceSolver


def _run(kind, w, V, h, tol, eta, device, options=None):
    n, c = V.shape
    assert w.shape == (n,) and h.shape[0] >= c
    wbuf = np.ascontiguousarray(w, dtype=np.complex128)
    hbuf = np.zeros(c, np.complex128)
    with DeviceSolver(n, max(c, 1), device=device) as dev:
        for key, value in (options or {}).items():
            dev.set_option(key, value)
        dev.set_columns(0, V)
        beta, broke = dev.ortho(c, wbuf, hbuf, tol, eta, kind)
    w[:] = wbuf
    h[:c] = hbuf
    return beta, broke


def dgks_gs(w, V, h, tol=1e-8, eta=np.sqrt(0.5), *, device=0, options=None):
    """Classical Gram-Schmidt, repeated once when the DGKS test fires (ortho.py:56-107)."""
    return _run(_lib.ORTHO_CGS2, w, V, h, tol, eta, device, options)


def dgks_mgs(w, V, h, tol=1e-8, eta=np.sqrt(0.5), *, device=0):
    """Modified Gram-Schmidt, swept a second time when the DGKS test fires (ortho.py:9-53)."""
    return _run(_lib.ORTHO_MGS, w, V, h, tol, eta, device)
