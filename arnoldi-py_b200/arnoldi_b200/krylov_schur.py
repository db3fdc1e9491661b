"""Krylov-Schur driver: drop-in for ``arnoldi.krylov_schur.partial_schur``.

Reference: src/arnoldi/krylov_schur.py:10-114.  Same signature, defaults, assertions,
exceptions and return shapes.  The n x (max_dim+1) basis lives in HBM for the whole
solve; per restart the host sees only the (max_dim+1) x max_dim matrix H.

    expand  (device)  decomposition.py:56-66   SpMV + CGS2/DGKS (or MGS) per column
    rotate  (host)    krylov_schur.py:69-76    zgees, ordered Schur, Q = Q1 Q2
    truncate(device)  krylov_schur.py:78,81    V[:, :p] = V Q_p ; V[:, p] = V[:, m]
    test    (host)    krylov_schur.py:83-101   spike row, residual estimates, history
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import schur

from . import _lib
from .history import History
from .operator import as_csr
from .solver import DeviceSolver
from .utils import arg_largest_magnitude, ordered_schur, rand_normalized_vector

_ORTHO = {"cgs2": _lib.ORTHO_CGS2, "dgks_gs": _lib.ORTHO_CGS2, "mgs": _lib.ORTHO_MGS,
          "dgks_mgs": _lib.ORTHO_MGS}


def _as_scipy_csr(A):
    import scipy.sparse as sp
    indptr, indices, data, shape = as_csr(A)
    return sp.csr_matrix((data, indices, indptr), shape=shape, copy=False)


def _ortho_kind(ortho):
    if callable(ortho):
        ortho = getattr(ortho, "__name__", "")
    try:
        return _ORTHO[ortho]
    except KeyError:
        raise AssertionError(f"unknown orthonormalization {ortho!r}: use 'cgs2' or 'mgs'") from None


def partial_schur(
    A, nev, *, max_dim=None, stopping_criterion=None, max_restarts=100,
    sort_function=None, p=None,
    ortho="cgs2", v0=None, device=0, stats=None, raise_on_no_convergence=True, comm=None,
    halo="auto", real_storage=True,
):
    """Partial Schur decomposition ``A Q = Q T`` of the ``nev`` wanted eigenvalues.

    Positional / keyword arguments up to ``p`` are the reference's.  Extra keyword-only
    arguments (defaults reproduce the reference):

    ortho : "cgs2" (dgks_gs, the reference's hard-wired choice) or "mgs" (dgks_mgs)
    v0    : start vector; default draws ``rand_normalized_vector`` like the reference
    device: CUDA device ordinal
    stats : dict filled with true matvec count, DGKS rounds, per-kernel time/bytes
    raise_on_no_convergence : False returns the current (Q, T, history) instead of raising
        (used by the benchmark to time a bounded number of restart cycles)
    comm  : a ``distributed.TorchComm`` (one process per GPU).  ``A`` is then either the whole
        scipy CSR matrix (each rank slices its block of rows) or this rank's ``RowBlock``; every
        rank must seed NumPy's global RNG identically (v0 is drawn globally and sliced); the
        returned Q holds this rank's rows only, T and history are identical on all ranks.
    real_storage : keep the basis as float64 on the device for as long as it is provably real
        (float64 A, real v0, real Schur vectors so far -- e.g. every symmetric real operator,
        and the first expansion of any real one); it is converted to complex128 in place the
        moment a complex Q arrives.  Real parts are what the complex path would compute; the
        returned arrays are complex128 either way.  False forces complex128 storage throughout.
    halo  : "pull" (each rank gathers the remote entries of v it needs straight from peer HBM),
        "push" (the owner gathers locally and streams them into the peer's buffer) or "auto"
        (push when some rank's halo has more than 65 536 scattered entries)

    Returns ``(Q, T, history)``: Q (n, nev) complex128, T (nev, nev) complex128.
    """
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
    if p is None:
        p = min(nev + 5, max_dim - 1)
    assert nev <= p < max_dim
    kind = _ortho_kind(ortho)

    import time
    clock = time.perf_counter
    phases = {}
    t_mark = clock()

    def lap(name):
        nonlocal t_mark
        now = clock()
        phases[name] = phases.get(name, 0.0) + (now - t_mark)
        t_mark = now

    H = np.zeros((max_dim + 1, max_dim), dtype=np.complex128)
    history = History.from_k(nev)
    converged = False
    multi = comm is not None and comm.world > 1
    if multi:
        from .distributed import RowBlock, RowPartition, build_halo_plan, slice_rows
        part = RowPartition(n, comm.world)
        r0, r1 = part.rows(comm.rank)
        block = A if isinstance(A, RowBlock) else slice_rows(_as_scipy_csr(A), r0, r1)
        assert block.row0 == r0 and block.nrows == r1 - r0, "RowBlock does not match the partition"
        plan = build_halo_plan(block)
        solver_args = dict(device=device, row0=r0, nrows_local=r1 - r0)
    else:
        indptr, indices, data, _ = as_csr(A)
        r0, r1 = 0, n
        solver_args = dict(device=device)

    lap("host_prepare")
    with DeviceSolver(n, max_dim, **solver_args) as dev:
        if not real_storage:
            dev.set_option("real_mode", 0)
        lap("device_alloc")
        if stats is not None:
            dev.set_timing(True)
        if multi:
            dev.connect(comm, part)
            dev.set_halo(plan.ghost_cols)
            assert halo in ("auto", "pull", "push")
            biggest = max(int.from_bytes(b, "little") for b in comm.all_gather_bytes(
                int(plan.ghost_cols.shape[0]).to_bytes(8, "little")))
            if halo == "push" or (halo == "auto" and biggest > 65536):
                dev.connect_halo_push(comm, part, plan.ghost_cols)
            dev.set_csr(plan.indptr, plan.indices, plan.data)
        else:
            dev.set_csr(indptr, indices, data)
        lap("upload_csr")
        if v0 is None:
            v0 = rand_normalized_vector(n, np.complex128)
        lap("host_v0")
        dev.set_columns(0, v0[r0:r1])
        # (multi-GPU: ab200_expand opens with a device-side peer barrier, so every rank's
        #  columns are in place before any halo read -- no host barrier needed here)

        def grow(start):
            cols, n_iter, _ = dev.expand(start, max_dim, tol, ortho=kind)
            for j in range(start, n_iter):
                H[: j + 2, j] = cols[: j + 2, j]
            return n_iter

        lap("upload_v0")
        m = grow(0)
        lap("expand")
        for restart in range(max_restarts):
            if m != max_dim:
                raise ValueError("Happy breakdown not supported yet")
            reported = restart * (max_dim - nev) + (m - nev)  # krylov_schur.py:63

            # rotate: zgees on H_m, then the reference's ordered_schur on the triangular result
            # (its second zgees is an exact no-op on triangular input and is skipped, its
            # ztrexc sequence is kept call for call so Q matches to rounding)
            T1, Q1 = schur(H[:m, :m], output="complex")
            T2, Q2 = ordered_schur(T1, output="complex", sort_function=sort_function)
            Q = Q1 @ Q2
            Qp = Q[:, :p]
            spike = H[m, :m] @ Qp
            last_beta = H[m, m - 1]

            lap("host_schur")
            # truncate
            dev.restart(Q, m, p)
            lap("restart")
            H[:p, :p] = T2[:p, :p]
            H[p, :p] = spike
            H[p, p:] = 0

            # convergence estimates |beta q_{m-1,k}| / |t_kk|
            estimate = np.abs(last_beta * Q[m - 1, :]) / np.abs(np.diag(T2))
            hit = estimate[:nev] <= tol
            history.matvecs[hit] = reported
            history.restarts[hit] = restart + 1
            if np.all(estimate[:nev] < tol):
                converged = True
                break
            m = grow(p)
            lap("expand")

        if stats is not None:
            stats.update(dev.stats())
            stats["true_matvecs"] = stats["arnoldi_steps"]
            stats["restart_cycles"] = restart + 1
            stats["converged"] = converged
        if not converged and raise_on_no_convergence:
            raise ValueError("Has not converged !")
        lap("host_schur")
        Qout = dev.get_columns(0, nev, hugepages=True)
        lap("download_q")
        if multi:
            comm.barrier()   # nobody unmaps while a peer may still be in its last kernel
    lap("device_free")
    if stats is not None:
        stats["host_phases_s"] = phases

    return Qout, H[:nev, :nev].copy(), history
