"""Krylov-Schur driver: drop-in for ``arnoldi.krylov_schur.partial_schur``.

Reference: src/arnoldi/krylov_schur.py:10-114.  Same signature, defaults, assertions,
exceptions and return shapes.  The n x (max_dim+1) basis lives in HBM for the whole
solve; per restart the host sees only the (max_dim+1) x max_dim matrix H.

    expand  (device)  decomposition.py:56-66   SpMV + CGS2/DGKS (or MGS) per column
    rotate  (host)    krylov_schur.py:69-76    zgees, ordered Schur, Q = Q1 Q2
    truncate(device)  krylov_schur.py:78,81    V[:, :p] = V Q_p ; V[:, p] = V[:, m]
    test    (host)    krylov_schur.py:83-101   spike row, residual estimates, history

Beyond the reference (all keyword-only, all defaulting to the reference's behaviour):
real arithmetic for real operators (``real_arith``), happy-breakdown handling
(``on_breakdown``), locking and a dynamic restart size (``lock``, ``dynamic_p``) -- the rows
SURVEY.md section 8f lists as "next" (README.md:116-118 of the reference lists them as TODO).
"""
from __future__ import annotations

import threading
import time

import numpy as np

from . import _lib
from .history import History
from .operator import as_csr, credit_matvecs, is_device_operator, unwrap
from .rotate import rotate
from .solver import DeviceSolver
from .utils import arg_largest_magnitude, rand_normalized_vector

_ORTHO = {"cgs2": _lib.ORTHO_CGS2, "dgks_gs": _lib.ORTHO_CGS2, "mgs": _lib.ORTHO_MGS,
          "dgks_mgs": _lib.ORTHO_MGS}

MAX_DIM_LIMIT = 256   # ab200_create: reduction slots / column tiles are sized for this


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
        raise AssertionError(
            f"unknown orthonormalization {ortho!r}: the device path offers the reference's two "
            "plugs, 'cgs2' (dgks_gs) and 'mgs' (dgks_mgs); an arbitrary Python callable cannot "
            "run inside the device loop") from None


def real_invariant_basis(Qp):
    """Real orthonormal basis of ``span(Qp)`` when that subspace is closed under complex
    conjugation, else ``None``.

    ``Qp`` (m x p) holds the leading Schur vectors of a REAL matrix.  The span is
    conjugation-closed exactly when the cut at p does not separate a complex-conjugate pair
    of Ritz values; then ``P = Qp Qp^H`` is a real projector and
    ``[Re Qp, Im Qp] [Re Qp, Im Qp]^T = Re P = P``: the p non-zero singular values of
    ``[Re Qp, Im Qp]`` are all 1 and its leading left singular vectors are the basis.  When
    a pair is split, singular values strictly between 0 and 1 appear.
    """
    m, p = Qp.shape
    U, s, _ = np.linalg.svd(np.hstack([Qp.real, Qp.imag]), full_matrices=False)
    closed = s[p - 1] > 1.0 - 1e-8 and (s.shape[0] == p or s[p] < 1e-8)
    return np.ascontiguousarray(U[:, :p]) if closed else None


def partial_schur(
    A, nev, *, max_dim=None, stopping_criterion=None, max_restarts=100,
    sort_function=None, p=None,
    ortho="cgs2", v0=None, device=0, stats=None, raise_on_no_convergence=True, comm=None,
    halo="auto", real_storage=True, real_arith="lossless", on_breakdown="raise", lock=False,
    dynamic_p=False, spmv_algo="auto", fast_real_schur=False,
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
        rank must seed NumPy's global RNG identically (v0 is drawn globally and sliced; the
        ranks cross-check it); the returned Q holds this rank's rows only, T and history are
        identical on all ranks.
    real_storage : False forces complex128 storage throughout (same as ``real_arith="off"``).
    real_arith : how long the basis of a REAL operator (float64 A, real v0) is kept as float64
        on the device (half the bytes of every n-length kernel; the returned arrays are
        complex128 either way).
          "lossless" (default; "auto" is an alias): float64 storage for as long as the complex
                    arithmetic of the reference would produce exactly-zero imaginary parts, i.e.
                    until a Schur basis Q with a non-zero imaginary part is applied; from then
                    on complex128.  The iteration is the reference's, restart for restart
                    (symmetric operators stay real for the whole solve; nonsymmetric ones
                    usually turn complex at their first restart).
          "pairs" : real arithmetic throughout (README.md:118 of the reference lists it as
                    TODO).  At each restart a real orthonormal basis Z of the kept subspace
                    ``span(Q[:, :p])`` is applied instead of the complex Schur vectors -- the
                    same Krylov-Schur iteration in another basis (``real_invariant_basis``).
                    That needs the subspace to be closed under conjugation; when the cut at p
                    separates a complex-conjugate pair of Ritz values the pair is kept whole
                    (p + 1 vectors for that restart; p - 1 when p + 1 == max_dim).  This is NOT
                    the reference's iteration: restart counts differ (usually fewer matvecs);
                    converged Ritz values agree; T is the same triangular factor up to the
                    phases of the Schur vectors.
          "off"   : complex128 storage from the start.
    on_breakdown : "raise" -- ``ValueError("Happy breakdown not supported yet")`` exactly like
        krylov_schur.py:57-59 -- or "continue": the invariant subspace found so far is kept,
        H[m, m-1] is set to zero, a fresh random direction orthonormalised against the basis is
        appended on the device and the expansion goes on (a Krylov decomposition with a
        block-triangular H); when the basis spans the whole space the solve ends there.
    lock  : True deflates a leading wanted Ritz pair once its residual estimate is below
        ``tol`` by zeroing its entry of the spike row (it stays in the basis, so later vectors
        are still orthogonalised against it).  The reference never locks.
    dynamic_p : True keeps ``p + min(#converged, (max_dim - p) // 2)`` vectors at a restart
        (ARPACK's adjustment); the reference keeps ``p`` (README.md:117 lists this as TODO).
    fast_real_schur : True factors H_m with dgees + 2 x 2 block rotations instead of zgees while
        H_m is real (real basis): a third of the host arithmetic of a restart, which is what the
        GPUs wait for; a real H_m that is symmetric to 1e-12 (a symmetric operator) is
        eigendecomposed with dsyevd instead and ordered by a column permutation.  A valid ordered
        Schur form, but not the reference's bit for bit (signs / phases of the Schur vectors,
        rounding): converged values agree, restart counts may move by rounding noise exactly as
        they do between OpenBLAS builds.
    halo  : "pull" (each rank reads the remote entries of v it needs straight from peer HBM),
        "push" (the owner gathers locally and streams them into the peer's buffer) or "auto"
        (push when some rank's halo has more than 65 536 scattered entries)
    spmv_algo : "auto" | "stream" | "vector" | "merge" (``ab200_set_csr``)

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
    if real_arith == "auto":
        real_arith = "lossless"
    assert real_arith in ("lossless", "pairs", "off"), real_arith
    assert on_breakdown in ("raise", "continue"), on_breakdown
    if not real_storage:
        real_arith = "off"
    if max_dim > MAX_DIM_LIMIT:
        raise ValueError(
            f"max_dim = {max_dim} exceeds this build's limit of {MAX_DIM_LIMIT} basis columns "
            "(the reference has no limit; see INTEGRATION.md section 3)")

    clock = time.perf_counter
    phases = {}
    t_mark = clock()

    def lap(name):
        nonlocal t_mark
        now = clock()
        phases[name] = phases.get(name, 0.0) + (now - t_mark)
        t_mark = now

    # v0 = randn(n) costs as much as the CSR upload at the benchmark sizes: draw it on a
    # second thread while the device is being set up (same global-RNG call as the reference)
    v0_box = {}
    v0_thread = None
    if v0 is None:
        def _draw():
            v0_box["v0"] = rand_normalized_vector(n, np.complex128)
        v0_thread = threading.Thread(target=_draw)
        v0_thread.start()

    H = np.zeros((max_dim + 1, max_dim), dtype=np.complex128)
    history = History.from_k(nev)
    converged = False
    multi = comm is not None and comm.world > 1
    wrappers = []
    device_op = None
    if multi:
        from .distributed import RowBlock, RowPartition, build_halo_plan, slice_rows
        part = RowPartition(n, comm.world)
        r0, r1 = part.rows(comm.rank)
        if isinstance(A, RowBlock):
            block = A
        else:
            _, wrappers = unwrap(A)
            block = slice_rows(_as_scipy_csr(A), r0, r1)
        assert block.row0 == r0 and block.nrows == r1 - r0, "RowBlock does not match the partition"
        plan = build_halo_plan(block)
        a_is_complex = np.iscomplexobj(plan.data)
        solver_args = dict(device=device, row0=r0, nrows_local=r1 - r0)
    elif is_device_operator(A):
        device_op = A
        r0, r1 = 0, n
        a_is_complex = np.dtype(A.dtype).kind == "c"
        solver_args = dict(device=device)
    else:
        _, wrappers = unwrap(A)
        indptr, indices, data, _ = as_csr(A)
        r0, r1 = 0, n
        a_is_complex = np.iscomplexobj(data)
        solver_args = dict(device=device)

    lap("host_prepare")
    dev = DeviceSolver(n, max_dim, **solver_args)
    try:
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
            dev.set_csr(plan.indptr, plan.indices, plan.data, algo=spmv_algo)
        elif device_op is not None:
            dev.set_operator(device_op)
        else:
            dev.set_csr(indptr, indices, data, algo=spmv_algo)
        lap("upload_csr")
        if v0_thread is not None:
            v0_thread.join()
            v0 = v0_box["v0"]
        v0 = np.asarray(v0)
        lap("host_v0")

        # Real or complex storage is ONE decision for the whole solve, not one per rank: a rank
        # whose slice of v0 happens to be real must not stay in float64 while its peers switch
        basis_real = (real_arith != "off" and not a_is_complex
                      and not (np.iscomplexobj(v0) and np.any(v0.imag != 0)))
        if multi:
            votes = comm.all_gather_bytes(bytes([1 if basis_real else 0]))
            basis_real = all(v == b"\x01" for v in votes)
            # every rank drew the n-vector from its own copy of NumPy's global RNG: make sure
            # they really are the same vector (same seed on every rank)
            probe = np.array([v0[0], v0[n // 2], v0[-1], np.vdot(v0, v0)], np.complex128).tobytes()
            assert all(b == probe for b in comm.all_gather_bytes(probe)), \
                "ranks disagree on v0: seed NumPy's global RNG identically on every rank"
        if not basis_real:
            dev.set_option("real_mode", 0)
        dev.set_columns(0, v0[r0:r1])
        # (multi-GPU: ab200_expand opens with a device-side peer barrier, so every rank's
        #  columns are in place before any halo read -- no host barrier needed here)
        lap("upload_v0")

        injected = 0

        def grow_once(start):
            cols, n_iter, _ = dev.expand(start, max_dim, tol, ortho=kind)
            for j in range(start, n_iter):
                rows = j + 2 if (j + 1 < n_iter or n_iter == max_dim) else j + 1
                H[:rows, j] = cols[:rows, j]
            return n_iter

        def grow(start):
            """Expand towards max_dim.  Returns (m, exhausted).  A breakdown at column
            m < max_dim is left to the caller when on_breakdown="raise"; with "continue" it is
            repaired by appending a fresh direction until max_dim is reached or the basis
            spans the whole space (exhausted)."""
            nonlocal injected
            m = grow_once(start)
            while m != max_dim and on_breakdown == "continue":
                if m >= n:
                    return m, True
                # V[:, :m] spans an invariant subspace: A V_m = V_m H_m.  Decouple it and go on.
                fresh = rand_normalized_vector(n, np.complex128)
                dev.set_columns(m, fresh[r0:r1])
                beta = dev.orthonormalize_column(m, m, tol, ortho=kind)
                if beta < tol:      # nothing left outside the basis
                    return m, True
                injected += 1
                H[m:, m - 1] = 0
                m = grow_once(m)
            return m, False

        m, exhausted = grow(0)
        lap("expand")
        locked = 0
        p_grown = 0
        restart = -1
        Tout = np.array(H[:nev, :nev])
        pending = None      # (W, pe): kept block still expressed in a real basis
        for restart in range(max_restarts):
            if m != max_dim and not exhausted:
                raise ValueError("Happy breakdown not supported yet")   # krylov_schur.py:57-59
            reported = restart * (max_dim - nev) + (m - nev)  # krylov_schur.py:63

            # rotate: zgees on H_m, then the reference's ordered_schur on the triangular result
            # (its second zgees is an exact no-op on triangular input and is skipped, its
            # ztrexc sequence is kept call for call so Q matches to rounding)
            T2, Q = rotate(H[:m, :m], sort_function, fast_real=fast_real_schur and basis_real)
            last_beta = 0.0 if exhausted else H[m, m - 1]

            # convergence estimates |beta q_{m-1,k}| / |t_kk|     (krylov_schur.py:91-92)
            estimate = np.abs(last_beta * Q[m - 1, :]) / np.abs(np.diag(T2))
            hit = estimate[:nev] <= tol
            history.matvecs[hit] = reported
            history.restarts[hit] = restart + 1
            converged = bool(np.all(estimate[:nev] < tol)) or exhausted
            Tout = np.array(T2[:nev, :nev])
            lap("host_schur")

            # ---- truncate                                         (krylov_schur.py:72-88)
            pe = min(p, m - 1)
            if dynamic_p and not converged:
                nconv = int(np.count_nonzero(estimate[:nev] < tol))
                pe = min(p + min(nconv, (max_dim - p) // 2), m - 1)
            Z = None
            if basis_real and not converged:
                if not np.any(Q[:, :pe].imag):
                    Z = np.ascontiguousarray(Q[:, :pe].real)     # the reference's own basis
                elif real_arith == "pairs":
                    Z = real_invariant_basis(Q[:, :pe])
                    if Z is None:
                        alt = pe + 1 if pe + 1 < m else pe - 1
                        Z = real_invariant_basis(Q[:, :alt]) if alt >= nev else None
                        if Z is not None:
                            pe = alt
                            p_grown += 1
            pending = None
            if converged:
                # the reference's own last step: complex Schur vectors into the first columns
                if m > pe:
                    dev.restart(Q, m, pe)
                else:
                    dev.combine(Q[:, :nev], 0, m, nev)
                basis_real = False
                lap("restart")
                break
            if Z is not None and not np.any(Q[:, :pe].imag):
                B = Q[:, :pe]
                block = T2[:pe, :pe]
            elif Z is not None:
                B = Z.astype(np.complex128)
                block = (Z.T @ H[:m, :m].real @ Z).astype(np.complex128)
                pending = (Z.T @ Q[:, :pe], pe)
            else:
                B = Q[:, :pe]
                block = T2[:pe, :pe]
                basis_real = False
            spike = H[m, :m] @ B
            if lock and pending is None:
                while locked < nev and estimate[locked] < tol:
                    locked += 1
                spike[:locked] = 0
            dev.restart(B, m, pe)
            lap("restart")
            H[:pe, :pe] = block
            H[pe, :pe] = spike
            H[pe, pe:] = 0
            H[pe + 1:, :] = 0
            m, exhausted = grow(pe)
            lap("expand")

        if pending is not None:
            # not converged, and the kept block is expressed in a real basis Z: rotate it to the
            # Schur vectors the complex iteration would hold (Q_p = Z W), as the caller expects
            W, pe = pending
            dev.combine(W[:, :nev], 0, pe, nev)
            basis_real = False

        if stats is not None:
            stats.update(dev.stats())
            stats["true_matvecs"] = stats["arnoldi_steps"]
            stats["restart_cycles"] = restart + 1
            stats["converged"] = converged
            stats["real_arith"] = real_arith
            stats["pairs_kept_whole"] = p_grown
            stats["breakdowns_repaired"] = injected
            stats["locked"] = locked
        credit_matvecs(wrappers, dev.true_matvecs())
        if not converged and raise_on_no_convergence:
            raise ValueError("Has not converged !")
        lap("host_schur")
        Qout = dev.get_columns(0, nev, hugepages=True)
        lap("download_q")
    finally:
        # two-phase teardown: unmap the peers' buffers, meet, and only then free what the
        # peers had mapped (cudaFree of an exported allocation before the importer closed
        # its handle is undefined)
        if multi:
            try:
                dev.disconnect()
                comm.barrier()
            finally:
                dev.close()
        else:
            dev.close()
    lap("device_free")
    if stats is not None:
        stats["host_phases_s"] = phases

    return Qout, Tout, history
