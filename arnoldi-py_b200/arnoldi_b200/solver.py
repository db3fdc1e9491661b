"""Thin object wrapper over one ``ab200_solver`` handle (one GPU, one row block)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class DeviceSolver:
    """Owns the device state of one Krylov-Schur solve: basis V, CSR block of A, H copy.

    Replaces the allocations at krylov_schur.py:42-43 and the three n-length call
    sites decomposition.py:58,60 and krylov_schur.py:78-81.
    """

    def __init__(self, n, max_dim, *, device=0, row0=0, nrows_local=None):
        self.lib = _lib.load()
        self.n_global = int(n)
        self.n = int(n if nrows_local is None else nrows_local)
        self.row0 = int(row0)
        self.max_dim = int(max_dim)
        self._h = C.c_void_p()
        _lib.check(self.lib.ab200_create(C.byref(self._h), int(device), self.n_global, self.row0,
                                         self.n, self.max_dim))
        # column-major staging area for the H columns an expansion returns
        self._hbuf = np.zeros((self.max_dim + 1, self.max_dim), np.complex128, order="F")

    # -- lifetime ---------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.ab200_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- operator ---------------------------------------------------------------
    def set_csr(self, indptr, indices, data, *, algo=_lib.SPMV_AUTO):
        """Upload a CSR row block exactly as scipy stores it (no re-ordering)."""
        if isinstance(algo, str):
            assert algo in _lib.SPMV_ALGOS, f"spmv_algo must be one of {sorted(_lib.SPMV_ALGOS)}"
            algo = _lib.SPMV_ALGOS[algo]
        indptr = np.ascontiguousarray(indptr)
        if indptr.dtype not in (np.int32, np.int64):
            indptr = indptr.astype(np.int64)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        if np.iscomplexobj(data):
            data = np.ascontiguousarray(data, dtype=np.complex128)
            kind = _lib.C128
        else:
            data = np.ascontiguousarray(data, dtype=np.float64)
            kind = _lib.F64
        assert indptr.shape == (self.n + 1,), "indptr must have nrows_local + 1 entries"
        assert indices.shape == data.shape, "indices and data differ in length"
        _lib.check(self.lib.ab200_set_csr(self._h, _ptr(indptr), indptr.dtype.itemsize * 8,
                                          _ptr(indices), _ptr(data), kind, int(data.shape[0]),
                                          int(algo)))
        self.nnz = int(data.shape[0])
        self.value_kind = kind

    def set_operator(self, op):
        """Use a device operator (``device_operator.DeviceOperator`` protocol) instead of CSR:
        ``op.device_apply(x_ptr, y_ptr, n, is_real, stream_ptr)`` must enqueue ``y = A x`` on
        the given CUDA stream.  The callback object is kept alive by this solver."""
        def trampoline(_user, x, y, n, is_real, stream):
            try:
                op.device_apply(int(x), int(y), int(n), bool(is_real), int(stream or 0))
                return 0
            except Exception:      # an exception cannot cross the C frame
                import traceback
                traceback.print_exc()
                return 1
        self._op_cb = _lib.APPLY_FN(trampoline)
        kind = _lib.C128 if np.dtype(op.dtype).kind == "c" else _lib.F64
        _lib.check(self.lib.ab200_set_operator(self._h, C.cast(self._op_cb, C.c_void_p), None, kind))
        self.nnz = 0
        self.value_kind = kind

    # -- basis ------------------------------------------------------------------
    def set_columns(self, col0, cols):
        cols = np.asarray(cols, dtype=np.complex128)
        if cols.ndim == 1:
            cols = cols.reshape(-1, 1)
        cols = np.asfortranarray(cols)
        assert cols.shape[0] == self.n
        _lib.check(self.lib.ab200_set_columns(self._h, int(col0), int(cols.shape[1]), _ptr(cols),
                                              int(cols.shape[0])))

    @staticmethod
    def _hugepage_hint(arr):
        """Ask for transparent huge pages on a freshly allocated result array: the device-to-
        host copy of a multi-GB Q is otherwise dominated by one page fault per 4 KiB."""
        try:
            libc = C.CDLL(None, use_errno=True)
            addr = arr.ctypes.data
            start = (addr + 0x1FFFFF) & ~0x1FFFFF
            length = (addr + arr.nbytes - start) & ~0x1FFFFF
            if length > 0:
                libc.madvise(C.c_void_p(start), C.c_size_t(length), 14)   # MADV_HUGEPAGE
        except Exception:
            pass

    def get_columns(self, col0, ncols, out=None, hugepages=False):
        if out is None:
            out = np.empty((self.n, ncols), np.complex128, order="F")
            if hugepages and out.nbytes >= (64 << 20):
                self._hugepage_hint(out)
        assert out.flags.f_contiguous and out.shape == (self.n, ncols)
        _lib.check(self.lib.ab200_get_columns(self._h, int(col0), int(ncols), _ptr(out), self.n))
        return out

    # -- hot path ---------------------------------------------------------------
    def expand(self, start_dim, end_dim, tol, *, eta=np.sqrt(0.5), ortho=_lib.ORTHO_CGS2):
        """decomposition.py:56-66 on the device.  Returns (H columns view, n_iter, breakdown).

        The returned array is column-major (max_dim+1) x max_dim; only columns
        [start_dim, n_iter) are meaningful (rows 0..j+1 of column j).
        """
        n_iter, brk = C.c_int(0), C.c_int(0)
        _lib.check(self.lib.ab200_expand(self._h, int(start_dim), int(end_dim), float(tol),
                                         float(eta), int(ortho), _ptr(self._hbuf),
                                         C.byref(n_iter), C.byref(brk)))
        return self._hbuf, n_iter.value, bool(brk.value)

    def restart(self, Q, m, p):
        """krylov_schur.py:78,81:  V[:, :p] = V[:, :m] Q[:, :p];  V[:, p] = V[:, m]."""
        q = np.asfortranarray(Q[:m, :p], dtype=np.complex128)
        _lib.check(self.lib.ab200_restart(self._h, _ptr(q), int(m), int(m), int(p)))

    def combine(self, Q, col0, m, p):
        """V[:, col0:col0+p] = V[:, col0:col0+m] Q[:m, :p]; no other column is touched."""
        q = np.asfortranarray(np.asarray(Q)[:m, :p], dtype=np.complex128)
        _lib.check(self.lib.ab200_combine(self._h, _ptr(q), int(m), int(col0), int(m), int(p)))

    def orthonormalize_column(self, col, ncols, tol, *, eta=np.sqrt(0.5), ortho=_lib.ORTHO_CGS2):
        """Orthonormalise basis column ``col`` against columns [0, ncols) on the device;
        returns its norm after the projections (below ``tol``: left un-normalised)."""
        beta, brk = C.c_double(0.0), C.c_int(0)
        _lib.check(self.lib.ab200_orthonormalize_column(self._h, int(col), int(ncols), float(tol),
                                                        float(eta), int(ortho), C.byref(beta),
                                                        C.byref(brk)))
        return beta.value

    def project(self, col, nrows):
        """h[i] = <V_i, A V_col> for i < nrows (explicit_restarts.py:150-151)."""
        h = np.empty(nrows, np.complex128)
        _lib.check(self.lib.ab200_project(self._h, int(col), int(nrows), _ptr(h)))
        return h

    def spmv(self, x):
        x = np.ascontiguousarray(x, dtype=np.complex128)
        assert x.shape == (self.n_global,)
        y = np.empty(self.n, np.complex128)
        _lib.check(self.lib.ab200_spmv(self._h, _ptr(x), _ptr(y)))
        return y

    def ortho(self, ncols, w, h, tol, eta, kind):
        """In place on contiguous complex128 ``w`` (n) and ``h`` (ncols)."""
        beta, brk = C.c_double(0.0), C.c_int(0)
        _lib.check(self.lib.ab200_ortho(self._h, int(ncols), _ptr(w), _ptr(h), float(tol),
                                        float(eta), int(kind), C.byref(beta), C.byref(brk)))
        return beta.value, bool(brk.value)

    # -- multi-GPU ---------------------------------------------------------------
    def connect(self, comm, partition):
        """Exchange CUDA-IPC handles with the other ranks (``comm`` = TorchComm-like) and map
        their basis / reduction buffers.  Collective: every rank must call it."""
        blob = C.create_string_buffer(256)
        _lib.check(self.lib.ab200_comm_export(self._h, blob))
        blobs = b"".join(comm.all_gather_bytes(blob.raw))
        starts = np.ascontiguousarray(partition.starts, dtype=np.int64)
        _lib.check(self.lib.ab200_comm_connect(self._h, int(comm.rank), int(comm.world), blobs,
                                               _ptr(starts)))
        comm.barrier()

    def disconnect(self):
        """First half of the multi-GPU teardown (unmap the peers); follow with a barrier."""
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.check(self.lib.ab200_comm_disconnect(self._h))

    def comm_bench(self, iters=1000):
        """Microseconds per in-kernel peer reduction (1-block exchange kernel, back to back)."""
        us = C.c_double(0.0)
        _lib.check(self.lib.ab200_comm_bench(self._h, int(iters), C.byref(us)))
        return us.value

    def set_halo(self, ghost_cols):
        g = np.ascontiguousarray(ghost_cols, dtype=np.int64)
        _lib.check(self.lib.ab200_set_halo(self._h, _ptr(g), int(g.shape[0])))

    def connect_halo_push(self, comm, partition, ghost_cols):
        """Switch the halo exchange to owner-side push (scattered halos).  Collective, after
        ``set_halo`` on every rank: ranks swap their ghost lists, each works out which of its
        rows every peer reads and where they land in that peer's ghost buffer."""
        blob = C.create_string_buffer(256)
        _lib.check(self.lib.ab200_halo_export(self._h, blob))
        blobs = b"".join(comm.all_gather_bytes(blob.raw))
        from .distributed import plan_halo_push
        lists = comm.all_gather_bytes(np.ascontiguousarray(ghost_cols, dtype=np.int64).tobytes())
        send_idx, send_ptr, dst_off = plan_halo_push(
            [np.frombuffer(b, dtype=np.int64) for b in lists], partition, comm.rank)
        if send_idx.shape[0] == 0:
            send_idx = np.zeros(1, np.int64)
        _lib.check(self.lib.ab200_halo_connect(self._h, blobs, _ptr(send_idx), _ptr(send_ptr),
                                               _ptr(dst_off)))
        comm.barrier()

    # -- measurement ------------------------------------------------------------
    def set_timing(self, on=True):
        _lib.check(self.lib.ab200_set_timing(self._h, int(bool(on))))

    def reset_stats(self):
        _lib.check(self.lib.ab200_reset_stats(self._h))

    def stats(self):
        st = _lib.Stats()
        _lib.check(self.lib.ab200_get_stats(self._h, C.byref(st)))
        return st.as_dict()

    def true_matvecs(self):
        """Operator applications since the last reset (device-side counter)."""
        return int(self.stats()["arnoldi_steps"])

    def synchronize(self):
        _lib.check(self.lib.ab200_synchronize(self._h))

    def timer_start(self):
        _lib.check(self.lib.ab200_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_double(0.0)
        _lib.check(self.lib.ab200_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def set_option(self, key, value):
        _lib.check(self.lib.ab200_set_option(self._h, key.encode(), int(value)))
