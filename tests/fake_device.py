"""Test double for ``arnoldi_b200.solver.DeviceSolver`` (TEST INFRASTRUCTURE ONLY).

Lets the CPU suite execute the host driver (krylov_schur.py / decomposition.py: defaults,
H bookkeeping, spike row, convergence test, history, exceptions) without a GPU by answering the
device calls with the oracle's NumPy restatement.  It is injected with ``monkeypatch`` inside
tests only; the product never imports it and has no such fallback.
"""
import numpy as np
import scipy.sparse as sp

import oracle


class FakeDeviceSolver:
    def __init__(self, n, max_dim, *, device=0, row0=0, nrows_local=None):
        assert row0 == 0 and nrows_local in (None, n), "the test double is single-rank"
        self.n, self.max_dim = int(n), int(max_dim)
        self.V = np.zeros((self.n, self.max_dim + 1), np.complex128, order="F")
        self.H = np.zeros((self.max_dim + 1, self.max_dim), np.complex128)
        self.counters = {}
        self.options = {}

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        pass

    def set_option(self, key, value):
        self.options[key] = value

    def set_timing(self, on=True):
        pass

    def set_csr(self, indptr, indices, data, **kw):
        self.A = sp.csr_matrix((data, indices, indptr), shape=(self.n, self.n))

    def set_columns(self, col0, cols):
        cols = np.asarray(cols, dtype=np.complex128)
        if cols.ndim == 1:
            cols = cols.reshape(-1, 1)
        self.V[:, col0:col0 + cols.shape[1]] = cols

    def get_columns(self, col0, ncols, out=None, hugepages=False):
        return np.array(self.V[:, col0:col0 + ncols], order="F")

    def expand(self, start_dim, end_dim, tol, *, eta=np.sqrt(0.5), ortho=0):
        fn = oracle.cgs_dgks if ortho == 0 else oracle.mgs_dgks
        _, _, k = oracle.arnoldi_expand(self.A, self.V, self.H, tol, start_dim=start_dim,
                                        max_dim=end_dim, ortho=fn, counters=self.counters)
        return np.asfortranarray(self.H), k, k != end_dim

    def restart(self, Q, m, p):
        oracle.restart_update(self.V, np.asarray(Q), m, p)

    def combine(self, Q, col0, m, p):
        Q = np.asarray(Q)[:m, :p]
        self.V[:, col0:col0 + p] = self.V[:, col0:col0 + m] @ Q

    def orthonormalize_column(self, col, ncols, tol, *, eta=np.sqrt(0.5), ortho=0):
        fn = oracle.cgs_dgks if ortho == 0 else oracle.mgs_dgks
        w = self.V[:, col]
        h = np.zeros(max(ncols, 1), np.complex128)
        if ncols and ortho == 1 and eta == 0.0:
            # a single modified Gram-Schmidt sweep (explicit_restarts.py:63-77)
            for j in range(ncols):
                w -= np.vdot(self.V[:, j], w) * self.V[:, j]
            beta = float(np.linalg.norm(w))
        elif ncols:
            beta, _ = fn(w, self.V[:, :ncols], h, tol, eta)
        else:
            beta = float(np.linalg.norm(w))
        if beta >= tol:
            w /= beta
        return float(beta)

    def project(self, col, nrows):
        return self.V[:, :nrows].conj().T @ (self.A @ self.V[:, col])

    def disconnect(self):
        pass

    def true_matvecs(self):
        return self.counters.get("matvecs", 0)

    def stats(self):
        c = self.counters
        return {"arnoldi_steps": c.get("matvecs", 0), "ortho_rounds": c.get("rounds", 0),
                "second_rounds": c.get("rounds", 0) - c.get("calls", 0), "kernel_launches": 0,
                "real_storage": 0}
