"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container, where
``/root/reference`` is mounted:

    python oracle/make_golden.py

Nothing here is imported at test or bench time; the GPU box has no
``/root/reference`` and only reads the committed ``.npz`` files.  Every array
below is an output of the reference's own functions (``arnoldi.*``) on seeded
inputs; the inputs are stored next to the outputs so the tests never have to
re-create a random stream.
"""

from __future__ import annotations

import os
import sys

import numpy as np
import scipy
import scipy.sparse as sp

REF_SRC = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _import_reference():
    sys.path.insert(0, REF_SRC)
    import arnoldi  # noqa: F401
    import arnoldi.decomposition as dec
    import arnoldi.krylov_schur as ks
    import arnoldi.matrices as mats
    import arnoldi.ortho as ortho
    import arnoldi.utils as utils
    return dec, ks, mats, ortho, utils


class CountingOperator:
    """Duck-typed operator (shape, dtype, @) that counts applications."""

    def __init__(self, A):
        self.A = A
        self.shape = A.shape
        self.dtype = A.dtype
        self.count = 0

    def __matmul__(self, x):
        self.count += 1
        return self.A @ x


def lap2d(N):
    T = sp.diags_array([-np.ones(N - 1), 2 * np.ones(N), -np.ones(N - 1)],
                       offsets=[-1, 0, 1])
    I = sp.eye_array(N)
    return (sp.kron(I, T) + sp.kron(T, I)).tocsr()


def csr_parts(A):
    A = A.tocsr()
    return dict(indptr=A.indptr.astype(np.int64), indices=A.indices.astype(np.int32),
                data=A.data, shape=np.array(A.shape, np.int64))


def solve_record(ks, utils, A, seed, **kw):
    op = CountingOperator(A)
    np.random.seed(seed)
    Q, T, hist = ks.partial_schur(op, kw.pop("nev"), sort_function=utils.arg_largest_real, **kw)
    Q = np.array(Q)
    T = np.array(T)
    lam = np.diag(T)
    # true residuals of the Schur relation and of the eigenpairs
    schur_res = np.linalg.norm(A @ Q - Q @ T, axis=0)
    w, S = np.linalg.eig(T)
    X = Q @ S
    eig_res = np.linalg.norm(A @ X - X * w, axis=0) / np.abs(w)
    return dict(diagT=lam, T=T, hist_matvecs=hist.matvecs, hist_restarts=hist.restarts,
                true_matvecs=np.int64(op.count), schur_res=schur_res, eig_vals=w,
                eig_res=eig_res, orth_err=np.abs(Q.conj().T @ Q - np.eye(Q.shape[1])).max())


def main():
    os.makedirs(OUT, exist_ok=True)
    dec, ks, mats, ortho, utils = _import_reference()
    meta = dict(numpy=np.__version__, scipy=scipy.__version__)

    # ---- matrices: golden CSR content (tests/test_matrices.py:8-59) ----------
    mat = {}
    for m in (2, 3, 10, 17, 20, 50):
        for k, v in csr_parts(mats.mark(m)).items():
            mat[f"mark{m}_{k}"] = v
    for k, v in csr_parts(mats.laplace(5)).items():
        mat[f"laplace5_{k}"] = v
    mat["laplace_eigen5"] = mats.laplace_eigen(5)
    np.savez_compressed(os.path.join(OUT, "matrices.npz"), **mat)

    # ---- orthogonalisation in isolation (ortho.py) ----------------------------
    rng = np.random.default_rng(1234)
    orth = {}
    cases = []
    for name, n, c in (("small", 64, 5), ("mid", 300, 12), ("tall", 1200, 33), ("one", 100, 1)):
        M = rng.standard_normal((n, c)) + 1j * rng.standard_normal((n, c))
        V, _ = np.linalg.qr(M)
        V = np.asfortranarray(V)
        w = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        cases.append((name, V, w))
        # nearly dependent: triggers the second DGKS round
        w2 = V @ (rng.standard_normal(c) + 1j * rng.standard_normal(c)) \
            + 1e-3 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
        cases.append((name + "_dgks", V, w2))
        # in span: breakdown
        w3 = V @ (rng.standard_normal(c) + 1j * rng.standard_normal(c))
        cases.append((name + "_brk", V, w3))
    names = []
    for name, V, w in cases:
        names.append(name)
        orth[f"{name.split('_')[0]}_V"] = V  # one copy per basis, shared by its 3 cases
        orth[f"{name}_w"] = w
        for tag, fn in (("gs", ortho.dgks_gs), ("mgs", ortho.dgks_mgs)):
            ww = w.copy()
            h = np.zeros(V.shape[1], np.complex128)
            beta, brk = fn(ww, V, h, 1e-8)
            orth[f"{name}_{tag}_w"] = ww
            orth[f"{name}_{tag}_h"] = h
            orth[f"{name}_{tag}_beta"] = np.float64(beta)
            orth[f"{name}_{tag}_brk"] = np.bool_(brk)
    orth["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "ortho.npz"), **orth)

    # ---- Arnoldi expansion (decomposition.py:13-68) ---------------------------
    arn = {}
    rng = np.random.default_rng(77)
    # (a) complex sparse, like tests/test_decomposition.py:72-90 but seeded
    n, m = 40, 12
    A = sp.random(n, n, density=5 / n, dtype=np.complex128, random_state=rng).tocsr()
    A = (A + sp.diags_array(np.ones(n))).tocsr()
    v0 = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    v0 /= np.linalg.norm(v0)
    V = np.zeros((n, m + 1), np.complex128, order="F")
    H = np.zeros((m + 1, m), np.complex128)
    V[:, 0] = v0
    _, _, n_iter = dec.arnoldi_decomposition(A, V, H, 1e-8)
    for k, v in csr_parts(A).items():
        arn[f"cplx_{k}"] = v
    arn.update(cplx_v0=v0, cplx_V=V, cplx_H=H, cplx_niter=np.int64(n_iter))
    # (b) real mark(10), max_dim < m  (tests/test_decomposition.py:92-113)
    A = mats.mark(10)
    n, m, md = A.shape[0], 10, 6
    v0 = rng.standard_normal(n).astype(np.complex128)
    v0 /= np.linalg.norm(v0)
    V = np.zeros((n, m + 1), np.complex128, order="F")
    H = np.zeros((m + 1, m), np.complex128)
    V[:, 0] = v0
    _, _, n_iter = dec.arnoldi_decomposition(A, V, H, 1e-8, max_dim=md)
    arn.update(mark10_v0=v0, mark10_V=V, mark10_H=H, mark10_niter=np.int64(n_iter),
               mark10_maxdim=np.int64(md))
    # (c) breakdown: v0 an eigenvector (tests/test_decomposition.py:115-139)
    Ad = np.diag(np.arange(1.0, 9.0))
    n, m = 8, 5
    v0 = np.zeros(n, np.complex128)
    v0[2] = 1.0
    V = np.zeros((n, m + 1), np.complex128, order="F")
    H = np.zeros((m + 1, m), np.complex128)
    V[:, 0] = v0
    _, _, n_iter = dec.arnoldi_decomposition(Ad, V, H, 1e-8)
    arn.update(brk_A=Ad, brk_v0=v0, brk_V=V, brk_H=H, brk_niter=np.int64(n_iter))
    np.savez_compressed(os.path.join(OUT, "arnoldi.npz"), **arn)

    # ---- restart update + ordered Schur (krylov_schur.py:69-88, utils.py:32-67)
    rst = {}
    rng = np.random.default_rng(5)
    n, m, p = 700, 20, 10
    M = rng.standard_normal((n, m + 1)) + 1j * rng.standard_normal((n, m + 1))
    V, _ = np.linalg.qr(M)
    V = np.asfortranarray(V)
    Hm = np.triu(rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m)), -1)
    from scipy.linalg import schur
    T1, Q1 = schur(Hm, output="complex")
    T2, Q2 = utils.ordered_schur(T1, output="complex", sort_function=utils.arg_largest_real)
    Q = Q1 @ Q2
    Vout = V.copy(order="F")
    Vout[:, :p] = V[:, :m] @ Q[:, :p]
    Vout[:, p] = V[:, m]
    rst.update(V=V, Hm=Hm, T2=T2, Q2=Q2, Q=Q, Vout=Vout, m=np.int64(m), p=np.int64(p))
    np.savez_compressed(os.path.join(OUT, "restart.npz"), **rst)

    # ---- full solves (krylov_schur.py:10-114) ---------------------------------
    sol = {}

    def put(tag, rec):
        for k, v in rec.items():
            sol[f"{tag}_{k}"] = v

    A50 = mats.mark(50)
    for seed in (0, 1, 42):
        put(f"mark50_s{seed}", solve_record(ks, utils, A50, seed, nev=5, max_dim=20,
                                            stopping_criterion=1e-8, max_restarts=1000))
    put("mark10_s0", solve_record(ks, utils, mats.mark(10), 0, nev=3, max_dim=5,
                                  max_restarts=1000))
    put("mark100_s0", solve_record(ks, utils, mats.mark(100), 0, nev=20, max_dim=60,
                                   stopping_criterion=1e-8, max_restarts=1000))
    put("lap2d32_s0", solve_record(ks, utils, lap2d(32), 0, nev=10, max_dim=40,
                                   stopping_criterion=1e-8, max_restarts=1000))
    put("lap2d64_s0", solve_record(ks, utils, lap2d(64), 0, nev=10, max_dim=40,
                                   stopping_criterion=1e-8, max_restarts=1000))
    # default arguments (largest magnitude, tol = sqrt(eps), max_dim/p defaults)
    np.random.seed(3)
    Q, T, hist = ks.partial_schur(mats.mark(20), 4)
    sol.update(mark20_default_diagT=np.diag(T), mark20_default_restarts=hist.restarts,
               mark20_default_matvecs=hist.matvecs)
    # complex-valued operator
    rng = np.random.default_rng(9)
    n = 400
    Ac = (sp.random(n, n, density=6 / n, dtype=np.complex128, random_state=rng)
          + sp.diags_array(np.linspace(1, 3, n) + 0.5j * np.linspace(-1, 1, n))).tocsr()
    for k, v in csr_parts(Ac).items():
        sol[f"cplx400_{k}"] = v
    put("cplx400_s0", solve_record(ks, utils, Ac, 0, nev=4, max_dim=24,
                                   stopping_criterion=1e-8, max_restarts=2000))

    # MGS plug: the reference hard-wires dgks_gs (decomposition.py:6,60); swap it
    saved = dec.dgks_gs
    dec.dgks_gs = ortho.dgks_mgs
    try:
        put("mark50_mgs_s0", solve_record(ks, utils, A50, 0, nev=5, max_dim=20,
                                          stopping_criterion=1e-8, max_restarts=1000))
        put("lap2d32_mgs_s0", solve_record(ks, utils, lap2d(32), 0, nev=10, max_dim=40,
                                           stopping_criterion=1e-8, max_restarts=1000))
    finally:
        dec.dgks_gs = saved
    sol["meta"] = np.array(repr(meta))
    np.savez_compressed(os.path.join(OUT, "solves.npz"), **sol)

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
    for tag in ("mark50_s0", "mark50_s1", "mark50_s42", "mark10_s0", "mark100_s0",
                "lap2d32_s0", "lap2d64_s0", "cplx400_s0", "mark50_mgs_s0", "lap2d32_mgs_s0"):
        print(tag, "R=", sol[f"{tag}_hist_restarts"], "mv=", sol[f"{tag}_hist_matvecs"][:1],
              "true=", sol[f"{tag}_true_matvecs"], "maxres=%.2e" % sol[f"{tag}_eig_res"].max())


if __name__ == "__main__":
    main()
