"""Round-2 additions to tests/golden: records of the UNMODIFIED reference on operators the
round-1 set did not cover.  TEST INFRASTRUCTURE ONLY; run in the build container:

    python oracle/make_golden_r2.py          # writes tests/golden/solves_r2.npz

* ``rect{32,64}``: anisotropic 5-point Laplacian on an N x (N+1) grid (weights 1, 0.75).  Same
  family as BASELINE config 2 (symmetric, real, <= 5 entries per row) but with SIMPLE
  eigenvalues, so the restart count and every Ritz value are well defined and can be asserted
  tightly at any rank count (the square isotropic grid has double eigenvalues and its counts are
  decided by rounding noise).
* ``mark200_s0``: the reference at the next size of the config-3 family (R = 55).

Inputs (CSR arrays) are stored next to the outputs, so the tests never rebuild a matrix.
"""
from __future__ import annotations

import os

import numpy as np
import scipy.sparse as sp

from make_golden import OUT, _import_reference, csr_parts, solve_record


def lap2d_rect(nx, ny, wx=1.0, wy=0.75):
    T = lambda N: sp.diags_array([-np.ones(N - 1), 2 * np.ones(N), -np.ones(N - 1)],  # noqa: E731
                                 offsets=[-1, 0, 1])
    A = (wy * sp.kron(T(ny), sp.eye_array(nx)) + wx * sp.kron(sp.eye_array(ny), T(nx))).tocsr()
    A.sum_duplicates()
    A.eliminate_zeros()
    A.sort_indices()
    return A


def main():
    dec, ks, mats, ortho, utils = _import_reference()
    sol = {}

    def put(tag, rec):
        for k, v in rec.items():
            sol[f"{tag}_{k}"] = v

    for N in (32, 64):
        A = lap2d_rect(N, N + 1)
        for k, v in csr_parts(A).items():
            sol[f"rect{N}_{k}"] = v
        for seed in (0, 1):
            put(f"rect{N}_s{seed}", solve_record(ks, utils, A, seed, nev=10, max_dim=40,
                                                 stopping_criterion=1e-8, max_restarts=2000))
    put("mark200_s0", solve_record(ks, utils, mats.mark(200), 0, nev=20, max_dim=60,
                                   stopping_criterion=1e-8, max_restarts=2000))
    # ---- explicit restarts with deflation (explicit_restarts.py:80-168) -------------
    import arnoldi.explicit_restarts as er
    ex = {}
    for tag, A, kw in (("mark10", mats.mark(10), dict(nev=3, max_dim=10, stopping_criterion=1e-8)),
                       ("mark20", mats.mark(20), dict(nev=4, max_dim=20, stopping_criterion=1e-8,
                                                      max_restarts=400)),
                       ("rect12", lap2d_rect(12, 13), dict(nev=4, max_dim=24, stopping_criterion=1e-9,
                                                           max_restarts=400))):
        np.random.seed(0)
        nev = kw.pop("nev")
        vals, vecs, hist = er.explicit_restarts_with_deflation(
            A, nev, sort_function=utils.arg_largest_real, **kw)
        for k, v in csr_parts(A).items():
            ex[f"{tag}_{k}"] = v
        ex[f"{tag}_vals"] = vals
        ex[f"{tag}_vecs"] = vecs
        ex[f"{tag}_hist_matvecs"] = hist.matvecs
        ex[f"{tag}_hist_restarts"] = hist.restarts
        ex[f"{tag}_res"] = np.linalg.norm(A @ vecs - vals * vecs, axis=0)
        print("explicit", tag, "vals", vals, "restarts", hist.restarts, "res %.2e" % ex[f"{tag}_res"].max())
    np.savez_compressed(os.path.join(OUT, "explicit.npz"), **ex)
    np.savez_compressed(os.path.join(OUT, "solves_r2.npz"), **sol)
    for tag in ("rect32_s0", "rect32_s1", "rect64_s0", "rect64_s1", "mark200_s0"):
        print(tag, "R=", sol[f"{tag}_hist_restarts"][:1], "true=", sol[f"{tag}_true_matvecs"],
              "maxres=%.2e" % sol[f"{tag}_eig_res"].max(), "diagT", sol[f"{tag}_diagT"][:3])


if __name__ == "__main__":
    main()
