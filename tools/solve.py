#!/usr/bin/env python
"""Run partial_schur to convergence on one of the BASELINE configurations and report
time-to-k-converged, restart / matvec counts, true residuals and per-kernel figures.

    python tools/solve.py --matrix lap2d --grid 4096 --nev 10 --max-dim 40
    python tools/solve.py --matrix mark --grid 4000 --nev 20 --max-dim 60
    (tests/solve_vs_oracle.py runs the same solve with the CPU oracle beside it)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "arnoldi-py_b200")):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--matrix", default="lap2d")
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--nev", type=int, default=10)
    ap.add_argument("--max-dim", type=int, default=40)
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--max-restarts", type=int, default=100000)
    ap.add_argument("--ortho", default="cgs2")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--real-arith", default="lossless", choices=["lossless", "pairs", "off"])
    ap.add_argument("--fast-real-schur", action="store_true")
    ap.add_argument("--dynamic-p", action="store_true")
    ap.add_argument("--lock", action="store_true")
    args = ap.parse_args()

    from arnoldi_b200 import matrices, partial_schur
    from arnoldi_b200.utils import arg_largest_real
    A = getattr(matrices, args.matrix)(args.grid)
    n = A.shape[0]
    np.random.seed(args.seed)
    stats = {}
    t0 = time.perf_counter()
    Q, T, hist = partial_schur(A, args.nev, max_dim=args.max_dim, stopping_criterion=args.tol,
                               max_restarts=args.max_restarts, sort_function=arg_largest_real,
                               ortho=args.ortho, stats=stats, real_arith=args.real_arith,
                               fast_real_schur=args.fast_real_schur, dynamic_p=args.dynamic_p,
                               lock=args.lock)
    dt = time.perf_counter() - t0
    w, S = np.linalg.eig(T)
    X = Q @ S
    res = np.linalg.norm(A @ X - X * w, axis=0) / np.abs(w)
    out = {
        "config": f"{args.matrix}({args.grid}) n={n} nnz={A.nnz} K={args.nev} max_dim={args.max_dim} "
                  f"LR tol={args.tol} seed={args.seed} ortho={args.ortho} real_arith={args.real_arith} "
                  f"fast_real_schur={args.fast_real_schur} dynamic_p={args.dynamic_p} lock={args.lock}",
        "pairs_kept_whole": int(stats.get("pairs_kept_whole", 0)),
        "real_storage_at_end": int(stats.get("real_storage", 0)),
        "time_to_k_converged_s": dt, "restarts": int(hist.restarts[0]),
        "history_matvecs": int(hist.matvecs[0]), "true_matvecs": int(stats["true_matvecs"]),
        "matvecs_per_s": stats["true_matvecs"] / dt,
        "dgks_second_round_fraction": stats["second_rounds"] / max(1, stats["arnoldi_steps"]),
        "max_true_residual": float(res.max()), "ritz_real": np.diag(T).real.tolist(),
        "orthogonality": float(np.abs(Q.conj().T @ Q - np.eye(args.nev)).max()),
        "host_phases_s": {k: round(v, 3) for k, v in stats["host_phases_s"].items()},
        "kernels": {k: {"ms": round(stats[k + "_ms"], 2), "launches": stats[k + "_launches"],
                        "gbs": round(stats[k + "_bytes"] / max(stats[k + "_ms"], 1e-9) / 1e6, 1)}
                    for k in ("spmv", "ortho_pass1", "ortho_fused", "ortho_pass2", "mgs", "restart")
                    if stats[k + "_launches"]},
    }
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
