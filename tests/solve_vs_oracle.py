#!/usr/bin/env python
"""Checker script (test infrastructure): one solve on the GPU path and the same seeded solve on
the CPU oracle, printed side by side.  Lives under tests/ because only tests may use oracle/.

    python tests/solve_vs_oracle.py --matrix mark --grid 400 --nev 20 --max-dim 60
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
    args = ap.parse_args()
    import oracle
    from arnoldi_b200 import matrices, partial_schur
    from arnoldi_b200.utils import arg_largest_real
    A = getattr(matrices, args.matrix)(args.grid)
    kw = dict(max_dim=args.max_dim, stopping_criterion=args.tol, max_restarts=args.max_restarts)
    np.random.seed(args.seed)
    stats = {}
    t0 = time.perf_counter()
    Q, T, hist = partial_schur(A, args.nev, sort_function=arg_largest_real, ortho=args.ortho,
                               stats=stats, **kw)
    dt = time.perf_counter() - t0
    np.random.seed(args.seed)
    cnt = {}
    t0 = time.perf_counter()
    Qo, To, ho = oracle.partial_schur(
        A, args.nev, sort_function=oracle.arg_largest_real, counters=cnt,
        ortho=oracle.cgs_dgks if args.ortho == "cgs2" else oracle.mgs_dgks, **kw)
    dto = time.perf_counter() - t0
    rel = np.abs(np.diag(T) - np.diag(To)) / np.abs(np.diag(To))
    print(json.dumps({
        "config": f"{args.matrix}({args.grid}) n={A.shape[0]} K={args.nev} max_dim={args.max_dim}",
        "gpu": {"time_s": dt, "restarts": int(hist.restarts[0]),
                "true_matvecs": int(stats["true_matvecs"])},
        "oracle": {"time_s": dto, "restarts": int(ho.restarts[0]), "true_matvecs": cnt["matvecs"],
                   "cpu_count": os.cpu_count()},
        "rel_ritz_diff": rel.tolist()}), flush=True)


if __name__ == "__main__":
    main()
