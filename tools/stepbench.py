#!/usr/bin/env python
"""BASELINE config 5: Arnoldi-step microbenchmark, SpMV + CGS2 versus SpMV + MGS.

A real Arnoldi expansion is run one step at a time (so the basis is genuinely
orthonormal and the DGKS test behaves as in a solve) and the CUDA-event time and
algorithmic bytes of every kernel class are recorded per basis size c = j + 1.

    python tools/stepbench.py --rows 10000000 --cmax 100 --report 20,40,60,100
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "arnoldi-py_b200")):
    sys.path.insert(0, p)

CLASSES = ("spmv", "ortho_pass1", "ortho_fused", "ortho_pass2", "mgs")


def run(A, cmax, kind, report, variant):
    from arnoldi_b200.solver import DeviceSolver
    from arnoldi_b200.utils import rand_normalized_vector
    n = A.shape[0]
    np.random.seed(0)
    v0 = rand_normalized_vector(n, np.complex128)
    out = []
    with DeviceSolver(n, cmax) as dev:
        dev.set_timing(True)
        dev.set_option("ortho_variant", variant)
        dev.set_csr(A.indptr, A.indices, A.data)
        dev.set_columns(0, v0)
        for j in range(cmax):
            dev.reset_stats()
            _, n_iter, brk = dev.expand(j, j + 1, 1e-8, ortho=kind)
            if brk:
                break
            c = j + 1
            if c in report:
                st = dev.stats()
                rec = {"c": c, "rounds": st["ortho_rounds"]}
                tot = 0.0
                for k in CLASSES:
                    if st[k + "_launches"]:
                        rec[k] = {"ms": round(st[k + "_ms"], 4),
                                  "gbs": round(st[k + "_bytes"] / st[k + "_ms"] / 1e6, 1)}
                        tot += st[k + "_ms"]
                rec["step_ms"] = round(tot, 4)
                out.append(rec)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--matrix", default="lap2d", choices=["lap2d", "powerlaw"])
    ap.add_argument("--cmax", type=int, default=100)
    ap.add_argument("--report", default="20,40,60,100")
    args = ap.parse_args()
    from arnoldi_b200 import _lib, matrices
    if args.matrix == "lap2d":
        side = int(round(args.rows ** 0.5))
        A = matrices.lap2d(side)
    else:
        A = matrices.powerlaw(args.rows)
    report = {int(x) for x in args.report.split(",")}
    res = {"matrix": f"{args.matrix} n={A.shape[0]} nnz={A.nnz}", "cmax": args.cmax}
    res["cgs2_two_sweep"] = run(A, args.cmax, _lib.ORTHO_CGS2, report, 1)
    res["cgs2_fused"] = run(A, args.cmax, _lib.ORTHO_CGS2, report, 3)
    res["mgs"] = run(A, args.cmax, _lib.ORTHO_MGS, report, 0)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
