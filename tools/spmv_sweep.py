#!/usr/bin/env python
"""SpMV kernel-shape sweep on one GPU: one operator, several option sets; per set the SpMV is
timed (CUDA events, inside a real Arnoldi expansion so x and y are basis columns in HBM) in
real (float64) and complex128 storage.

    python tools/spmv_sweep.py --matrix lap2d --size 4096 "spmv_variant=2" "spmv_tile=1280,spmv_stages=3" ...
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "arnoldi-py_b200")):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--matrix", default="lap2d")
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--modes", default="real,complex")
    ap.add_argument("--algo", default="auto")
    ap.add_argument("sets", nargs="*", default=[""])
    args = ap.parse_args()
    from arnoldi_b200 import matrices
    from arnoldi_b200.solver import DeviceSolver
    from arnoldi_b200.utils import rand_normalized_vector
    A = getattr(matrices, args.matrix)(args.size)
    n = A.shape[0]
    np.random.seed(0)
    v0 = rand_normalized_vector(n, np.complex128)
    peak = 6522.7
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    for mode in args.modes.split(","):
        for spec in args.sets:
            with DeviceSolver(n, args.steps + 2) as dev:
                if mode == "complex":
                    dev.set_option("real_mode", 0)
                for kv in filter(None, spec.split(",")):
                    k, v = kv.split("=")
                    dev.set_option(k, int(v))
                dev.set_timing(True)
                dev.set_csr(A.indptr, A.indices, A.data, algo=args.algo)
                dev.set_columns(0, v0)
                dev.expand(0, 2, 1e-8)                   # warm-up
                dev.reset_stats()
                dev.expand(2, args.steps + 2, 1e-8)
                st = dev.stats()
            ms = st["spmv_ms"] / st["spmv_launches"]
            gbs = st["spmv_bytes"] / st["spmv_ms"] / 1e6
            print(json.dumps({"matrix": f"{args.matrix}({args.size})", "mode": mode, "options": spec,
                              "spmv_ms": round(ms, 4), "gbs": round(gbs, 1),
                              "frac_measured": round(gbs / peak, 3), "frac_8tbs": round(gbs / 8000, 3)}),
                  flush=True)


if __name__ == "__main__":
    main()
