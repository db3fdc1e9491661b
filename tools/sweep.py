#!/usr/bin/env python
"""Kernel-variant sweep on one GPU: same restart cycles as bench.py, one solver, several
option sets; prints per-kernel-class average time and achieved GB/s (CUDA events).

    python tools/sweep.py --grid 4096 --cycles 2 "ortho_variant=0" "ortho_variant=2" ...
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "arnoldi-py_b200")):
    sys.path.insert(0, p)

from scipy.linalg import schur  # noqa: E402

from arnoldi_b200.solver import DeviceSolver  # noqa: E402
from arnoldi_b200.utils import arg_largest_real, ordered_schur, rand_normalized_vector  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--matrix", default="lap2d")
    ap.add_argument("--nev", type=int, default=10)
    ap.add_argument("--max-dim", type=int, default=40)
    ap.add_argument("--cycles", type=int, default=2)
    ap.add_argument("--ortho", default="cgs2")
    ap.add_argument("--complex-storage", action="store_true")
    ap.add_argument("sets", nargs="*", default=[""])
    args = ap.parse_args()
    from arnoldi_b200 import _lib
    from arnoldi_b200 import matrices
    A = getattr(matrices, args.matrix)(args.grid)
    n = A.shape[0]
    m, nev = args.max_dim, args.nev
    p = min(nev + 5, m - 1)
    kind = _lib.ORTHO_MGS if args.ortho == "mgs" else _lib.ORTHO_CGS2
    np.random.seed(0)
    v0 = rand_normalized_vector(n, np.complex128)
    H = np.zeros((m + 1, m), np.complex128)
    dev = DeviceSolver(n, m)
    if args.complex_storage:
        dev.set_option("real_mode", 0)
    dev.set_timing(True)
    dev.set_csr(A.indptr, A.indices, A.data)
    dev.set_columns(0, v0)

    def grow(start):
        cols, n_iter, brk = dev.expand(start, m, 1e-8, ortho=kind)
        for j in range(start, n_iter):
            H[: j + 2, j] = cols[: j + 2, j]

    def cycle():
        T1, Q1 = schur(H[:m, :m], output="complex")
        T2, Q2 = ordered_schur(T1, output="complex", sort_function=arg_largest_real)
        Q = Q1 @ Q2
        spike = H[m, :m] @ Q[:, :p]
        dev.restart(Q, m, p)
        H[:p, :p] = T2[:p, :p]
        H[p, :p] = spike
        H[p, p:] = 0
        grow(p)

    grow(0)
    cycle()
    for spec in args.sets:
        opts = dict(kv.split("=") for kv in spec.split(",") if kv)
        for k in ("ortho_variant", "fused_ct", "grid_mult", "restart_variant", "fused_stages", "fused_r", "spmv_variant"):
            dev.set_option(k, int(opts.get(k, 0)))
        if "spmv_tile" in opts or "spmv_threads" in opts:
            dev.set_option("spmv_tile", int(opts.get("spmv_tile", 0)))
            dev.set_option("spmv_threads", int(opts.get("spmv_threads", 0)))
            dev.set_csr(A.indptr, A.indices, A.data)
        cycle()
        dev.reset_stats()
        dev.timer_start()
        for _ in range(args.cycles):
            cycle()
        ms = dev.timer_stop()
        st = dev.stats()
        out = {"opts": spec, "ms_per_cycle": ms / args.cycles, "real_storage": st["real_storage"]}
        for key in ("spmv", "ortho_pass1", "ortho_fused", "ortho_pass2", "mgs", "restart"):
            if st[key + "_launches"]:
                out[key] = {"avg_ms": round(st[key + "_ms"] / st[key + "_launches"], 4),
                            "gbs": round(st[key + "_bytes"] / st[key + "_ms"] / 1e6, 1),
                            "n": st[key + "_launches"]}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
