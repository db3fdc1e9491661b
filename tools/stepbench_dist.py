#!/usr/bin/env python
"""BASELINE config 5 on several GPUs: Arnoldi-step microbenchmark, SpMV + CGS2 (two-sweep rounds
and the fused three-sweep schedule) versus SpMV + MGS, at n = 1e7 .. 2e8 rows of the 2-D 5-point
Laplacian sharded by rows (each rank generates only its own block), c = 20 .. 100 basis columns.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 \
        tools/stepbench_dist.py --rows 100000000 --cmax 100

A real Arnoldi expansion is run one step at a time (so the basis is genuinely orthonormal and
the DGKS test behaves as in a solve); CUDA-event time and algorithmic bytes per kernel class are
rank 0's (per-GPU figures), step time is the max over ranks."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "arnoldi-py_b200")):
    sys.path.insert(0, p)

CLASSES = ("spmv", "ortho_pass1", "ortho_fused", "ortho_pass2", "mgs")


def run(comm, part, plan, v0loc, n, cmax, kind, report, variant, local, complex_storage):
    from arnoldi_b200.solver import DeviceSolver
    r0, r1 = part.rows(comm.rank)
    out = []
    dev = DeviceSolver(n, cmax, device=local, row0=r0, nrows_local=r1 - r0)
    try:
        if complex_storage:
            dev.set_option("real_mode", 0)
        dev.set_timing(True)
        dev.set_option("ortho_variant", variant)
        if comm.world > 1:
            dev.connect(comm, part)
            dev.set_halo(plan.ghost_cols)
        dev.set_csr(plan.indptr, plan.indices, plan.data)
        dev.set_columns(0, v0loc)
        for j in range(cmax):
            dev.reset_stats()
            dev.timer_start()
            _, n_iter, brk = dev.expand(j, j + 1, 1e-8, ortho=kind)
            ms = comm.max_float(dev.timer_stop())
            if brk:
                break
            c = j + 1
            if c in report:
                st = dev.stats()
                rec = {"c": c, "rounds": st["ortho_rounds"], "step_ms_max_over_ranks": round(ms, 4)}
                for k in CLASSES:
                    if st[k + "_launches"]:
                        rec[k] = {"ms": round(st[k + "_ms"], 4),
                                  "gbs_per_gpu": round(st[k + "_bytes"] / st[k + "_ms"] / 1e6, 1)}
                out.append(rec)
    finally:
        if comm.world > 1:
            dev.disconnect()
            comm.barrier()
        dev.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--cmax", type=int, default=100)
    ap.add_argument("--report", default="20,40,60,100")
    ap.add_argument("--complex-storage", action="store_true")
    ap.add_argument("--skip-mgs", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    from arnoldi_b200 import _lib, matrices
    from arnoldi_b200.distributed import RowPartition, TorchComm, build_halo_plan
    comm = TorchComm()
    side = int(round(args.rows ** 0.5))
    n = side * side
    part = RowPartition(n, comm.world)
    r0, r1 = part.rows(rank)
    plan = build_halo_plan(matrices.lap2d_rows(side, r0, r1))
    # a unit-norm start vector without drawing n normals on every rank: rank-local stream
    rng = np.random.default_rng(1234 + rank)
    v0loc = rng.standard_normal(r1 - r0).astype(np.complex128)
    sq = sum(np.frombuffer(b, np.float64)[0] for b in comm.all_gather_bytes(
        np.array([np.vdot(v0loc, v0loc).real]).tobytes()))
    v0loc /= np.sqrt(sq)
    report = {int(x) for x in args.report.split(",")}
    res = {"matrix": f"lap2d({side}) n={n} sharded x{comm.world}", "cmax": args.cmax,
           "storage": "complex128" if args.complex_storage else "float64 (real operator, real v0)"}
    a = (comm, part, plan, v0loc, n, args.cmax)
    res["cgs2_two_sweep"] = run(*a, _lib.ORTHO_CGS2, report, 1, local, args.complex_storage)
    res["cgs2_fused"] = run(*a, _lib.ORTHO_CGS2, report, 3, local, args.complex_storage)
    if not args.skip_mgs:
        res["mgs"] = run(*a, _lib.ORTHO_MGS, report, 0, local, args.complex_storage)
    if rank == 0:
        print(json.dumps(res), flush=True)
    comm.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
