#!/usr/bin/env python
"""What the per-kernel CUDA events cost: config-2 restart cycles with the instrumentation switched
on and off in alternation (same clocks, same data), wall time per cycle (a cycle ends with a
stream synchronise, so the host clock brackets the device work).  1 GPU, or N under torchrun.

    python tools/event_cost.py [--grid 4096] [--cycles 40]
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
    ap.add_argument("--grid", type=int, default=4096)
    ap.add_argument("--cycles", type=int, default=40)
    args = ap.parse_args()
    from arnoldi_b200.matrices import lap2d
    from arnoldi_b200.rotate import rotate
    from arnoldi_b200.solver import DeviceSolver
    from arnoldi_b200.utils import arg_largest_real, rand_normalized_vector
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    MAX_DIM, P = 40, 15
    A = lap2d(args.grid)
    n = A.shape[0]
    np.random.seed(0)
    v0 = rand_normalized_vector(n, np.complex128)
    H = np.zeros((MAX_DIM + 1, MAX_DIM), np.complex128)
    comm = None
    if world > 1:
        import torch
        import torch.distributed as dist
        from arnoldi_b200.distributed import RowPartition, TorchComm, build_halo_plan, slice_rows
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        comm = TorchComm()
        part = RowPartition(n, world)
        r0, r1 = part.rows(rank)
        plan = build_halo_plan(slice_rows(A, r0, r1))
        dev = DeviceSolver(n, MAX_DIM, device=local, row0=r0, nrows_local=r1 - r0)
        dev.connect(comm, part)
        dev.set_halo(plan.ghost_cols)
        dev.set_csr(plan.indptr, plan.indices, plan.data)
        dev.set_columns(0, v0[r0:r1])
    else:
        dev = DeviceSolver(n, MAX_DIM, device=local)
        dev.set_csr(A.indptr, A.indices, A.data)
        dev.set_columns(0, v0)

    def grow(start):
        cols, n_iter, brk = dev.expand(start, MAX_DIM, 1e-8)
        for j in range(start, n_iter):
            H[: j + 2, j] = cols[: j + 2, j]

    def cycle():
        T2, Q = rotate(H[:MAX_DIM, :MAX_DIM], arg_largest_real, fast_real=True)
        spike = H[MAX_DIM, :MAX_DIM] @ Q[:, :P]
        dev.restart(Q, MAX_DIM, P)
        H[:P, :P] = T2[:P, :P]
        H[P, :P] = spike
        H[P, P:] = 0
        grow(P)

    grow(0)
    for _ in range(4):
        cycle()
    t = {True: [], False: []}
    for i in range(args.cycles):
        on = (i % 2 == 0)
        dev.set_timing(on)
        if comm is not None:
            comm.barrier()
        t0 = time.perf_counter()
        cycle()
        dt = time.perf_counter() - t0
        if comm is not None:
            dt = comm.max_float(dt)
        t[on].append(dt)
    dev.stats()
    if comm is not None:
        dev.disconnect()
        comm.barrier()
    dev.close()
    if rank == 0:
        a, b = 1e3 * np.median(t[True]), 1e3 * np.median(t[False])
        print(json.dumps({"ranks": world, "grid": args.grid, "cycles_each": args.cycles // 2,
                          "ms_per_cycle_with_kernel_events": round(a, 4),
                          "ms_per_cycle_without": round(b, 4), "event_cost_ms_per_cycle": round(a - b, 4),
                          "launches_per_cycle": 101}))
    if comm is not None:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
