#!/usr/bin/env python
"""Latency of the fused peer reduction (csrc/ortho.cu `peer_allreduce`) at the rank count of
the launch: a 1-block kernel that does nothing but one exchange, launched back to back.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/comm_bench.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "arnoldi-py_b200")):
    sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    from arnoldi_b200.distributed import RowPartition, TorchComm
    from arnoldi_b200.solver import DeviceSolver
    comm = TorchComm()
    n = 1024 * comm.world
    part = RowPartition(n, comm.world)
    r0, r1 = part.rows(rank)
    dev = DeviceSolver(n, 8, device=local, row0=r0, nrows_local=r1 - r0)
    dev.connect(comm, part)
    out = []
    for iters in (200, 2000, 2000):
        comm.barrier()
        out.append(comm.max_float(dev.comm_bench(iters)))
    # NCCL all-reduce of the same payload (57 doubles), for scale: launch + ring/tree latency
    x = torch.zeros(57, dtype=torch.float64, device="cuda")
    for _ in range(20):
        dist.all_reduce(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(1000):
        dist.all_reduce(x)
    b.record()
    torch.cuda.synchronize()
    nccl_us = comm.max_float(a.elapsed_time(b))
    dev.disconnect()
    comm.barrier()
    dev.close()
    if rank == 0:
        print(json.dumps({"ranks": comm.world, "peer_exchange_us": [round(v, 2) for v in out],
                          "includes": "kernel launch + remote stores + system fence + flag wait + sum",
                          "nccl_allreduce_57_doubles_us": round(nccl_us, 2)}), flush=True)
    comm.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
