"""Host logic of the multi-GPU path on CPU: partition + halo plan against scipy slicing
(bit-exact), and the same under a real world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import csr_from_golden


def _matrices(golden):
    from arnoldi_b200.matrices import lap2d, mark
    rng = np.random.default_rng(4)
    R = sp.random(501, 501, density=0.01, random_state=rng, format="csr")
    return {"mark50": csr_from_golden(golden("matrices"), "mark50"), "lap2d_23": lap2d(23),
            "random": R, "mark7": mark(7)}


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_partition_and_halo_plan_match_scipy(golden, world):
    from arnoldi_b200.distributed import RowPartition, build_halo_plan, slice_rows
    for name, A in _matrices(golden).items():
        n = A.shape[0]
        part = RowPartition(n, world)
        assert part.starts[0] == 0 and part.starts[-1] == n
        sizes = np.diff(part.starts)
        assert sizes.max() - sizes.min() <= 1           # balanced by rows
        rng = np.random.default_rng(1)
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        y = A @ x
        covered = 0
        for r in range(world):
            r0, r1 = part.rows(r)
            blk = slice_rows(A, r0, r1)
            S = A[r0:r1]                                 # scipy's own row slice
            np.testing.assert_array_equal(blk.indptr, S.indptr)
            np.testing.assert_array_equal(blk.indices, S.indices)
            np.testing.assert_array_equal(blk.data, S.data)
            plan = build_halo_plan(blk)
            g = plan.ghost_cols
            assert np.all(np.diff(g) > 0) and not np.any((g >= r0) & (g < r1))
            assert np.all(part.owner(g) != r)
            # renumbered block applied to [local x ; ghost x] == rows of A @ x, bit for bit
            L = sp.csr_matrix((plan.data, plan.indices, plan.indptr),
                              shape=(r1 - r0, (r1 - r0) + len(g)))
            np.testing.assert_array_equal(L @ np.concatenate([x[r0:r1], x[g]]), y[r0:r1])
            # inverse map gives back the global ids
            back = np.where(plan.indices < plan.nloc, plan.indices + r0,
                            g[np.maximum(plan.indices - plan.nloc, 0)] if len(g) else 0)
            np.testing.assert_array_equal(back, S.indices)
            covered += r1 - r0
        assert covered == n
    # banded operators only talk to their neighbours
    from arnoldi_b200.matrices import lap2d
    A = lap2d(16)
    part = RowPartition(256, 4)
    plan = build_halo_plan(slice_rows(A, *part.rows(1)))
    assert set(part.owner(plan.ghost_cols)) == {0, 2} and len(plan.ghost_cols) == 32


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(os.path.dirname(here), "arnoldi-py_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from arnoldi_b200.distributed import RowPartition, TorchComm, build_halo_plan, slice_rows
        from arnoldi_b200.matrices import mark
        comm = TorchComm()
        assert comm.rank == rank and comm.world == world
        A = mark(30)
        n = A.shape[0]
        part = RowPartition(n, world)
        r0, r1 = part.rows(rank)
        plan = build_halo_plan(slice_rows(A, r0, r1))
        np.random.seed(0)
        x = np.random.randn(n) + 0j                      # same stream on every rank
        # "halo exchange": every rank publishes its block, takes the ghost entries it needs
        blocks = comm.all_gather_bytes(x[r0:r1].tobytes())
        xg = np.concatenate([np.frombuffer(b, np.complex128) for b in blocks])
        ghost = xg[plan.ghost_cols]
        L = sp.csr_matrix((plan.data, plan.indices, plan.indptr),
                          shape=(r1 - r0, (r1 - r0) + len(plan.ghost_cols)))
        y_loc = L @ np.concatenate([x[r0:r1], ghost])
        ok = np.array_equal(y_loc, (A @ x)[r0:r1])
        # rank-ordered sum of partial inner products is the same bits on every rank
        part_dot = np.vdot(x[r0:r1], y_loc)
        parts = comm.all_gather_bytes(np.array([part_dot]).tobytes())
        total = sum(np.frombuffer(b, np.complex128)[0] for b in parts)
        totals = comm.all_gather_bytes(np.array([total]).tobytes())
        same = all(t == totals[0] for t in totals)
        mx = comm.max_float(float(rank))
        comm.barrier()
        q.put((rank, bool(ok), bool(same), mx))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(o[0] for o in out) == [0, 1]
    assert all(o[1] and o[2] for o in out)
    assert all(o[3] == 1.0 for o in out)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_halo_push_plan_is_the_transpose_of_the_pull_plan(world):
    """Owner-side push lists: what rank q sends to rank r must be exactly the entries of r's
    ghost list that q owns, in r's order, landing at the right offset of r's ghost buffer."""
    from arnoldi_b200.distributed import (RowPartition, build_halo_plan, plan_halo_push,
                                          slice_rows)
    from arnoldi_b200.matrices import lap2d, powerlaw
    for A in (lap2d(19), powerlaw(3000)):
        n = A.shape[0]
        part = RowPartition(n, world)
        ghosts = [build_halo_plan(slice_rows(A, *part.rows(r))).ghost_cols for r in range(world)]
        rng = np.random.default_rng(0)
        x = rng.standard_normal(n)
        received = [np.full(len(g), np.nan) for g in ghosts]
        for q in range(world):
            q0, q1 = part.rows(q)
            send_idx, send_ptr, dst_off = plan_halo_push(ghosts, part, q)
            assert send_ptr[0] == 0 and send_ptr[q + 1] == send_ptr[q]      # nothing to itself
            for r in range(world):
                rows = send_idx[send_ptr[r]:send_ptr[r + 1]]
                assert np.all((rows >= 0) & (rows < q1 - q0)) and np.all(np.diff(rows) > 0)
                received[r][dst_off[r]:dst_off[r] + len(rows)] = x[q0 + rows]   # the "push"
        for r in range(world):
            np.testing.assert_array_equal(received[r], x[ghosts[r]])          # every slot filled
