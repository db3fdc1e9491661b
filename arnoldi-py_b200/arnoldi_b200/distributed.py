"""Host-side logic of the multi-GPU path: row partition, halo plan, bootstrap.

One process per GPU.  A and the Krylov basis are block-row sharded; rank r owns the
contiguous rows ``[starts[r], starts[r+1])``.  The reference has no distributed path
(SURVEY.md section 5), so nothing here mirrors a reference file; the partition and the
column renumbering are checked bit-exactly against scipy row slicing in the tests.

Everything in this module is NumPy on the host and runs without a GPU.
"""
from __future__ import annotations

import dataclasses

import numpy as np
import scipy.sparse as sp


@dataclasses.dataclass
class RowPartition:
    """Contiguous, balanced-by-rows partition of ``n`` rows over ``world`` ranks."""

    n: int
    world: int

    def __post_init__(self):
        assert 1 <= self.world <= 8, "one box: at most 8 ranks"
        assert self.n >= self.world, "fewer rows than ranks"
        self.starts = np.array([(r * self.n) // self.world for r in range(self.world + 1)],
                               dtype=np.int64)

    def rows(self, rank):
        return int(self.starts[rank]), int(self.starts[rank + 1])

    def owner(self, cols):
        """Rank owning each global row/column id."""
        return np.searchsorted(self.starts, np.asarray(cols), side="right") - 1


@dataclasses.dataclass
class RowBlock:
    """Rows ``[row0, row0 + nrows)`` of a global CSR matrix, columns still GLOBAL ids.

    What a rank hands to ``partial_schur`` when it never holds the whole matrix
    (e.g. the 100 M-row synthetic operator generated shard by shard)."""

    indptr: np.ndarray
    indices: np.ndarray
    data: np.ndarray
    row0: int
    shape: tuple

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def nrows(self):
        return self.indptr.shape[0] - 1


def slice_rows(A, r0, r1):
    """Rows [r0, r1) of a scipy CSR matrix as a RowBlock (bit-identical to ``A[r0:r1]``)."""
    assert sp.issparse(A) and A.format == "csr"
    lo, hi = int(A.indptr[r0]), int(A.indptr[r1])
    return RowBlock(A.indptr[r0:r1 + 1] - A.indptr[r0], A.indices[lo:hi], A.data[lo:hi], r0,
                    tuple(A.shape))


@dataclasses.dataclass
class HaloPlan:
    """Local CSR block in LOCAL column numbering + the remote entries of v it reads.

    Column id < nloc  -> local row id (global id - row0)
    Column id >= nloc -> nloc + index into ``ghost_cols`` (sorted global ids, so the entries
                         owned by one peer are contiguous and ascending)
    """

    indptr: np.ndarray
    indices: np.ndarray       # int32, local numbering
    data: np.ndarray
    ghost_cols: np.ndarray    # int64, strictly increasing global ids outside the block
    row0: int
    nloc: int


def build_halo_plan(block: RowBlock) -> HaloPlan:
    r0, nloc = block.row0, block.nrows
    idx = np.asarray(block.indices)
    local = (idx >= r0) & (idx < r0 + nloc)
    outside = idx[~local].astype(np.int64)          # usually a tiny fraction of the entries
    ghost = np.unique(outside)
    assert nloc + ghost.shape[0] < 2**31, "local + ghost columns exceed int32"
    renum = (idx - r0).astype(np.int32)              # wraps for outside entries: overwritten next
    if outside.shape[0]:
        renum[~local] = (nloc + np.searchsorted(ghost, outside)).astype(np.int32)
    return HaloPlan(np.asarray(block.indptr), renum, np.asarray(block.data), ghost, r0, nloc)


def plan_halo_push(ghost_lists, partition: RowPartition, rank: int):
    """Owner-side view of the halo: given every rank's sorted ghost list, work out which LOCAL
    rows of ``rank`` each peer reads and where they land in that peer's ghost buffer.

    Returns ``(send_idx, send_ptr, dst_off)``: rows ``send_idx[send_ptr[r]:send_ptr[r+1]]`` (local
    numbering, ascending) go to rank r, starting at index ``dst_off[r]`` of r's ghost buffer.
    """
    r0, r1 = partition.rows(rank)
    send, ptr, dst = [], [0], []
    for r, g in enumerate(ghost_lists):
        g = np.asarray(g, dtype=np.int64)
        lo, hi = (0, 0) if r == rank else (int(x) for x in np.searchsorted(g, [r0, r1]))
        send.append(g[lo:hi] - r0)
        ptr.append(ptr[-1] + (hi - lo))
        dst.append(lo)
    send_idx = np.ascontiguousarray(np.concatenate(send), dtype=np.int64)
    return send_idx, np.array(ptr, dtype=np.int64), np.array(dst, dtype=np.int64)


class TorchComm:
    """Bootstrap over torch.distributed (NCCL on the GPU box, gloo in CPU tests): only
    small host objects travel through it (IPC handles, barriers, max of timings)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_gather_bytes(self, payload: bytes):
        out = [None] * self.world
        self.dist.all_gather_object(out, payload, group=self.group)
        return out

    def barrier(self):
        self.dist.barrier(group=self.group)

    def max_float(self, x: float) -> float:
        out = [None] * self.world
        self.dist.all_gather_object(out, float(x), group=self.group)
        return max(out)
