"""Multi-GPU plumbing of the RPD path: one process per GPU, tets sharded in contiguous blocks with the
sites replicated (the cells of different tets are independent, reference convex_cell.cu:1162-1163),
compact results gathered on one rank over torch.distributed (NCCL over NVLink on the GPU box, gloo
in the CPU tests).  No data-path collective is needed before the gather.

The gathered blob is the concatenation of the ranks' compact blobs in rank order, which IS the
global (tet, site) order because shards are contiguous in tet order; cell byte offsets are rebased
by the preceding ranks' blob sizes, cell ids by the preceding ranks' cell counts.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard(n_tet: int, rank: int, world: int) -> tuple[int, int]:
    """rank r owns tets [first, first + count): contiguous, sizes differ by at most one"""
    base, rem = divmod(n_tet, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


class DeviceView:
    """__cuda_array_interface__ wrapper of a raw device pointer owned by libmat_b200 (mb_rpd_device_buffers)"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (max(int(nbytes), 1),), "typestr": "|u1",
                                         "data": (int(ptr), False), "version": 3}


def as_u8_tensor(ptr: int, nbytes: int, device) -> torch.Tensor:
    return torch.as_tensor(DeviceView(ptr, nbytes), device=device)[:nbytes]


def gather_varlen(local: torch.Tensor, dst: int = 0, group=None, out: torch.Tensor | None = None):
    """Gather 1-D uint8 tensors of different lengths on rank `dst` (all-gather of the sizes, then a
    grouped send/recv).  Returns (concatenation or None, sizes[list]) -- `out` may supply a large
    enough destination buffer to avoid reallocating every step."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = torch.zeros(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(sizes, torch.tensor([local.numel()], dtype=torch.int64, device=local.device), group=group)
    sz = [int(x) for x in sizes.tolist()]
    if world == 1:
        return local, sz
    ops = []
    result = None
    if rank == dst:
        total = sum(sz)
        if out is None or out.numel() < total:
            out = torch.empty(total, dtype=torch.uint8, device=local.device)
        result = out[:total]
        off = 0
        for r in range(world):
            if r == dst:
                result[off:off + sz[r]].copy_(local, non_blocking=True)
            elif sz[r]:
                ops.append(dist.P2POp(dist.irecv, result[off:off + sz[r]], r, group))
            off += sz[r]
    elif local.numel():
        ops.append(dist.P2POp(dist.isend, local, dst, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return result, sz


def rebase_offsets(cell_offsets: list[np.ndarray], blob_sizes: list[int]) -> np.ndarray:
    """per-rank cell byte offsets (each n_cells_r + 1 long, starting at 0) -> global offsets"""
    out = [np.zeros(1, dtype=np.int64)]
    base = 0
    for offs, size in zip(cell_offsets, blob_sizes):
        offs = np.asarray(offs, dtype=np.int64)
        assert offs[0] == 0 and offs[-1] == size
        out.append(offs[1:] + base)
        base += size
    return np.concatenate(out)
