"""The multi-GPU host logic (tet shards + variable-length gather) on CPU: world_size 2, gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from libmat_b200.dist import gather_varlen, shard
    n_tet = 1001
    first, count = shard(n_tet, rank, world)
    # fake compact blobs: tet t contributes (t % 5) cells of 8 + t % 3 bytes, each byte = t & 0xff
    rng_bytes = []
    for t in range(first, first + count):
        for c in range(t % 5):
            rng_bytes.append(np.full(8 + t % 3, t & 0xff, np.uint8))
    local = torch.from_numpy(np.concatenate(rng_bytes) if rng_bytes else np.zeros(0, np.uint8))
    got, sizes = gather_varlen(local, dst=0)
    if rank == 0:
        q.put((got.numpy().copy(), sizes))
    else:
        assert got is None
    # an empty shard on one side must not hang
    empty = torch.zeros(0, dtype=torch.uint8) if rank == 1 else local
    got2, sizes2 = gather_varlen(empty, dst=0)
    if rank == 0:
        q.put((got2.numpy().copy(), sizes2))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_partition():
    from libmat_b200.dist import shard
    for n, w in ((10, 3), (196608, 8), (7, 8), (2058000, 8)):
        parts = [shard(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == n
        for (f0, c0), (f1, _) in zip(parts[:-1], parts[1:]):
            assert f0 + c0 == f1
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def test_gather_varlen_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, sizes = q.get(timeout=120)
    got2, sizes2 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = []
    for t in range(1001):
        for c in range(t % 5):
            want.append(np.full(8 + t % 3, t & 0xff, np.uint8))
    want = np.concatenate(want)
    assert sum(sizes) == want.size and np.array_equal(got, want)  # rank order == tet order
    assert sizes2[1] == 0 and np.array_equal(got2, want[:sizes2[0]])


def test_rebase_offsets():
    from libmat_b200.dist import rebase_offsets
    a = np.array([0, 10, 30])
    b = np.array([0, 5])
    c = np.array([0])
    out = rebase_offsets([a, c, b], [30, 0, 5])
    assert out.tolist() == [0, 10, 30, 35]


def _balanced_worker(rank, world, port, q):
    import numpy as np
    import torch.distributed as dist

    from libmat_b200.dist import balanced_shards, shard
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n = 1001
    cells = (np.arange(n) % 7 + (np.arange(n) > 600) * 9).astype(np.int64)  # heavier tail
    first, count = shard(n, rank, world)
    cuts, _ = balanced_shards(n, first, cells[first:first + count])
    from libmat_b200.dist import allreduce_site_volumes
    vol, bary = allreduce_site_volumes(np.full(5, rank + 1.0, np.float32), np.arange(15, dtype=np.float32) * (rank + 1))
    assert np.allclose(vol, 3.0) and np.allclose(bary, np.arange(15) * 3.0)
    q.put((rank, cuts.tolist()))
    dist.destroy_process_group()


def test_balanced_shards_world2():
    """work-balanced contiguous shards: every rank computes the same cuts; the heavier tail gets fewer tets"""
    import numpy as np
    import torch.multiprocessing as mp

    from libmat_b200.dist import balanced_cuts
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29641
    ps = [ctx.Process(target=_balanced_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
    assert got[0] == got[1]
    cuts = got[0]
    n = 1001
    cells = (np.arange(n) % 7 + (np.arange(n) > 600) * 9).astype(np.int64)
    assert cuts[0] == 0 and cuts[-1] == n and cuts[1] > n // 2
    w = cells + 1.5
    assert abs(w[:cuts[1]].sum() - w[cuts[1]:].sum()) <= 2 * w.max()  # the cut falls within one tet of the midpoint
    # measured rebalancing: the second half really costs 3x per unit of weight -> the cut moves right
    from libmat_b200.dist import rebalance_cuts
    true_cost = w * np.where(np.arange(n) > 600, 3.0, 1.0)
    c = np.array(cuts)
    for _ in range(4):
        c = rebalance_cuts(c, w, np.array([true_cost[c[0]:c[1]].sum(), true_cost[c[1]:c[2]].sum()]))
    assert c[1] > cuts[1] and abs(true_cost[:c[1]].sum() - true_cost[c[1]:].sum()) <= 0.02 * true_cost.sum()
    # pure function: degenerate inputs
    assert balanced_cuts(np.zeros(0), 4).tolist() == [0, 0, 0, 0, 0]
    assert balanced_cuts(np.ones(3), 8)[-1] == 3 and (np.diff(balanced_cuts(np.ones(3), 8)) >= 0).all()
