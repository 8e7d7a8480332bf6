"""The fused multi-GPU gather (libmat_b200.dist.ShardSink over mb_rpd_run_to_sink): two ranks -- on two GPUs
when the box has them, else both on cuda:0 -- shard the tets, stream their ordered records into rank 0's
device memory through a CUDA-IPC peer mapping ("device") or into a shared page-locked host segment ("host");
rank 0 must end up with exactly the single-process result (same bytes, same offsets)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, kind, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from libmat_b200 import synth
    from libmat_b200.dist import ShardSink, shard
    from libmat_b200.rpd import Context
    ndev = torch.cuda.device_count()
    device = rank % ndev
    torch.cuda.set_device(device)
    mesh = synth.make_ball_mesh(12)
    sites = synth.make_spheres(600)
    ctx = Context(device)
    ctx.set_mesh(mesh)
    first, count = shard(mesh.n_tet, rank, world)
    ctx.set_tet_range(first, count)
    ctx.upload_sites(sites.site_soa, sites.weights, sites.flags, None, 0)
    sink = ShardSink(ctx, 12 << 20, 60000, kind=kind, tag=f"mb_test_{port}")
    for n_chunks in (1, 3):
        res, directory = sink.run(n_chunks=n_chunks)
        assert directory[rank, 0] == res.compact_bytes and directory[rank, 1] == res.n_cells
        res.free()
        if rank == 0:
            blob, offs = sink.read_host(directory)
            np.savez(os.path.join(out_dir, f"got_{kind}_{n_chunks}.npz"), blob=blob, offs=offs)
        dist.barrier()
    # lean transport format through the sink; rank 0 (which holds the whole mesh + sites) expands all shards
    res, directory = sink.run(n_chunks=2, lean=2)  # slim transport format
    res.free()
    if rank == 0:
        blob, offs = sink.read_host(directory)
        recs = ctx.expand_compact(blob.view(np.uint32) if blob.size % 4 == 0 else blob, offs)
        np.save(os.path.join(out_dir, f"lean_{kind}.npy"), recs)
    dist.barrier()
    # per-site volume / barycentre sums of the shards add up to the single-process sums
    from libmat_b200.dist import allreduce_site_volumes
    rv = ctx.run(want_volumes=True)
    vol, bary = allreduce_site_volumes(*rv.site_volumes())
    rv.free()
    if rank == 0:
        np.savez(os.path.join(out_dir, f"vol_{kind}.npz"), vol=vol, bary=bary)
    dist.barrier()
    # a sink that is too small is an error, not a truncation
    small = ShardSink(ctx, 4096, 8, kind=kind, tag=f"mb_test_small_{port}")
    try:
        small.run(n_chunks=2)[0].free()
        raise AssertionError("expected MB_ERR_NOMEM")
    except Exception as exc:  # LibMatError
        assert "sink too small" in str(exc), exc
    dist.barrier()
    small.close()
    sink.close()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["device", "host"])
def test_two_rank_sink_equals_single_process(ctx, synth, tmp_path, kind):
    import torch.multiprocessing as mp
    mesh = synth.make_ball_mesh(12)
    sites = synth.make_spheres(600)
    ctx.set_mesh(mesh)
    ctx.upload_sites(sites.site_soa, sites.weights, sites.flags, None, 0)
    rv = ctx.run(want_volumes=True)
    want_vol, want_bary = rv.site_volumes()
    rv.free()
    res = ctx.run()
    blob, offs = res.compact()
    want_blob = blob[: res.compact_bytes // 4].view(np.uint8).copy()
    want_offs = offs.copy()
    want_recs = res.records()
    res.free()
    mpc = mp.get_context("spawn")
    port = _free_port()
    procs = [mpc.Process(target=_worker, args=(r, 2, port, kind, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    for p in procs:
        if p.is_alive():
            p.kill()
            pytest.fail("worker hung")
        assert p.exitcode == 0
    for n_chunks in (1, 3):
        got = np.load(tmp_path / f"got_{kind}_{n_chunks}.npz")
        assert np.array_equal(got["offs"], want_offs)
        assert np.array_equal(got["blob"], want_blob)
    lean = np.load(tmp_path / f"lean_{kind}.npy")
    assert lean.tobytes() == want_recs.tobytes()
    v = np.load(tmp_path / f"vol_{kind}.npz")
    # float atomics in a different order: 1e-4 of the largest sum
    assert np.abs(v["vol"] - want_vol).max() <= 1e-4 * np.abs(want_vol).max()
    assert np.abs(v["bary"] - want_bary).max() <= 1e-4 * np.abs(want_bary).max()
