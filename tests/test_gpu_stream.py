"""Streamed run (mb_rpd_run_to_host): the chunked, D2H-overlapped pipeline must deliver exactly the
one-shot result -- same blob bytes, same offsets, same counters -- for any chunk count, in both modes,
for tet ranges and tet subsets."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def one_shot(ctx, **kw):
    res = ctx.run(**kw)
    blob, offs = res.compact()
    out = (blob[: res.compact_bytes // 4].copy(), offs.copy(), res.n_cells, res.n_pairs, res.status_histogram.copy())
    res.free()
    return out


def streamed(ctx, n_chunks, want_records=False, **kw):
    res = ctx.run_to_host(n_chunks=n_chunks, **kw)
    blob, offs = res.host_compact()
    out = (blob.copy(), offs.copy(), res.n_cells, res.n_pairs, res.status_histogram.copy())
    recs = res.records() if want_records else None
    res.free()
    return out, recs


def same(a, b):
    assert a[2] == b[2] and a[3] == b[3], (a[2:4], b[2:4])
    assert np.array_equal(a[4], b[4])
    assert np.array_equal(a[1], b[1])
    assert np.array_equal(a[0], b[0])


@pytest.mark.parametrize("n_chunks", [1, 2, 3, 7, 0])
def test_stream_grid_mode(ctx, cfg1_rt, n_chunks):
    mesh, sites, knn, k = cfg1_rt
    ctx.set_mesh(mesh)
    ctx.upload_sites(sites.site_soa, sites.weights, sites.flags, None, 0)
    want = one_shot(ctx)
    got, _ = streamed(ctx, n_chunks)
    same(want, got)
    got2, _ = streamed(ctx, n_chunks)  # second run reuses the pinned destination
    same(want, got2)


@pytest.mark.parametrize("n_chunks", [1, 4])
def test_stream_given_mode_records(ctx, O, cfg1, cfg1_oracle, n_chunks):
    mesh, sites, knn, k = cfg1
    pt, ps, ra, sa = cfg1_oracle
    ctx.set_mesh(mesh)
    ctx.upload_sites(sites.site_soa, sites.weights, sites.flags, knn, k)
    want = one_shot(ctx)
    got, recs = streamed(ctx, n_chunks, want_records=True)
    same(want, got)
    ref = ra[ra["status"] == 4]
    d = O.defined_equal(ref, recs)
    assert all(v == 0 for f, v in d.items() if f != "cells_compared"), d
    assert np.array_equal(recs["id"], np.arange(len(recs)))


def test_stream_range_and_subset(ctx, cfg1_rt):
    mesh, sites, knn, k = cfg1_rt
    ctx.set_mesh(mesh)
    ctx.upload_sites(sites.site_soa, sites.weights, sites.flags, None, 0)
    ctx.set_tet_range(1000, 7001)
    want = one_shot(ctx)
    got, _ = streamed(ctx, 5)
    same(want, got)
    ctx.set_tet_range(0, -1)
    ids = np.unique(np.random.default_rng(5).integers(0, mesh.n_tet, 3000)).astype(np.int32)
    ctx.set_tet_subset(ids)
    want = one_shot(ctx)
    got, _ = streamed(ctx, 3)
    same(want, got)
    ctx.set_tet_subset(None)


def test_stream_more_chunks_than_tets_and_guards(ctx, cfg1_rt):
    from libmat_b200.capi import LibMatError
    mesh, sites, knn, k = cfg1_rt
    ctx.set_mesh(mesh)
    ctx.upload_sites(sites.site_soa, sites.weights, sites.flags, None, 0)
    ctx.set_tet_range(10, 3)
    want = one_shot(ctx)
    res = ctx.run_to_host(n_chunks=9)
    blob, offs = res.host_compact()
    assert np.array_equal(want[0], blob) and np.array_equal(want[1], offs)
    with pytest.raises(LibMatError):
        res.pairs()
    with pytest.raises(LibMatError):
        res.emit(mesh.n_surf_faces - 1)
    res.free()
    ctx.set_tet_range(0, -1)


def test_shard_upload_with_tet_id_base(ctx, cfg1_rt):
    """a rank that uploads only its contiguous shard of the tets (global vertices) + mb_set_tet_id_base returns
    exactly the corresponding slice of the full run, global tet ids included"""
    mesh, sites, knn, k = cfg1_rt
    ctx.set_mesh(mesh)
    ctx.upload_sites(sites.site_soa, sites.weights, sites.flags, None, 0)
    first, count = 5000, 9000
    ctx.set_tet_range(first, count)
    want = one_shot(ctx)
    ctx.set_tetmesh(mesh.vertices, mesh.indices[first:first + count], mesh.v_adjs, mesh.f_adjs[first:first + count],
                    mesh.f_ids[first:first + count], e_adj6=mesh.e_adj6[first:first + count])
    ctx.set_tet_id_base(first)
    got = one_shot(ctx)
    same(want, got)
    got2, recs = streamed(ctx, 3, want_records=True)
    same(want, got2)
    assert recs["tet_id"].min() >= first and recs["tet_id"].max() < first + count
    pt, ps, st = ctx.run().pairs()
    assert pt.min() >= first and pt.max() < first + count


@pytest.mark.parametrize("fmt,ratio", [(1, 0.75), (2, 0.45)])
@pytest.mark.parametrize("mode", ["grid", "given"])
def test_lean_records_expand_to_the_full_records(ctx, cfg1, cfg1_rt, mode, fmt, ratio):
    """lean (1: records without plane equations) and slim (2: also one neighbour id per bisector instead of three id
    words per plane) transport formats: the host expansion recomputes tet-face planes, power bisectors and plane ids
    and must reproduce the full-format records byte for byte -- plane equations included -- in both modes, through
    fetch_records and through expand_compact"""
    mesh, sites, knn, k = cfg1_rt if mode == "grid" else cfg1
    ctx.set_mesh(mesh)
    ctx.upload_sites(sites.site_soa, sites.weights, sites.flags, None if mode == "grid" else knn, 0 if mode == "grid" else k)
    full = ctx.run()
    want = full.records()
    full_bytes = full.compact_bytes
    full.free()
    res = ctx.run_to_host(n_chunks=3, lean=fmt)
    assert res.n_cells == len(want)
    assert res.compact_bytes < ratio * full_bytes
    got = res.records()
    blob, offs = res.host_compact()
    got2 = ctx.expand_compact(blob, offs)
    res.free()
    for f in want.dtype.names:
        assert np.ascontiguousarray(want[f]).tobytes() == np.ascontiguousarray(got[f]).tobytes(), f
        assert np.ascontiguousarray(want[f]).tobytes() == np.ascontiguousarray(got2[f]).tobytes(), f
    # the one-shot, device-resident run refuses the transport format
    from libmat_b200 import capi
    import ctypes as C
    opts = capi.RpdOpts(0, 0, 0, 0, 1)
    h = C.c_void_p()
    assert ctx.lib.mb_rpd_run(ctx._ctx, C.byref(opts), C.byref(h)) != 0


def test_speculative_capacity_overflow_is_redone(ctx, cfg1_rt):
    """the speculative span launch sizes its pair arrays from the pairs-per-tet seen so far; an estimate that is far
    too small must only cost a redone span (one-shot and streamed), never a different result"""
    mesh, sites, knn, k = cfg1_rt
    ctx.set_mesh(mesh)
    ctx.upload_sites(sites.site_soa, sites.weights, sites.flags, None, 0)
    want = one_shot(ctx)
    for hint in (0.05, 1.0):
        ctx.lib.mb_debug_set_pair_hint(ctx._ctx, hint)
        same(want, one_shot(ctx))
        ctx.lib.mb_debug_set_pair_hint(ctx._ctx, hint)
        got, _ = streamed(ctx, 4)
        same(want, got)
    ctx.lib.mb_debug_set_pair_hint(ctx._ctx, 0.0)  # learn again
    same(want, one_shot(ctx))


def test_unaligned_ranges_concatenate_grid_mode(ctx, cfg1_rt):
    """work-balanced shards start anywhere: tet ranges whose first tet is not a multiple of the K2 cluster size, one-shot
    and streamed, must concatenate to the full grid-mode result byte for byte (offsets rebased)"""
    mesh, sites, knn, k = cfg1_rt
    ctx.set_mesh(mesh)
    ctx.upload_sites(sites.site_soa, sites.weights, sites.flags, None, 0)
    full = one_shot(ctx)
    cuts = [0, 1237, 1238, 5003, 11111, mesh.n_tet]
    for how in ("one_shot", "streamed"):
        blobs, n_cells, n_pairs, base, offs = [], 0, 0, 0, [np.zeros(1, np.int64)]
        for a, b in zip(cuts[:-1], cuts[1:]):
            ctx.set_tet_range(a, b - a)
            part = one_shot(ctx) if how == "one_shot" else streamed(ctx, 3)[0]
            blobs.append(part[0])
            offs.append(part[1][1:] + base)
            base += int(part[1][-1])
            n_cells += part[2]
            n_pairs += part[3]
        ctx.set_tet_range(0, -1)
        assert n_cells == full[2] and n_pairs == full[3]
        assert np.array_equal(np.concatenate(offs), full[1])
        assert np.array_equal(np.concatenate(blobs), full[0])


def test_cluster_search_equals_per_tet_search(cfg1_rt, monkeypatch):
    """K2: the cluster search (one grid walk per 6 tets) and the per-tet search (MB_K2_VARIANT=1) produce the same
    candidate pairs and, from them, the same records"""
    from libmat_b200.rpd import Context
    mesh, sites, knn, k = cfg1_rt
    out = []
    for variant in ("0", "1"):
        monkeypatch.setenv("MB_K2_VARIANT", variant)
        c = Context(0)
        c.set_mesh(mesh)
        c.upload_sites(sites.site_soa, sites.weights, sites.flags, None, 0)
        r = c.run()
        pt, ps, st = r.pairs()
        blob, offs = r.compact()
        out.append((pt.copy(), ps.copy(), st.copy(), blob[: r.compact_bytes // 4].copy(), offs.copy()))
        r.free()
        c.close()
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("env", [{"MB_STREAM_VARIANT": "3", "MB_DEBUG_SMALL_SCRATCH": "1"}, {"MB_STREAM_VARIANT": "3"},
                                 {"MB_STREAM_VARIANT": "1"}])
def test_streamed_run_variants_and_overflow_fallback(cfg1_rt, monkeypatch, env):
    """the host-pipelined staged run (opt-in) -- also forced to overflow its scratch bound, which drains it and redoes
    the run one range at a time -- and the per-span run deliver the one-shot result, like the default staged run"""
    from libmat_b200.rpd import Context
    mesh, sites, knn, k = cfg1_rt
    for key, val in env.items():
        monkeypatch.setenv(key, val)
    c = Context(0)
    c.set_mesh(mesh)
    c.upload_sites(sites.site_soa, sites.weights, sites.flags, None, 0)
    want = one_shot(c)
    for n_chunks in (1, 3, 6):
        got, _ = streamed(c, n_chunks)
        same(want, got)
        got, _ = streamed(c, n_chunks, lean=2)
        assert got[2] == want[2] and got[3] == want[3] and np.array_equal(got[4], want[4])
    c.close()
