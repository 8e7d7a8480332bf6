"""GPU parity tests of the RPD path, through the C ABI (libmat_b200.so) against the oracle, the
reference build (oracle/_ref, when present) and the committed golden vectors."""
import numpy as np
import pytest

from conftest import golden
from oracle.gen_golden import kat1_inputs, mini_inputs

pytestmark = pytest.mark.gpu

FIELDS = ("status", "voro_id", "tet_id", "weight", "nb_v", "nb_p", "nb_e", "ver", "id2", "edge")


def expected_hist(recs, stat):
    """per-pair Status histogram (index = status+1): `success` where the reference copied a valid record
    (records of cells with |vol| < 0.1 stay valid although gpu_stat flips afterwards, convex_cell.cu:1040),
    else the reference's gpu_stat value."""
    st = np.where(recs["status"] == 4, 4, stat)
    return np.bincount(st + 1, minlength=10)


def assert_defined_equal(O, a, b):
    d = O.defined_equal(a, b)
    assert all(v == 0 for f, v in d.items() if f != "cells_compared"), d
    assert d["cells_compared"] == len(a)


def run_given(ctx, mesh, sites, knn, k, **kw):
    ctx.set_mesh(mesh)
    return ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, knn, k, **kw)


def test_kat1_golden(ctx, O):
    g = golden("kat1_rpd.npz")
    mesh, sites, knn, k = kat1_inputs()
    res = run_given(ctx, mesh, sites, knn, k)
    recs = res.records()
    assert res.n_cells == 2 and res.n_pairs == 2
    for f in FIELDS:
        assert np.array_equal(recs[f], g[f]), f
    assert np.array_equal(recs["clip"][..., :5].view(np.uint32), g["clip"][..., :5].view(np.uint32))
    assert recs["id"].tolist() == [0, 1] and (recs["is_active"] == 1).all()
    assert (recs["euler"] == -1).all() and (recs["cell_vol"] == -1).all()  # copy(), convex_cell.cu:933-949


@pytest.mark.parametrize("lanes", [8, 16, 32])
def test_mini_golden(ctx, O, lanes):
    g = golden("mini_rpd.npz")
    mesh, sites, knn, k = mini_inputs()
    res = run_given(ctx, mesh, sites, knn, k, lanes_per_cell=lanes)
    ok = g["status"] == 4
    assert res.n_pairs == len(g["pair_tet"])
    assert res.n_cells == int(ok.sum())
    recs = res.records()
    for f in FIELDS:
        assert np.array_equal(recs[f], g[f][ok]), f
    assert np.array_equal(recs["clip"][..., :5].view(np.uint32), g["clip"][ok][..., :5].view(np.uint32))
    # per-pair status histogram (index = status + 1) equals the reference's
    assert np.array_equal(res.status_histogram, expected_hist(g, g["stat"]))


@pytest.mark.parametrize("lanes", [8, 16, 32])
def test_cfg1_given_vs_oracle(ctx, O, cfg1, cfg1_oracle, lanes):
    """BASELINE configs[0] (20 250 tets, 1 000 spheres, k=80): byte-identical on defined entries."""
    mesh, sites, knn, k = cfg1
    pt, ps, ra, sa = cfg1_oracle
    res = run_given(ctx, mesh, sites, knn, k, lanes_per_cell=lanes)
    assert res.n_pairs == len(pt)
    want = ra[ra["status"] == 4]
    assert res.n_cells == len(want)
    recs = res.records()
    assert_defined_equal(O, want, recs)
    assert np.array_equal(recs["id"], np.arange(len(recs)))
    key = recs["tet_id"].astype(np.int64) * sites.n_site + recs["voro_id"]
    assert (np.diff(key) > 0).all()  # sorted by (tet, site), voronoi.cu:744-769
    assert np.array_equal(res.status_histogram, expected_hist(ra, sa))


def test_cfg1_given_vs_reference_build(ctx, O, cfg1, cfg1_oracle):
    if O.ref("rpd") is None:
        pytest.skip("oracle/_ref not built")
    mesh, sites, knn, k = cfg1
    pt, ps, _, _ = cfg1_oracle
    rb, sb, _ = O.run_pairs(mesh, sites, knn, k, pt, ps, impl="ref")
    res = run_given(ctx, mesh, sites, knn, k)
    assert_defined_equal(O, rb[rb["status"] == 4], res.records())


def test_overflow_classes(ctx, O, synth):
    """dense sites on a tiny mesh: plane / triangle / edge overflow cells are dropped like the
    reference drops them (status parity per pair)."""
    mesh = synth.make_ball_mesh(2)
    sites = synth.make_spheres(400, stream=5)
    knn, k = synth.knn_site_lists(sites, 120)
    pt, ps = O.tet_sphere_relation(mesh, sites, knn, k)
    ra, sa, _ = O.run_pairs(mesh, sites, knn, k, pt, ps, impl="oracle")
    res = run_given(ctx, mesh, sites, knn, k)
    assert res.n_pairs == len(pt)
    assert np.array_equal(res.status_histogram, expected_hist(ra, sa))
    assert res.status_histogram[[1, 2, 8]].sum() > 0  # the config really exercises overflow statuses
    assert_defined_equal(O, ra[ra["status"] == 4], res.records())


def test_tet_range_shards_concatenate(ctx, O, cfg1):
    """tet shards (multi-GPU partition): the concatenation of shard results == the full result."""
    mesh, sites, knn, k = cfg1
    full = run_given(ctx, mesh, sites, knn, k).records()
    parts = []
    cuts = [0, 5000, 5001, 13000, mesh.n_tet]
    for a, b in zip(cuts[:-1], cuts[1:]):
        ctx.set_tet_range(a, b - a)
        parts.append(ctx.run().records())
    ctx.set_tet_range(0, -1)
    cat = np.concatenate(parts)
    assert len(cat) == len(full)
    assert np.array_equal(np.concatenate([p["id"] for p in parts]),
                          np.concatenate([np.arange(len(p)) for p in parts]))  # ids restart per shard
    for f in full.dtype.names:
        if f not in ("id", "thread_id"):
            assert np.ascontiguousarray(cat[f]).tobytes() == np.ascontiguousarray(full[f]).tobytes(), f


def test_empty_range_and_unselected(ctx, O, cfg1):
    mesh, sites, knn, k = cfg1
    ctx.set_mesh(mesh)
    ctx.upload_sites(sites.site_soa, sites.weights, sites.flags, knn, k)
    ctx.set_tet_range(10, 0)
    r = ctx.run()
    assert r.n_cells == 0 and r.n_pairs == 0 and len(r.records()) == 0
    ctx.set_tet_range(0, -1)
    # no site selected -> no cells (voronoi.cu:165-169)
    r = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, np.zeros(sites.n_site, np.uint32), knn, k)
    assert r.n_cells == 0


def test_partially_selected_sites(ctx, O, cfg1, synth):
    """N+1ring selected, 2-ring clipped against only (rpd_api.cxx:371-373)."""
    mesh, sites, knn, k = cfg1
    flags = (np.arange(sites.n_site) % 3 != 0).astype(np.uint32)
    s2 = synth.Sites(sites.site_soa, sites.weights, flags, sites.radii)
    pt, ps = O.tet_sphere_relation(mesh, s2, knn, k)
    assert flags[ps].all()
    ra, _, _ = O.run_pairs(mesh, s2, knn, k, pt, ps)
    res = run_given(ctx, mesh, s2, knn, k)
    assert res.n_pairs == len(pt)
    assert_defined_equal(O, ra[ra["status"] == 4], res.records())


def test_partial_tets_and_dense_e_adjs(ctx, O, synth):
    """partial-tet call (rpd_api.cxx:254-281): subset idx/f_adjs/f_ids with GLOBAL vertices, and
    the reference's dense e_adjs table instead of the compact 6-per-tet form."""
    mesh = synth.make_ball_mesh(6)
    sites = synth.make_spheres(120)
    knn, k = synth.knn_site_lists(sites, 60)
    sel = np.arange(mesh.n_tet)[::3]
    sub = synth.TetMesh(mesh.vertices, mesh.indices[sel], mesh.v_adjs, mesh.e_adj6[sel], mesh.f_adjs[sel],
                        mesh.f_ids[sel], mesh.n_surf_faces)
    pt, ps = O.tet_sphere_relation(sub, sites, knn, k)
    ra, _, _ = O.run_pairs(sub, sites, knn, k, pt, ps)
    want = ra[ra["status"] == 4]
    ctx.set_tetmesh(sub.vertices, sub.indices, sub.v_adjs, sub.f_adjs, sub.f_ids, e_adjs_dense=mesh.dense_e_adjs())
    res = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, knn, k)
    assert_defined_equal(O, want, res.records())


def test_site_without_neighbours(ctx, O, synth):
    """a site with zero real neighbours relates to every tet and owns every tet whole (voronoi.cu:171-190)."""
    mesh = synth.make_ball_mesh(3)
    sites = synth.make_spheres(1)
    knn = np.full((2, 1), -1, np.int32)
    res = run_given(ctx, mesh, sites, knn, 1)
    assert res.n_cells == mesh.n_tet
    recs = res.records()
    assert (recs["nb_v"] == 4).all() and (recs["nb_p"] == 4).all() and (recs["nb_e"] == 6).all()
    pt, ps = O.tet_sphere_relation(mesh, sites, knn, 1)
    ra, _, _ = O.run_pairs(mesh, sites, knn, 1, pt, ps)
    assert_defined_equal(O, ra, recs)


def test_compact_blob_round_trip(ctx, O, cfg1):
    """the compact result (what NCCL gathers) expands to exactly the fetched records"""
    mesh, sites, knn, k = cfg1
    res = run_given(ctx, mesh, sites, knn, k)
    blob, offs = res.compact()
    assert offs[0] == 0 and offs[-1] == res.compact_bytes and (np.diff(offs) > 0).all()
    recs = res.records()
    w = blob[offs[:-1] // 4 + 2]
    assert np.array_equal(w & 0xff, recs["nb_v"]) and np.array_equal((w >> 8) & 0xff, recs["nb_p"])
    assert np.array_equal(blob[offs[:-1] // 4].astype(np.int32), recs["tet_id"])
    sizes = 4 * (4 + recs["nb_v"].astype(np.int64) + 7 * recs["nb_p"] + (3 * recs["nb_e"].astype(np.int64) + 3) // 4)
    assert np.array_equal(np.diff(offs), sizes)


def test_error_paths(ctx):
    from libmat_b200.rpd import Context, LibMatError
    c = Context(0)
    with pytest.raises(LibMatError, match="mb_set_tetmesh"):
        c.upload_sites(np.zeros(3, np.float32), np.ones(1, np.float32), np.ones(1, np.uint32))
        c.run()
    with pytest.raises(LibMatError):
        c.run(lanes_per_cell=5)
    c.close()


# ------------------------------------------------------------------------------------------------
# grid-kNN mode
# ------------------------------------------------------------------------------------------------
def canon_equal(O, a, b):
    ca, cb = O.canonicalize(a), O.canonicalize(b)
    d = O.defined_equal(ca, cb)
    return d


def grid_vs_given(ctx, O, mesh, sites, knn, k, allow_capped=True, **kw):
    """grid-kNN mode against the reference semantics with a sufficient neighbour list, on the canonical parity
    form of SURVEY 8a.  No numeric allowance: EVERY difference must belong to a flagged class --
      * det-flagged: a conflict determinant of the (tet, site) pair fell under the predicate_generator bound
        (mb_rpd_fetch_flags on the GPU side, the reference's USE_ARITHMETIC_FILTER test on the oracle side);
      * the reference's own degenerate-volume class: a cell whose a12 volume is below 0.1 keeps a valid record
        while gpu_stat flips to no_intersection (convex_cell.cu:1040) -- such a sliver may exist on one side only;
      * cells dropped at the 64-plane / 96-vertex / 152-edge caps (dead entries count, so the clip order matters),
        reported in the status histogram."""
    ns = sites.n_site
    pt, ps = O.tet_sphere_relation(mesh, sites, knn, k)
    ra, sa, _ = O.run_pairs(mesh, sites, knn, k, pt, ps)
    fo = O.flagged_pairs(mesh, sites, knn, k, pt, ps, "ref" if O.ref("rpd_filter") is not None else "oracle")
    ok = ra["status"] == 4
    want = ra[ok]
    want_flag = fo[ok].astype(bool)
    want_sliver = sa[ok] == O.STATUS["no_intersection"]  # record valid, |vol| < 0.1
    ctx.set_mesh(mesh)
    res = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, None, 0, **kw)
    assert res.n_cand_overflow == 0
    got = res.records()
    cf, pf = res.flags(pairs=True)
    gpt, gps, _ = res.pairs()
    assert int(cf.sum()) == res.n_flagged_cells and int(pf.sum()) == res.n_flagged_pairs
    gpu_flagged_pairs = set((gpt.astype(np.int64)[pf.astype(bool)] * ns + gps[pf.astype(bool)]).tolist())
    ref_flagged_pairs = set((pt.astype(np.int64)[fo.astype(bool)] * ns + ps[fo.astype(bool)]).tolist())
    n_capped = int(res.status_histogram[[1, 2, 8]].sum()) if allow_capped else 0
    ka = want["tet_id"].astype(np.int64) * ns + want["voro_id"]
    kb = got["tet_id"].astype(np.int64) * ns + got["voro_id"]
    common = np.intersect1d(ka, kb)
    in_a, in_b = np.isin(ka, common), np.isin(kb, common)
    # cells of the reference that are missing here: flagged / sliver, or dropped at a cap on this side
    unexplained = [int(key) for key, fl, sl in zip(ka[~in_a], want_flag[~in_a], want_sliver[~in_a])
                   if not (fl or sl or int(key) in gpu_flagged_pairs)]
    assert len(unexplained) <= n_capped, (unexplained[:10], n_capped)
    # cells here that the reference does not have: flagged, or a sliver by the exact volume
    vb = O.cell_volumes(got[~in_b])
    extra = [int(key) for key, fl, v in zip(kb[~in_b], cf[~in_b], vb)
             if not (fl or abs(v) < 0.1 or int(key) in ref_flagged_pairs)]
    assert not extra, extra[:10]
    # common cells: canonical records identical except on det-flagged cells
    ca, cb = O.canonicalize(want[in_a]), O.canonicalize(got[in_b])
    bad = np.zeros(len(ca), bool)
    for f in ("nb_v", "nb_p", "nb_e"):
        bad |= ca[f] != cb[f]
    same = ~bad
    iv = np.arange(96)[None, :] < ca["nb_v"][:, None]
    ip = np.arange(64)[None, :] < ca["nb_p"][:, None]
    ie = np.arange(152)[None, :] < ca["nb_e"][:, None]
    bad |= same & ((ca["ver"] != cb["ver"]).any(axis=2) & iv).any(axis=1)
    bad |= same & ((ca["id2"] != cb["id2"]).any(axis=2) & ip).any(axis=1)
    bad |= same & ((ca["clip"][..., :5].view(np.uint32) != cb["clip"][..., :5].view(np.uint32)).any(axis=2) & ip).any(axis=1)
    bad |= same & ((ca["edge"] != cb["edge"]).any(axis=2) & ie).any(axis=1)
    flagged_common = want_flag[in_a] | cf[in_b].astype(bool)
    assert not (bad & ~flagged_common).any(), (int((bad & ~flagged_common).sum()), int(bad.sum()))
    grid_vs_given.last = {"cells": len(common), "mismatching_flagged": int(bad.sum()), "one_sided": int((~in_a).sum() + (~in_b).sum()),
                          "gpu_flagged_cells": int(cf.sum()), "ref_flagged_cells": int(want_flag.sum())}
    return want, got


@pytest.mark.parametrize("n,ns", [(4, 40), (8, 150)])
def test_grid_mode_canonical_parity_all_neighbours(ctx, O, synth, n, ns):
    """grid-kNN mode vs the reference semantics given ALL other sites as neighbours: same cells,
    same canonical combinatorics (SURVEY 8a parity form)."""
    mesh = synth.make_ball_mesh(n)
    sites = synth.make_spheres(ns)
    knn, k = synth.site_lists_from_sets([[m for m in range(ns) if m != s] for s in range(ns)], ns)
    grid_vs_given(ctx, O, mesh, sites, knn, k)


def test_grid_mode_canonical_parity_cfg1_rt(ctx, O, cfg1_rt):
    """config 1 with regular-triangulation lists: the library's own neighbour search must give the
    cells the reference computes from the RT neighbours."""
    mesh, sites, knn, k = cfg1_rt
    want, got = grid_vs_given(ctx, O, mesh, sites, knn, k)
    assert len(got) > 3 * mesh.n_tet


def test_grid_mode_overflow_pass(ctx, O, synth):
    """many more sites than tets: per-tet survivor lists overflow the fast pass and are redone by the
    big-list pass; results must not change."""
    mesh = synth.make_ball_mesh(3)
    used_big_pass = False
    log = []
    for ns in (500, 700, 900):
        sites = synth.make_spheres(ns, stream=7)
        knn, k, valid = synth.rt_site_lists(sites)
        sites.flags[:] = valid.astype(np.uint32)
        ctx.set_mesh(mesh)
        res = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, None, 0, grid_k=32)
        log.append((ns, res.n_cand_overflow, res.n_big_pass_tets, res.n_cells, str(res.status_histogram.tolist())))
        if res.n_cand_overflow:
            continue  # more than 96 true candidates per tet: that is the documented truncation
        grid_vs_given(ctx, O, mesh, sites, knn, k, grid_k=32)
        used_big_pass |= res.n_big_pass_tets > 0
    assert used_big_pass, log


def test_grid_mode_tiles_every_tet(ctx, O, cfg1_rt):
    """property: the cells of each tet tile it (volume conservation), no dropped cells."""
    mesh, sites, _, _ = cfg1_rt
    ctx.set_mesh(mesh)
    res = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, None, 0)
    recs = res.records()
    cv = O.cell_volumes(recs)
    pv = np.zeros(mesh.n_tet)
    np.add.at(pv, recs["tet_id"], cv)
    tv = mesh.tet_volumes()
    rel = np.abs(pv - tv) / tv
    assert np.mean(rel) < 1e-4 and np.max(rel) < 0.1  # float planes of sliver tets dominate the max
    assert abs(pv.sum() - tv.sum()) / tv.sum() < 1e-6
    assert res.status_histogram[[1, 2, 3, 8, 9]].sum() == 0  # no overflow / inconsistent cells


def test_given_lists_with_grid_candidates(ctx, O, cfg1_rt):
    """opts.grid_candidates: pairs from the uniform-grid search, clipping by the given (RT) lists in list
    order -> the valid cells are byte-identical to the reference semantics; only empty pairs differ."""
    mesh, sites, knn, k = cfg1_rt
    ctx.set_mesh(mesh)
    ref = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, knn, k)
    fast = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, knn, k, grid_candidates=True)
    a, b = ref.records(), fast.records()
    assert fast.n_pairs <= ref.n_pairs and fast.n_cells == ref.n_cells
    for f in a.dtype.names:
        assert np.ascontiguousarray(a[f]).tobytes() == np.ascontiguousarray(b[f]).tobytes(), f
    pt, ps = O.tet_sphere_relation(mesh, sites, knn, k)
    ra, _, _ = O.run_pairs(mesh, sites, knn, k, pt, ps)
    assert_defined_equal(O, ra[ra["status"] == 4], b)


def test_site_volumes_and_barycentres(ctx, O, cfg1_rt):
    """a12 (atomic_add_bary_and_volume): per-site sums agree with the oracle's within float-atomic
    reordering; the volumes tile the mesh."""
    mesh, sites, knn, k = cfg1_rt
    pt, ps = O.tet_sphere_relation(mesh, sites, knn, k)
    _, _, _, vol_o, bary_o = O.run_pairs(mesh, sites, knn, k, pt, ps, n_threads=1, want_vol=True)
    ctx.set_mesh(mesh)
    ctx.upload_sites(sites.site_soa, sites.weights, sites.flags, knn, k)
    res = ctx.run(want_volumes=True)
    vol, bary = res.site_volumes()
    assert np.allclose(vol, vol_o, rtol=2e-5, atol=1e-2)
    assert np.allclose(bary, bary_o, rtol=2e-5, atol=20.0)
    assert abs(float(vol.astype(np.float64).sum()) - mesh.tet_volumes().sum()) / mesh.tet_volumes().sum() < 2e-3  # the FP32 formula of a12 itself
    cv = res.cell_volumes()
    assert len(cv) == res.n_cells and np.isfinite(cv).all()
    per_site = np.zeros(sites.n_site)
    recs = res.records()
    np.add.at(per_site, recs["voro_id"][np.abs(cv) >= 0.1], cv[np.abs(cv) >= 0.1].astype(np.float64))
    assert np.allclose(per_site, vol, rtol=1e-4, atol=1e-1)


@pytest.mark.parametrize("mode", ["given", "grid"])
def test_tet_subset_equals_full_restricted(ctx, O, cfg1_rt, mode):
    """mb_set_tet_subset (device-side partial-tet recompute): the cells of the listed tets equal the
    full run's cells of those tets, with global tet ids and (tet, site) order."""
    mesh, sites, knn, k = cfg1_rt
    ctx.set_mesh(mesh)
    args = (sites.site_soa, sites.weights, sites.flags) + ((knn, k) if mode == "given" else (None, 0))
    full = ctx.compute_clipped_voro_diagram(*args).records()
    rng = np.random.default_rng(5)
    sel = np.sort(rng.choice(mesh.n_tet, size=3000, replace=False)).astype(np.int32)
    ctx.set_tet_subset(sel)
    part = ctx.compute_clipped_voro_diagram(*args).records()
    ctx.set_tet_subset(None)
    want = full[np.isin(full["tet_id"], sel)]
    assert len(part) == len(want) > 0
    for f in full.dtype.names:
        if f not in ("id", "thread_id"):
            assert np.ascontiguousarray(part[f]).tobytes() == np.ascontiguousarray(want[f]).tobytes(), f
    assert np.array_equal(part["id"], np.arange(len(part)))
    from libmat_b200.rpd import LibMatError
    with pytest.raises(LibMatError):
        ctx.set_tet_subset(np.array([5, 3], np.int32))
    again = ctx.compute_clipped_voro_diagram(*args)
    assert again.n_cells == len(full)


def test_security_radius_exit(ctx, O, synth):
    """a9 (opt-in): distance-sorted kNN lists with the security-radius early exit the live reference comments out
    (is_security_radius_reached convex_cell.cu:240-268, used at :1285-1296 and :1304-1316): same cells, byte for byte, and
    the same per-pair statuses -- security_radius_not_reached included -- as the oracle port with the option on; far
    fewer clips than walking the whole list"""
    mesh = synth.make_ball_mesh(8)
    sites = synth.make_spheres(400)
    knn, k = synth.knn_site_lists(sites, 60, by_distance=True)
    pt, ps = O.tet_sphere_relation(mesh, sites, knn, k)
    O.set_security_radius(True)
    try:
        ra, sa, _ = O.run_pairs(mesh, sites, knn, k, pt, ps)
    finally:
        O.set_security_radius(False)
    rb, sb, _ = O.run_pairs(mesh, sites, knn, k, pt, ps)
    ctx.set_mesh(mesh)
    ctx.upload_sites(sites.site_soa, sites.weights, sites.flags, knn, k)
    res = ctx.run(security_radius=True)
    assert res.n_pairs == len(pt)
    assert_defined_equal(O, ra[ra["status"] == 4], res.records())
    assert np.array_equal(res.status_histogram, expected_hist(ra, sa))
    plain = ctx.run()
    # same surviving cells as the plain walk of the whole list (the exit only skips planes that cannot cut); the clip
    # COUNT is not comparable: the bisector cull (a library shortcut) is off in this mode, as in the reference
    # (histogram index = status + 1: [4] = security_radius_not_reached, [5] = success)
    assert res.status_histogram[4] > 0 and plain.status_histogram[4] == 0
    assert plain.status_histogram[5] >= res.status_histogram[5] > 0
    # a list cut so short that the radius cannot be reached: cells are dropped with the reference's status
    knn2, k2 = synth.knn_site_lists(sites, 3, by_distance=True)
    pt2, ps2 = O.tet_sphere_relation(mesh, sites, knn2, k2)
    O.set_security_radius(True)
    try:
        r2, s2, _ = O.run_pairs(mesh, sites, knn2, k2, pt2, ps2)
    finally:
        O.set_security_radius(False)
    ctx.upload_sites(sites.site_soa, sites.weights, sites.flags, knn2, k2)
    res2 = ctx.run(security_radius=True)
    assert np.array_equal(res2.status_histogram, expected_hist(r2, s2))
    assert res2.status_histogram[4] > 0  # security_radius_not_reached (status 3) occurs
    assert_defined_equal(O, r2[r2["status"] == 4], res2.records())
