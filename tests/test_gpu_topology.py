"""K6 (mb_rpd_topology): cell / half-plane-facet connected components and Euler sums per power cell against the
literal dict / set / BFS restatement of the reference (oracle.topology: rpd_update.cxx:303-316, 439-470, 497-503;
common_cxx.h:447-489; fix_topo.cxx:117-131)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def check(ctx, O, mesh, sites, knn, k):
    ctx.set_mesh(mesh)
    res = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, knn, k)
    recs = res.records()
    em = res.emit(mesh.n_surf_faces - 1)
    tp = res.topology()
    want = O.topology(em, recs["voro_id"], em["cell_euler"])
    assert np.array_equal(tp["cell_cc"], want["cell_cc"])
    # duplicate half-plane facets of one cell (same neighbour twice) cannot occur for a valid record
    assert np.array_equal(tp["facet_cc"], want["facet_cc"])
    assert len(tp["edge_cc"]) == len(em["edge_cell"]) and np.array_equal(tp["edge_cc"], want["edge_cc"])
    n_site = sites.n_site
    nc = np.zeros(n_site, np.int64)
    ncc = np.zeros(n_site, np.int64)
    es = np.zeros(n_site, np.float64)
    for s, (a, b, c) in want["site_stats"].items():
        nc[s], ncc[s], es[s] = a, b, c
    assert np.array_equal(tp["site_n_cells"], nc)
    assert np.array_equal(tp["site_n_cc"], ncc)
    assert np.array_equal(tp["site_euler_sum"], es)  # same summation order: bit-identical doubles
    got_pairs = {(int(s), int(n)): int(c) for s, n, c in zip(tp["pair_site"], tp["pair_neigh"], tp["pair_n_cc"])}
    assert got_pairs == want["pairs"]
    key = tp["pair_site"].astype(np.int64) * (1 << 32) + tp["pair_neigh"]
    assert (np.diff(key) > 0).all()
    res.free()
    return tp, want


def test_topology_rt_lists(ctx, O, cfg1_rt):
    """sufficient (regular-triangulation) neighbour lists: the power cells tile the mesh; every visible power cell of
    this convex-ish configuration is one component with Euler characteristic 1"""
    mesh, sites, knn, k = cfg1_rt
    tp, want = check(ctx, O, mesh, sites, knn, k)
    vis = tp["site_n_cells"] > 0
    euler = tp["site_euler_sum"] - tp["site_n_cells"]
    assert (tp["site_n_cc"][vis] >= 1).all()
    assert np.mean(np.abs(euler[vis] - 1.0) < 1e-3) > 0.95


def test_topology_insufficient_lists(ctx, O, cfg1):
    """k-nearest-centre lists with many hidden sites: overlapping, fragmented power cells -- components > 1 occur
    and must still match the BFS restatement exactly"""
    mesh, sites, knn, k = cfg1
    tp, want = check(ctx, O, mesh, sites, knn, k)
    assert tp["site_n_cc"].max() >= 1


def test_topology_needs_emit(ctx, cfg1_rt):
    from libmat_b200.capi import LibMatError
    mesh, sites, knn, k = cfg1_rt
    ctx.set_mesh(mesh)
    res = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, None, 0)
    with pytest.raises(LibMatError):
        res.topology()
    res.free()
