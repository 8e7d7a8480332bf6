"""K4 emission (get_all_voro_info keys) against the oracle restatement."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_emit_cfg1(ctx, O, cfg1):
    mesh, sites, knn, k = cfg1
    ctx.set_mesh(mesh)
    res = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, knn, k)
    recs = res.records()
    em = res.emit(mesh.n_surf_faces - 1)
    want = O.emit(recs, mesh.n_surf_faces - 1)
    for key in ("facet_cell", "facet_key", "facet_is_tet", "vert_cell", "vert_lvid", "vert_key", "vert_surf_fid",
                "edge_cell", "edge_key", "edge_lvid"):
        assert np.array_equal(em[key], want[key]), key
    # positions: same float expression (compute_vertex_coordinates) -> bitwise; contract is 1e-6 rel
    assert np.array_equal(em["vert_pos"].view(np.uint32), want["vert_pos"].view(np.uint32))
    _, _, eu = O.reload_active(recs, "oracle")
    assert np.array_equal(em["cell_euler"].view(np.uint32), eu.view(np.uint32))
    # pc_face centroids of the surface facets (cell_to_surfv2fid): same loop order, same float sums -> bitwise
    assert np.array_equal(em["facet_centroid"].view(np.uint32), want["facet_centroid"].view(np.uint32))
    assert (em["facet_centroid"] != 0).any()


def test_emit_against_the_reference_update_build(ctx, O, synth, cfg1_rt):
    """K4 against the reference's OWN get_all_voro_info (rpd_update.cxx compiled in place, oracle/_ref/libref_update.so):
    facets, surface centroids, vertices, edges, Euler values, covered feature edges and sharp-line end positions --
    and K6 against its update_pc_cc_info / update_pc_facet_cc_info / update_pc_edge_cc_info."""
    if O.ref("update") is None:
        pytest.skip("oracle/_ref/libref_update.so not built")
    mesh, sites, knn, k = cfg1_rt
    fe_map = synth.fake_feature_edges(mesh, every=5)
    ctx.set_mesh(mesh)
    ctx.set_feature_edges(fe_map)
    res = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, knn, k)
    recs = res.records()
    max_sf = mesh.n_surf_faces - 1
    em = res.emit(max_sf)
    tp = res.topology()
    ctx.set_feature_edges(None)
    R = O.ref_update(recs, sites.n_site, max_sf, fe_map)
    site_of = recs["voro_id"].astype(np.int64)

    def rows(a):
        a = np.asarray(a, dtype=np.int64).reshape(len(a), -1)
        return a[np.lexsort(a.T[::-1])] if len(a) else a

    hp = em["facet_is_tet"] == 0
    assert np.array_equal(rows(np.stack([site_of[em["facet_cell"][hp]], em["facet_key"][hp], em["facet_cell"][hp]], axis=1)), rows(R["facets"]))
    tf = ~hp
    assert np.array_equal(rows(np.stack([site_of[em["facet_cell"][tf]], em["facet_cell"][tf], em["facet_key"][tf]], axis=1)), rows(R["tfids"]))
    sf = tf & (em["facet_key"] <= max_sf)
    got = np.stack([site_of[em["facet_cell"][sf]], em["facet_cell"][sf], em["facet_key"][sf]], axis=1).astype(np.int64)
    want = R["surf"].astype(np.int64)
    og, ow = np.lexsort(got.T[::-1]), np.lexsort(want.T[::-1])
    assert np.array_equal(got[og], want[ow]) and len(want) > 0
    assert np.array_equal(em["facet_centroid"][sf][og].astype(np.float64), R["surf_pos"][ow])
    got = np.concatenate([site_of[em["vert_cell"]][:, None], em["vert_cell"][:, None], em["vert_lvid"][:, None], em["vert_key"],
                          em["vert_surf_fid"][:, None]], axis=1).astype(np.int64)
    want = R["vertices"].astype(np.int64)
    og, ow = np.lexsort(got.T[::-1]), np.lexsort(want.T[::-1])
    assert np.array_equal(got[og], want[ow])
    assert np.array_equal(em["vert_pos"][og].astype(np.float64), R["vertices_pos"][ow])
    assert np.array_equal(rows(np.concatenate([site_of[em["edge_cell"]][:, None], em["edge_key"], em["edge_cell"][:, None], em["edge_lvid"]], axis=1)), rows(R["edges"]))
    assert np.array_equal(em["cell_euler"].view(np.uint32), R["cell_euler"].view(np.uint32))
    # covered feature edges: (site, kind, cell, lv1, lv2, line, fe_id)
    h = em["fe_hit"]
    assert len(h) > 0
    assert np.array_equal(rows(np.concatenate([site_of[h[:, 0]][:, None], h[:, 1:2], h[:, 0:1], h[:, 2:6]], axis=1)), rows(R["fe"]))
    en, ep = em["fe_end"], em["fe_end_pos"]
    keep = en[:, 2] != -1
    got = np.concatenate([site_of[en[keep, 0]][:, None], en[keep]], axis=1).astype(np.int64)
    got, first = np.unique(got, axis=0, return_index=True)
    want = R["fe_end"].astype(np.int64)
    ow = np.lexsort(want.T[::-1])
    assert np.array_equal(got, want[ow])
    assert np.array_equal(ep[keep][first].astype(np.float64), R["fe_end_pos"][ow])
    # K6 partitions against the reference's component lists
    lab = {}
    for s, kcomp, c in R["cc"].astype(np.int64):
        lab.setdefault((s, kcomp), []).append(c)
    ref_label = np.full(len(recs), -1, np.int64)
    for cells in lab.values():
        ref_label[cells] = min(cells)
    assert np.array_equal(ref_label, tp["cell_cc"])
    n_cc = {}
    for s, n, kcomp, c in R["facet_cc"].astype(np.int64):
        n_cc.setdefault((int(s), int(n)), set()).add(int(kcomp))
    got_pairs = {(int(s), int(n)): int(c) for s, n, c in zip(tp["pair_site"], tp["pair_neigh"], tp["pair_n_cc"])}
    assert got_pairs == {k_: len(v) for k_, v in n_cc.items()}
    comp = {}
    for s, a, b, kcomp, c in R["edge_cc"].astype(np.int64):
        comp.setdefault((s, a, b, kcomp), []).append(c)
    first_edge = {}
    for e in range(len(em["edge_cell"])):
        c = int(em["edge_cell"][e])
        first_edge.setdefault((int(site_of[c]), int(em["edge_key"][e][0]), int(em["edge_key"][e][1]), c), e)
    for (s, a, b, kcomp), cells in comp.items():
        es = [first_edge[(s, a, b, c)] for c in cells]
        assert all(tp["edge_cc"][e] == min(es) for e in es)
    res.free()
