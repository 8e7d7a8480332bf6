"""K4 / K6 parity is pinned against the reference's OWN power-cell update: src/rpd3d_base/rpd_update.cxx
(get_all_voro_info, update_pc_cc_info, update_pc_facet_cc_info, update_pc_edge_cc_info, update_power_cells) compiled
in place with geogram stand-ins (oracle/_ref/libref_update.so).  This file checks the two restatements the GPU tests
compare against -- orc_emit (C) and oracle.topology (dict / set / BFS) -- item by item against that build."""
import numpy as np
import pytest


def _rows(a):
    """sorted unique rows of an int array, as a set-like comparable"""
    a = np.asarray(a, dtype=np.int64).reshape(len(a), -1)
    return a[np.lexsort(a.T[::-1])] if len(a) else a


@pytest.fixture(scope="module")
def case(O, synth):
    if O.ref("update") is None:
        pytest.skip("oracle/_ref/libref_update.so not built")
    mesh = synth.make_ball_mesh(8)
    sites = synth.make_spheres(150)
    knn, k, valid = synth.rt_site_lists(sites)
    sites.flags[:] = valid.astype(np.uint32)
    pt, ps = O.tet_sphere_relation(mesh, sites, knn, k)
    ra, _, _ = O.run_pairs(mesh, sites, knn, k, pt, ps)
    recs = ra[ra["status"] == 4].copy()
    recs["id"] = np.arange(len(recs))
    max_sf = mesh.n_surf_faces - 1
    fe_map = synth.fake_feature_edges(mesh, every=7)
    return mesh, sites, recs, max_sf, fe_map, O.ref_update(recs, sites.n_site, max_sf, fe_map)


def test_emit_restatement_equals_reference_update(O, case):
    mesh, sites, recs, max_sf, fe_map, R = case
    em = O.emit(recs, max_sf)
    site_of = recs["voro_id"].astype(np.int64)
    # facets: half-planes (site, neigh, cell) and tet faces (site, cell, tfid)
    hp = em["facet_is_tet"] == 0
    got = np.stack([site_of[em["facet_cell"][hp]], em["facet_key"][hp], em["facet_cell"][hp]], axis=1)
    assert np.array_equal(_rows(got), _rows(R["facets"]))
    tf = ~hp
    got = np.stack([site_of[em["facet_cell"][tf]], em["facet_cell"][tf], em["facet_key"][tf]], axis=1)
    assert np.array_equal(_rows(got), _rows(R["tfids"]))
    # surface facets: centroid of the face loop (get_cell_v2surffid, rpd_update.cxx:20-42), bit-for-bit as float
    sf = tf & (em["facet_key"] <= max_sf)
    got = np.stack([site_of[em["facet_cell"][sf]], em["facet_cell"][sf], em["facet_key"][sf]], axis=1).astype(np.int64)
    want = R["surf"].astype(np.int64)
    og, ow = np.lexsort(got.T[::-1]), np.lexsort(want.T[::-1])
    assert np.array_equal(got[og], want[ow]) and len(want) > 0
    assert np.array_equal(em["facet_centroid"][sf][og].astype(np.float64), R["surf_pos"][ow])
    # vertices: key, sorted neighbour triple, surface fid, position
    got = np.concatenate([site_of[em["vert_cell"]][:, None], em["vert_cell"][:, None], em["vert_lvid"][:, None], em["vert_key"],
                          em["vert_surf_fid"][:, None]], axis=1).astype(np.int64)
    want = R["vertices"].astype(np.int64)
    og, ow = np.lexsort(got.T[::-1]), np.lexsort(want.T[::-1])
    assert np.array_equal(got[og], want[ow]) and len(want) > 0
    assert np.array_equal(em["vert_pos"][og].astype(np.float64), R["vertices_pos"][ow])
    # bisector-bisector edges: key + end vertices
    got = np.concatenate([site_of[em["edge_cell"]][:, None], em["edge_key"], em["edge_cell"][:, None], em["edge_lvid"]], axis=1)
    assert np.array_equal(_rows(got), _rows(R["edges"]))
    got = np.concatenate([site_of[em["edge_cell"]][:, None], em["edge_key"], em["edge_cell"][:, None]], axis=1)
    assert np.array_equal(_rows(np.unique(got, axis=0)), _rows(R["e2cells"]))
    # per-cell Euler values (cal_cell_euler after reload_active)
    _, _, eu = O.reload_active(recs, "oracle")
    assert np.array_equal(eu.view(np.uint32), R["cell_euler"].view(np.uint32))
    # feature-edge hits (tet_es2fe_map) and sharp-line end positions
    fe = O.feature_edges(recs, fe_map)
    assert len(R["fe"]) > 0
    assert np.array_equal(_rows(fe["rows"]), _rows(R["fe"]))
    og, ow = np.lexsort(fe["end_rows"].astype(np.int64).T[::-1]), np.lexsort(R["fe_end"].astype(np.int64).T[::-1])
    assert np.array_equal(fe["end_rows"][og], R["fe_end"][ow])
    assert np.array_equal(fe["end_pos"][og].astype(np.float64), R["fe_end_pos"][ow])


def test_topology_restatement_equals_reference_update(O, case):
    mesh, sites, recs, max_sf, fe_map, R = case
    em = O.emit(recs, max_sf)
    _, _, eu = O.reload_active(recs, "oracle")
    tp = O.topology(em, recs["voro_id"], eu)
    site_of = recs["voro_id"].astype(np.int64)
    # cell components: same partition (labels = smallest member vs the reference's component index)
    want = R["cc"].astype(np.int64)
    lab = {}
    for s, kcomp, c in want:
        lab.setdefault((s, kcomp), []).append(c)
    ref_label = np.full(len(recs), -1, np.int64)
    for cells in lab.values():
        ref_label[cells] = min(cells)
    assert np.array_equal(ref_label, tp["cell_cc"])
    # half-plane facet components
    want = R["facet_cc"].astype(np.int64)
    comp = {}
    for s, n, kcomp, c in want:
        comp.setdefault((s, n, kcomp), []).append(c)
    n_cc = {}
    for (s, n, kcomp), cells in comp.items():
        n_cc[(s, n)] = n_cc.get((s, n), 0) + 1
    assert n_cc == {k: v for k, v in tp["pairs"].items()}
    hp = np.flatnonzero(em["facet_is_tet"] == 0)
    first_facet = {}
    for f in hp:
        first_facet.setdefault((int(site_of[em["facet_cell"][f]]), int(em["facet_key"][f]), int(em["facet_cell"][f])), int(f))
    for (s, n, kcomp), cells in comp.items():
        fs = [first_facet[(s, n, c)] for c in cells]
        assert all(tp["facet_cc"][f] == min(fs) for f in fs)
    # edge components
    want = R["edge_cc"].astype(np.int64)
    comp = {}
    for s, a, b, kcomp, c in want:
        comp.setdefault((s, a, b, kcomp), []).append(c)
    first_edge = {}
    for e in range(len(em["edge_cell"])):
        c = int(em["edge_cell"][e])
        first_edge.setdefault((int(site_of[c]), int(em["edge_key"][e][0]), int(em["edge_key"][e][1]), c), e)
    for (s, a, b, kcomp), cells in comp.items():
        es = [first_edge[(s, a, b, c)] for c in cells]
        assert all(tp["edge_cc"][e] == min(es) for e in es)
    # cell neighbours (first / last cell of a tet-face id's set)
    nb = R["neighbours"].astype(np.int64)
    assert len(nb) > 0 and (site_of[nb[:, 1]] == nb[:, 0]).all() and (site_of[nb[:, 2]] == nb[:, 0]).all()
