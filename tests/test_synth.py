"""Synthetic generators: determinism and the mesh invariants the kernels rely on."""
import numpy as np


def test_mesh_sizes_and_orientation(synth):
    m = synth.make_ball_mesh(6)
    assert m.n_tet == 6 * 6**3 and m.n_vert == 7**3
    assert (m.tet_volumes() > 0).all()  # positively oriented (convex_cell.cu:140-156 convention)
    assert m.vertices.min() >= 0 and m.vertices.max() <= 1000  # [0,1000]^3, params.h:15
    assert abs(m.tet_volumes().sum() - 4 / 3 * np.pi * 500**3) / (4 / 3 * np.pi * 500**3) < 0.08


def test_mesh_adjacency(synth):
    m = synth.make_ball_mesh(5)
    assert set(np.unique(m.f_adjs)) == {1, 2}
    # boundary faces are numbered first, every interior face id appears exactly twice
    fid = m.f_ids.ravel()
    fa = m.f_adjs.ravel()
    assert fid[fa == 1].max() == m.n_surf_faces - 1 and len(np.unique(fid[fa == 1])) == m.n_surf_faces
    u, c = np.unique(fid[fa == 2], return_counts=True)
    assert (c == 2).all() and u.min() == m.n_surf_faces
    assert m.v_adjs.sum() == 4 * m.n_tet
    # dense table round trip (reference get_edge_idx, convex_cell.h:46-59)
    dense = m.dense_e_adjs()
    a = m.indices[:, synth.TET_EDGE_PAIRS[:, 0]].astype(np.int64)
    b = m.indices[:, synth.TET_EDGE_PAIRS[:, 1]].astype(np.int64)
    assert np.array_equal(dense[synth.edge_idx(a, b, m.n_vert)], m.e_adj6)
    # Euler characteristic of a ball: V - E + F - T = 1
    ne = len(np.unique(np.minimum(a, b) * m.n_vert + np.maximum(a, b)))
    nf = len(np.unique(fid))
    assert m.n_vert - ne + nf - m.n_tet == 1


def test_determinism(synth):
    a, b = synth.make_ball_mesh(4), synth.make_ball_mesh(4)
    assert np.array_equal(a.vertices, b.vertices) and np.array_equal(a.indices, b.indices)
    s, t = synth.make_spheres(100), synth.make_spheres(100)
    assert np.array_equal(s.site_soa, t.site_soa) and np.array_equal(s.weights, t.weights)
    assert np.allclose(s.weights, s.radii**2)


def test_knn_layout(synth):
    s = synth.make_spheres(50)
    knn, k = synth.knn_site_lists(s, 10)
    assert knn.shape == (k + 1, 50) and (knn[-1] == -1).all()  # triangulation.cxx:245-256
    for j in range(50):
        col = knn[:k, j]
        assert (np.diff(col) > 0).all() and j not in col


def test_dist2mat_input(synth):
    d = synth.make_dist2mat(500, nu=20, nv=40, n_slabs=2400, n_cones=1200)
    assert d.offset[0] == 0 and np.array_equal(d.offset[1:], np.cumsum(d.count)[:-1].astype(np.uint32))
    assert d.prims.shape[0] == int(d.count.sum())
    kinds = np.where(d.prims[:, 1] == -1, 0, np.where(d.prims[:, 0] == -1, 1, 2))
    assert set(np.unique(kinds)) == {0, 1, 2}
    assert d.prims.max() < len(d.spheres)


def test_library_tet_adjacency_equals_generator(synth):
    """mb_tet_adjacency (f4: the sparse load_tet_adj_info, io.cxx:238-335) against the restatement in synth.tet_adjacency,
    on a full mesh and on an arbitrary sub-mesh (partial-tet calls, rpd_api.cxx:254-281)"""
    import numpy as np
    from libmat_b200 import capi
    mesh = synth.make_ball_mesh(9)
    v, e6, fa, fi, nb = capi.tet_adjacency(mesh.indices, mesh.n_vert)
    assert nb == mesh.n_surf_faces
    assert np.array_equal(v, mesh.v_adjs) and np.array_equal(e6, mesh.e_adj6)
    assert np.array_equal(fa, mesh.f_adjs) and np.array_equal(fi, mesh.f_ids)
    sub = mesh.indices[::3]
    want = synth.tet_adjacency(sub, mesh.n_vert)
    got = capi.tet_adjacency(sub, mesh.n_vert)
    for a, b in zip(got[:4], want[:4]):
        assert np.array_equal(a, np.asarray(b).reshape(a.shape))
    assert got[4] == want[4]
    # caller-supplied surface facet ids for the boundary faces, interior ids after n_sf_facets
    perm = np.random.default_rng(0).permutation(nb).astype(np.int32)
    _, _, _, fi2, _ = capi.tet_adjacency(mesh.indices, mesh.n_vert, boundary_sf_fids=perm, n_sf_facets=nb + 7)
    b = mesh.f_adjs == 1
    assert np.array_equal(fi2[b], perm) and np.array_equal(fi2[~b], mesh.f_ids[~b] + 7)
