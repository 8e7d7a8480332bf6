"""f3: dist2mat candidate lists built on the device (mb_dist2mat_set_medial_mesh / _set_face_sites / _by_face) against
the host restatement of the reference's list construction (gather_point_to_sites + gather_point_to_slab_and_cone,
fix_geo_error.cxx:149-215) and against mb_dist2mat fed with the replicated per-sample copies of those lists
(load_and_compute_sample_dist2mat_gpubuffer, :300-366)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_device_built_lists_equal_reference_construction(ctx, synth):
    d = synth.make_dist2mat(50000)
    off, prims = synth.reference_face_lists(len(d.spheres), d.mm_faces, d.mm_edges, d.fid_sites, d.n_fid)
    ctx.dist2mat_set_medial_mesh(d.spheres, d.mm_faces, d.mm_edges)
    rng = np.random.default_rng(3)
    rows = np.concatenate([d.fid_sites, d.fid_sites[rng.integers(0, len(d.fid_sites), 5000)]])  # duplicates, any order
    ctx.dist2mat_set_face_sites(rows[rng.permutation(len(rows))], d.n_fid)
    goff, gprims = ctx.dist2mat_face_lists()
    assert np.array_equal(goff, off) and np.array_equal(gprims, prims)
    # samples by face id == the reference interface fed with one private copy of the list per sample
    r, cid, tie, prim3 = ctx.dist2mat_by_face(d.samples, d.sample_fid)
    o, c, p = synth.replicate_lists(off, prims, d.sample_fid)
    r2, cid2, tie2 = ctx.compute_closest_dist2mat(d.spheres, d.samples, o, c, p)
    assert np.array_equal(r.view(np.uint32), r2.view(np.uint32)) and np.array_equal(cid, cid2) and np.array_equal(tie, tie2)
    assert (cid >= 0).all() and np.array_equal(prim3, p[o.astype(np.int64) + cid])
    # a face no power cell touches (fix_geo_error.cxx:168-170) and an out-of-range id: empty list
    fid = d.sample_fid[:4].copy()
    ctx.dist2mat_set_face_sites(d.fid_sites[d.fid_sites[:, 0] != fid[0]], d.n_fid)
    fid[1] = -1
    r, cid, _, prim3 = ctx.dist2mat_by_face(d.samples[:4], fid, want_tie=False)
    assert r[0] == np.float32(1e16) and cid[0] == -1 and (prim3[0] == -1).all() and cid[1] == -1 and cid[2] >= 0


def test_lists_from_rpd_surface_facets(ctx, O, synth, cfg1_rt):
    """RPD -> K4 -> dist2mat lists without leaving the device: fid2sites comes from the emitted surface facets
    (cell_to_surfv2fid), the medial mesh is a synthetic one over the RPD's sites"""
    mesh, sites, knn, k = cfg1_rt
    ns = sites.n_site
    spheres = np.concatenate([sites.centers(), sites.radii[:, None]], axis=1).astype(np.float32)
    nb0, nb1 = knn[0], knn[1]
    has2 = (nb0 >= 0) & (nb1 >= 0)
    mm_edges = np.stack([np.arange(ns)[nb0 >= 0], nb0[nb0 >= 0]], axis=1).astype(np.int32)
    mm_faces = np.stack([np.arange(ns)[has2], nb0[has2], nb1[has2]], axis=1).astype(np.int32)
    ctx.set_mesh(mesh)
    res = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, None, 0)
    max_sf = mesh.n_surf_faces - 1
    em = res.emit(max_sf)
    recs = res.records()
    ctx.dist2mat_set_medial_mesh(spheres, mm_faces, mm_edges)
    ctx.dist2mat_set_face_sites_from_rpd(res, max_sf)
    goff, gprims = ctx.dist2mat_face_lists()
    sf = (em["facet_is_tet"] == 1) & (em["facet_key"] <= max_sf)
    rows = np.stack([em["facet_key"][sf], recs["voro_id"][em["facet_cell"][sf]]], axis=1)
    off, prims = synth.reference_face_lists(ns, mm_faces, mm_edges, rows, max_sf + 1)
    assert np.array_equal(goff, off) and np.array_equal(gprims, prims) and len(prims) > 0
    res.free()
