"""The flagged class of SURVEY 8a: cells with a conflict determinant under the static-filter bound that the
reference's src/predicate_generator prints (KAT-3) and its USE_ARITHMETIC_FILTER branch consumes
(convex_cell.cu:479-497).  Pin: oracle/_ref/libref_rpd_filter.so = the reference's convex_cell.cu compiled for the
host WITH that branch switched on; a pair is flagged iff that build ends it as needs_exact_predicates.

Generic inputs never get near the bound (0 flagged pairs in configs 1 and 2), so the class is exercised with an
exactly degenerate input: an unjittered lattice mesh and equal spheres on a lattice, where bisectors pass through
mesh vertices and edges."""
import os

import numpy as np
import pytest


def lattice(synth, n=8, m=5, L=1024.0):
    mesh = synth.make_box_mesh(n, L)
    sites = synth.make_lattice_spheres(m, L)
    ns = sites.n_site
    knn, k = synth.site_lists_from_sets([[q for q in range(ns) if q != s] for s in range(ns)], ns)
    return mesh, sites, knn, k


def test_oracle_flagged_class_equals_reference_filter_build(O, synth, cfg1, cfg1_oracle, capfd):
    if O.ref("rpd_filter") is None:
        pytest.skip("oracle/_ref/libref_rpd_filter.so not built")
    mesh, sites, knn, k = lattice(synth)
    pt, ps = O.tet_sphere_relation(mesh, sites, knn, k)
    fo = O.flagged_pairs(mesh, sites, knn, k, pt, ps, "oracle")
    fr = O.flagged_pairs(mesh, sites, knn, k, pt, ps, "ref")
    capfd.readouterr()  # the reference prints its needs_perturb diagnostics
    assert np.array_equal(fo, fr)
    assert 0.2 * len(pt) < fr.sum() < len(pt)  # the input really is degenerate
    # a generic input has no flagged pair at all
    mesh, sites, knn, k = cfg1
    pt, ps, _, _ = cfg1_oracle
    assert O.flagged_pairs(mesh, sites, knn, k, pt, ps, "oracle").sum() == 0
    assert O.flagged_pairs(mesh, sites, knn, k, pt, ps, "ref").sum() == 0


@pytest.mark.gpu
def test_gpu_flags_equal_reference_filter_build(O, synth, capfd):
    """given-neighbours mode without the cull filter runs exactly the reference's conflict tests: the per-pair flags
    must EQUAL the reference's USE_ARITHMETIC_FILTER verdicts, and the records stay byte-identical to the live
    (unfiltered) reference on this degenerate input, overflow / perturb / inconsistent statuses included."""
    from libmat_b200.rpd import Context
    mesh, sites, knn, k = lattice(synth)
    pt, ps = O.tet_sphere_relation(mesh, sites, knn, k)
    ra, sa, _ = O.run_pairs(mesh, sites, knn, k, pt, ps, impl="ref" if O.ref("rpd") is not None else "oracle")
    fr = O.flagged_pairs(mesh, sites, knn, k, pt, ps, "ref" if O.ref("rpd_filter") is not None else "oracle")
    capfd.readouterr()
    os.environ["MB_NO_CULL"] = "1"
    try:
        c = Context(0)
    finally:
        del os.environ["MB_NO_CULL"]
    c.set_mesh(mesh)
    res = c.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, knn, k)
    gt, gs, gst = res.pairs()
    assert np.array_equal(gt, pt) and np.array_equal(gs, ps)
    cf, pf = res.flags(pairs=True)
    assert np.array_equal(pf, fr), (int(pf.sum()), int(fr.sum()), int((pf != fr).sum()))
    ok = ra["status"] == 4
    assert np.array_equal(cf, fr[ok]) and res.n_flagged_cells == int(fr[ok].sum()) and res.n_flagged_pairs == int(fr.sum())
    d = O.defined_equal(ra[ok], res.records())
    assert all(v == 0 for f, v in d.items() if f != "cells_compared"), d
    assert np.array_equal(gst.astype(np.int32), np.where(ok, 4, sa))
    c.close()


@pytest.mark.gpu
def test_gpu_flags_with_cull_are_a_subset_and_unflagged_records_unchanged(ctx, O, synth, capfd):
    """the default path culls listed neighbours whose bisector provably misses the tet, i.e. it skips conflict tests
    the reference performs: it can only flag FEWER pairs.  On an exactly degenerate input the skipped tests are not
    always no-ops in the reference (a cell vertex of three planes through one line has w = 0 and a determinant of
    rounding noise against ANY plane), so records may differ -- but only on pairs of the reference's flagged class;
    every other record stays byte-identical.  MB_NO_CULL=1 is the literal behaviour (previous test)."""
    mesh, sites, knn, k = lattice(synth)
    ns = sites.n_site
    pt, ps = O.tet_sphere_relation(mesh, sites, knn, k)
    ra, sa, _ = O.run_pairs(mesh, sites, knn, k, pt, ps)
    fr = O.flagged_pairs(mesh, sites, knn, k, pt, ps, "oracle").astype(bool)
    capfd.readouterr()
    ctx.set_mesh(mesh)
    res = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, knn, k)
    cf, pf = res.flags(pairs=True)
    assert not (pf.astype(bool) & ~fr).any() and pf.sum() > 0
    got = res.records()
    ok = ra["status"] == 4
    want = ra[ok]
    ka = want["tet_id"].astype(np.int64) * ns + want["voro_id"]
    kb = got["tet_id"].astype(np.int64) * ns + got["voro_id"]
    clean = set((pt.astype(np.int64)[~fr] * ns + ps[~fr]).tolist())  # pairs the reference's filter build certifies
    ia = np.array([int(x) in clean for x in ka])
    ib = np.array([int(x) in clean for x in kb])
    assert np.array_equal(ka[ia], kb[ib]) and ia.sum() > 0
    d = O.defined_equal(want[ia], got[ib])
    assert all(v == 0 for f, v in d.items() if f != "cells_compared"), d
    # the streamed run carries the same flags in its records
    rs = ctx.run_to_host(n_chunks=3)
    assert np.array_equal(rs.flags(), cf)
    rs.free()


@pytest.mark.gpu
def test_grid_mode_differences_on_degenerate_input_are_all_flagged(ctx, O, synth, capfd):
    """grid-kNN mode clips in a different order than the reference (per-tet candidate lists): on an exactly
    degenerate input the combinatorics legitimately differ -- and every differing cell must carry the flag."""
    from test_gpu_rpd import grid_vs_given
    mesh, sites, knn, k = lattice(synth)
    grid_vs_given(ctx, O, mesh, sites, knn, k)
    capfd.readouterr()
    info = grid_vs_given.last
    assert info["gpu_flagged_cells"] > 0 and info["ref_flagged_cells"] > 0
    print("degenerate lattice, grid vs reference:", info)
