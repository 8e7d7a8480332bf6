"""config-5 style loop: successive recomputes with sphere insertion / update on a resident mesh."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_loop_iterations_equal_fresh_runs(O, synth):
    from libmat_b200.loop import RpdLoop, evolve_sites
    from libmat_b200.rpd import Context
    mesh = synth.make_ball_mesh(10)
    sites = synth.make_spheres(600)
    a, b = Context(0), Context(0)
    loop = RpdLoop(a, mesh)
    n_prev = sites.n_site
    for it in range(4):
        sites, changed = evolve_sites(sites, it)
        assert sites.n_site > n_prev and len(changed) >= 2
        n_prev = sites.n_site
        res, dt = loop.step(sites)
        b.set_mesh(mesh)
        fresh = b.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, None, 0)
        ra, rb = res.records(), fresh.records()
        assert len(ra) == len(rb)
        for f in ra.dtype.names:
            assert np.ascontiguousarray(ra[f]).tobytes() == np.ascontiguousarray(rb[f]).tobytes(), (it, f)
        # and the cells still tile the mesh
        cv = O.cell_volumes(ra)
        assert abs(cv.sum() - mesh.tet_volumes().sum()) / mesh.tet_volumes().sum() < 1e-6
    a.close()
    b.close()


def test_loop_partial_subset_matches_full(O, synth):
    """a caller-supplied affected-tet subset reproduces the full run's cells of those tets"""
    from libmat_b200.loop import RpdLoop, evolve_sites
    from libmat_b200.rpd import Context
    mesh = synth.make_ball_mesh(8)
    sites = synth.make_spheres(300)
    c = Context(0)
    loop = RpdLoop(c, mesh)
    full, _ = loop.step(sites)
    rf = full.records()
    sites2, changed = evolve_sites(sites, 0)
    full2, _ = loop.step(sites2)
    r2 = full2.records()
    # affected tets = tets whose cell set differs between the two full runs (ground truth)
    def keyset(r):
        return set(zip(r["tet_id"].tolist(), r["voro_id"].tolist()))
    diff = keyset(rf) ^ keyset(r2)
    sel = np.unique(np.array(sorted({t for t, _ in diff}), dtype=np.int32))
    assert 0 < len(sel) < mesh.n_tet
    part, _ = loop.step(sites2, tet_subset=sel)
    rp = part.records()
    want = r2[np.isin(r2["tet_id"], sel)]
    assert len(rp) == len(want)
    for f in ("tet_id", "voro_id", "nb_v", "nb_p", "nb_e", "ver", "id2", "edge"):
        assert np.array_equal(rp[f], want[f]), f
    c.close()


def test_loop_streamed_step_equals_device_step(synth):
    """RpdLoop.step(to_host=True) (streamed D2H) delivers the same compact result as the device-resident step"""
    from libmat_b200.loop import RpdLoop, evolve_sites
    from libmat_b200.rpd import Context
    mesh = synth.make_ball_mesh(10)
    sites = synth.make_spheres(600)
    c = Context(0)
    loop = RpdLoop(c, mesh)
    for it in range(2):
        sites, _ = evolve_sites(sites, it)
        res, _ = loop.step(sites)
        blob, offs = res.compact()
        want = (blob[: res.compact_bytes // 4].copy(), offs.copy())
        res2, _ = loop.step(sites, to_host=True, n_chunks=3)
        b2, o2 = res2.host_compact()
        assert np.array_equal(want[0], b2) and np.array_equal(want[1], o2)
    c.close()


def test_incremental_run_equals_full_recompute(O, synth):
    """mb_rpd_run_incremental: the records of the affected tets merged into the previous result (mb_rpd_merge_compact)
    are byte-identical to a full recompute, iteration after iteration; unaffected tets really are untouched; an
    iteration without changes affects nothing"""
    from libmat_b200 import capi
    from libmat_b200.loop import RpdLoop, evolve_sites, site_rings
    from libmat_b200.rpd import Context
    mesh = synth.make_ball_mesh(12)
    sites = synth.make_spheres(1500)
    a, b = Context(0), Context(0)
    loop = RpdLoop(a, mesh)
    b.set_mesh(mesh)
    res, tets, _ = loop.step_incremental(sites)
    assert len(tets) == mesh.n_tet  # first call: everything
    blob, offs = res.compact()
    blob = blob[: res.compact_bytes // 4].copy()
    fractions = []
    for it in range(4):
        prev_sites = sites
        sites, changed = evolve_sites(sites, it, frac_insert=0.003, frac_update=0.003)
        res, tets, _ = loop.step_incremental(sites, to_host=(it % 2 == 1), n_chunks=2)
        fractions.append(len(tets) / mesh.n_tet)
        assert 0 < len(tets) < mesh.n_tet and (np.diff(tets) > 0).all()
        pb, po = res.host_compact() if it % 2 == 1 else res.compact()
        pb = pb[: res.compact_bytes // 4]
        assert np.isin(pb[po[:-1] // 4].astype(np.int64), tets).all()
        blob, offs = capi.merge_compact(blob, offs, pb, po, tets)
        full = b.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, None, 0)
        fb, fo = full.compact()
        assert np.array_equal(offs, fo) and np.array_equal(blob, fb[: full.compact_bytes // 4]), it
        # the tets of the changed spheres' previous / new cells are among the affected ones
        recs = full.records()
        touched = np.unique(recs["tet_id"][np.isin(recs["voro_id"], changed)])
        assert np.isin(touched, tets).all()
        full.free()
    assert max(fractions) < 0.6
    # nothing changed -> nothing affected, empty patch
    res, tets, _ = loop.step_incremental(sites)
    assert len(tets) == 0 and res.n_cells == 0
    # the neighbour rings of the changed spheres from K6's half-plane pairs (the reference's N + 1-ring + 2-ring)
    full = b.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, None, 0)
    full.emit(mesh.n_surf_faces - 1)
    tp = full.topology()
    rings = site_rings(tp["pair_site"], tp["pair_neigh"], changed, sites.n_site)
    assert len(rings) == 3 and len(rings[1]) > 0 and len(rings[2]) > 0
    a.close()
    b.close()
    print("incremental: affected-tet fractions per iteration", [round(f, 3) for f in fractions])
