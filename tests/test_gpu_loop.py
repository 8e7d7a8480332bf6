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
