"""The plain-C restatement against the reference's own sources compiled in place (oracle/_ref),
on seeded synthetic inputs.  Skipped when oracle/_ref has not been built."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def need_ref(O):
    if O.ref("rpd") is None or O.ref("host") is None or O.ref("d2m") is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")


def test_rpd_records_cfg1(O, cfg1, cfg1_oracle, need_ref):
    mesh, sites, knn, k = cfg1
    pt, ps, ra, sa = cfg1_oracle
    rb, sb, _ = O.run_pairs(mesh, sites, knn, k, pt, ps, impl="ref")
    assert np.array_equal(sa, sb)
    d = O.defined_equal(ra, rb)
    assert all(v == 0 for f, v in d.items() if f != "cells_compared"), d
    assert d["cells_compared"] == int((ra["status"] == 4).sum()) > 40000


def test_rpd_host_postprocessing(O, cfg1_oracle, need_ref):
    _, _, ra, _ = cfg1_oracle
    ok = ra[ra["status"] == 4][:5000]
    a = O.reload_active(ok, "oracle")
    b = O.reload_active(ok, "ref")
    for x, y in zip(a[:2], b[:2]):
        assert np.array_equal(x, y)
    assert np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))
    va = O.vertex_coordinates(ok, "oracle")
    vb = O.vertex_coordinates(ok, "ref")
    assert np.array_equal(va[..., :3].view(np.uint32), vb[..., :3].view(np.uint32))


def test_rpd_volumes_and_overflow_classes(O, synth, need_ref):
    """dense site set on a tiny mesh: exercises plane/vertex/edge overflow statuses."""
    mesh = synth.make_ball_mesh(2)
    sites = synth.make_spheres(400, stream=5)
    n = sites.n_site
    knn, k = synth.knn_site_lists(sites, 120)
    pt, ps = O.tet_sphere_relation(mesh, sites, knn, k)
    ra, sa, _, va, ba = O.run_pairs(mesh, sites, knn, k, pt, ps, impl="oracle", n_threads=1, want_vol=True)
    rb, sb, _, vb, bb = O.run_pairs(mesh, sites, knn, k, pt, ps, impl="ref", n_threads=1, want_vol=True)
    assert np.array_equal(sa, sb)
    d = O.defined_equal(ra, rb)
    assert all(v == 0 for f, v in d.items() if f != "cells_compared"), d
    assert np.array_equal(va[..., :3].view(np.uint32), vb[..., :3].view(np.uint32))
    assert np.array_equal(ba.view(np.uint32), bb.view(np.uint32))


def test_dist2mat(O, synth, need_ref):
    d = synth.make_dist2mat(20000)
    ra, ia, _, sec = O.dist2mat(d, "oracle", want_second=True)
    rb, ib, _ = O.dist2mat(d, "ref")
    rel = np.abs(ra - rb) / np.maximum(np.abs(rb), 1e-3)
    assert rel.max() <= 1e-6
    tie = (sec - ra) <= 1e-6 * np.maximum(np.abs(ra), np.abs(sec))
    assert not np.any((ia != ib) & ~tie)
