"""The C++ drop-in shims (include/libmat_b200_shim.hpp, libmat_b200_dist2mat_shim.hpp): a C++ host
program with the reference's exact signatures and the reference's own ConvexCellHost / GpuBuffer types
(tests/cxx/shim_driver.cpp, built against /root/reference's headers in place)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import ROOT

DRIVER = os.path.join(ROOT, "tests", "cxx", "_build", "shim_driver")


def _w(f, a):
    a = np.ascontiguousarray(a)
    f.write(np.int64(a.size).tobytes())
    f.write(a.tobytes())


def _r(f, dt):
    n = int(np.frombuffer(f.read(8), np.int64)[0])
    return np.frombuffer(f.read(n * np.dtype(dt).itemsize), dt).copy()


def test_shim_builds_against_reference_headers():
    if not os.path.isdir("/root/reference/src/rpd3d_base"):
        pytest.skip("reference tree absent")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cxx")])
    assert os.path.exists(DRIVER)


@pytest.mark.gpu
def test_shim_rpd_matches_oracle(O, synth):
    if not os.path.exists(DRIVER):
        pytest.skip("tests/cxx/_build/shim_driver not prebuilt")
    mesh = synth.make_ball_mesh(6)
    sites = synth.make_spheres(120)
    knn, k = synth.knn_site_lists(sites, 60)
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        with open(fin, "wb") as f:
            for a in (mesh.vertices.astype(np.float32), mesh.indices.astype(np.int32), mesh.v_adjs.astype(np.int32),
                      mesh.dense_e_adjs(), mesh.f_adjs.astype(np.int32), mesh.f_ids.astype(np.int32),
                      sites.site_soa, sites.weights, sites.flags, knn.astype(np.int32),
                      np.array([sites.n_site, k], np.int32)):
                _w(f, a)
        r = subprocess.run([DRIVER, "rpd", fin, fout], capture_output=True, text=True)
        assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
        with open(fout, "rb") as f:
            hdr = _r(f, np.int32).reshape(-1, 8)
            ver = _r(f, np.uint8).reshape(-1, 96, 4)
            clip = _r(f, np.float32).reshape(-1, 64, 5)
            id2 = _r(f, np.int32).reshape(-1, 64, 2)
            edge = _r(f, np.uint8).reshape(-1, 152, 3)
            euler = _r(f, np.float32)
            vol = _r(f, np.float32)
    pt, ps = O.tet_sphere_relation(mesh, sites, knn, k)
    ra, _, _ = O.run_pairs(mesh, sites, knn, k, pt, ps)
    want = O.zero_undefined(ra[ra["status"] == 4])
    assert len(hdr) == len(want) > 1000
    assert np.array_equal(hdr[:, 0], want["status"]) and np.array_equal(hdr[:, 1], want["voro_id"])
    assert np.array_equal(hdr[:, 2], want["tet_id"]) and np.array_equal(hdr[:, 3], np.arange(len(want)))
    assert np.array_equal(hdr[:, 4], want["nb_v"]) and np.array_equal(hdr[:, 5], want["nb_p"])
    assert np.array_equal(hdr[:, 6], want["nb_e"]) and (hdr[:, 7] == 1).all()
    assert np.array_equal(ver, want["ver"]) and np.array_equal(edge, want["edge"]) and np.array_equal(id2, want["id2"])
    assert np.array_equal(clip.view(np.uint32), want["clip"][..., :5].view(np.uint32))
    _, _, eu = O.reload_active(want, "oracle")
    assert np.array_equal(euler.view(np.uint32), eu.view(np.uint32))  # the reference's own cal_cell_euler
    assert len(vol) == sites.n_site and (vol == 0).all()              # voronoi.cu:501-502: never filled


@pytest.mark.gpu
def test_shim_dist2mat_matches_oracle(O, synth):
    if not os.path.exists(DRIVER):
        pytest.skip("tests/cxx/_build/shim_driver not prebuilt")
    d = synth.make_dist2mat(20000, nu=20, nv=40, n_slabs=2400, n_cones=1200)
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        with open(fin, "wb") as f:
            for a in (d.spheres, d.samples, d.offset, d.count, d.prims):
                _w(f, a)
        subprocess.check_call([DRIVER, "d2m", fin, fout])
        with open(fout, "rb") as f:
            res = _r(f, np.float32)
            cid = _r(f, np.int32)
    # the shim returns no tie flags: take them from the oracle's runner-up distance (scale-aware, as the kernel's)
    ro, co, _, sec = O.dist2mat(d, "oracle", want_second=True)
    tie = ((sec - ro) <= 1e-6 * np.maximum(np.maximum(np.abs(ro), np.abs(sec)), 0.1)).astype(np.uint8)
    from test_gpu_dist2mat import check_against_builds
    check_against_builds(O, d, res, cid, tie, min_bitwise=0.98)
