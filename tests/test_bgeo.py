"""The IO_CUDA result format: libmat_b200's native .bgeo writer (libmat_b200/csrc/bgeo.cu, host code) against the
reference's own save_convex_cells_houdini + IO::GeometryWriter compiled in place (oracle/_ref/libref_bgeo.so; geogram
/ json headers replaced by the stand-ins in oracle/stubs) -- byte for byte, for the 16-bit and the 32-bit point-index
variants and the boundary-only filter.  Runs on the CPU: the writer is host code and needs no device."""
import os
import struct

import numpy as np
import pytest

from libmat_b200 import capi


def oracle_records(O, mesh, sites, knn, k):
    pt, ps = O.tet_sphere_relation(mesh, sites, knn, k)
    ra, _, _ = O.run_pairs(mesh, sites, knn, k, pt, ps, impl="oracle")
    return O.zero_undefined(ra[ra["status"] == 4])


def parse_header(b):
    magic, v, ver, n_pts, n_prims = struct.unpack(">IcIII", b[:17])
    return magic, v, ver, n_pts, n_prims


@pytest.mark.parametrize("boundary_only", [False, True])
def test_bgeo_mini_16bit(O, tmp_path, boundary_only):
    if O.ref("bgeo") is None:
        pytest.skip("oracle/_ref not built")
    from oracle.gen_golden import mini_inputs
    mesh, sites, knn, k = mini_inputs()
    recs = oracle_records(O, mesh, sites, knn, k)
    want = O.ref_bgeo(recs, mesh.n_surf_faces - 1, boundary_only, str(tmp_path / "work"))
    path = str(tmp_path / "ours.bgeo")
    n_pts, n_poly = capi.bgeo_write_records(recs, path, mesh.n_surf_faces - 1, boundary_only)
    got = open(path, "rb").read()
    magic, v, ver, hp, hq = parse_header(got)
    assert magic == 0x4267656F and v == b"V" and ver == 5 and hp == n_pts and hq == n_poly
    assert n_pts == int(recs["nb_v"].sum()) <= 65536  # 16-bit indices
    assert got == want


def test_bgeo_cfg1_32bit(O, synth, tmp_path):
    if O.ref("bgeo") is None:
        pytest.skip("oracle/_ref not built")
    mesh = synth.make_ball_mesh(10)
    sites = synth.make_spheres(400)
    knn, k, valid = synth.rt_site_lists(sites)
    sites.flags[:] = valid.astype(np.uint32)
    recs = oracle_records(O, mesh, sites, knn, k)
    want = O.ref_bgeo(recs, mesh.n_surf_faces - 1, False, str(tmp_path / "work"))
    path = str(tmp_path / "ours.bgeo")
    n_pts, n_poly = capi.bgeo_write_records(recs, path, mesh.n_surf_faces - 1, False)
    assert n_pts > 65536  # 32-bit indices
    got = open(path, "rb").read()
    assert len(got) == len(want)
    assert got == want
    assert got[-2:] == b"\x00\xff"


def test_bgeo_empty_is_an_error(tmp_path):
    with pytest.raises(capi.LibMatError):
        capi.bgeo_write_records(np.zeros(0, capi.RECORD_DTYPE), str(tmp_path / "e.bgeo"), 10)
