"""mb_rpd_write_bgeo on the GPU path: the .bgeo written from a streamed run's compact records (full and lean format)
equals the reference's save_convex_cells_houdini output on the same cells, byte for byte."""
import numpy as np
import pytest

from libmat_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("lean", [0, 1, 2])
def test_bgeo_from_streamed_run(ctx, O, cfg1_rt, tmp_path, lean):
    mesh, sites, knn, k = cfg1_rt
    ctx.set_mesh(mesh)
    ctx.upload_sites(sites.site_soa, sites.weights, sites.flags, None, 0)
    res = ctx.run_to_host(n_chunks=3, lean=lean)
    blob, offs = res.host_compact()
    path = str(tmp_path / "gpu.bgeo")
    n_pts, n_poly = ctx.write_bgeo(blob, offs, path, mesh.n_surf_faces - 1)
    recs = res.records()
    res.free()
    assert n_pts == int(recs["nb_v"].sum()) and n_poly > 0
    got = open(path, "rb").read()
    path2 = str(tmp_path / "recs.bgeo")
    capi.bgeo_write_records(recs, path2, mesh.n_surf_faces - 1)
    assert got == open(path2, "rb").read()
    if O.ref("bgeo") is not None:
        assert got == O.ref_bgeo(recs, mesh.n_surf_faces - 1, False, str(tmp_path / "work"))
