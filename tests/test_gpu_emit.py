"""K4 emission (get_all_voro_info keys) against the oracle restatement."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_emit_cfg1(ctx, O, cfg1):
    mesh, sites, knn, k = cfg1
    ctx.set_mesh(mesh)
    res = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, knn, k)
    recs = res.records()
    em = res.emit(mesh.n_surf_faces - 1)
    want = O.emit(recs, mesh.n_surf_faces - 1)
    for key in ("facet_cell", "facet_key", "facet_is_tet", "vert_cell", "vert_lvid", "vert_key", "vert_surf_fid",
                "edge_cell", "edge_key", "edge_lvid"):
        assert np.array_equal(em[key], want[key]), key
    # positions: same float expression (compute_vertex_coordinates) -> bitwise; contract is 1e-6 rel
    assert np.array_equal(em["vert_pos"].view(np.uint32), want["vert_pos"].view(np.uint32))
    _, _, eu = O.reload_active(recs, "oracle")
    assert np.array_equal(em["cell_euler"].view(np.uint32), eu.view(np.uint32))
