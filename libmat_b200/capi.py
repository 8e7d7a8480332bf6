"""ctypes binding of the C ABI (include/libmat_b200.h).  The shared library is built in-tree by
libmat_b200/build.py; a missing library is an error -- there is no CPU or Python fallback."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmat_b200.so")

# every symbol include/libmat_b200.h declares
SYMBOLS = [
    "mb_create", "mb_destroy", "mb_last_error", "mb_version", "mb_predicate_bounds", "mb_tet_adjacency", "mb_measure_peaks", "mb_launch_count", "mb_set_stream", "mb_rpd_fetch_pairs", "mb_set_tetmesh", "mb_set_tet_range", "mb_set_tet_id_base", "mb_set_tet_subset",
    "mb_rpd3d", "mb_rpd_upload_sites", "mb_rpd_run", "mb_rpd_run_to_host", "mb_rpd_run_incremental", "mb_rpd_fetch_affected_tets", "mb_rpd_merge_compact", "mb_rpd_run_to_sink", "mb_rpd_expand_compact", "mb_rpd_spans", "mb_sink_create", "mb_sink_destroy", "mb_sink_open",
    "mb_sink_close", "mb_host_register", "mb_host_unregister", "mb_copy_to_host", "mb_rpd_sync", "mb_rpd_free", "mb_rpd_count",
    "mb_rpd_status_histogram", "mb_rpd_clip_passes", "mb_rpd_flagged", "mb_rpd_fetch_flags", "mb_debug_set_pair_hint", "mb_rpd_stats", "mb_rpd_kernel_ms", "mb_rpd_fetch_records", "mb_rpd_compact_bytes",
    "mb_rpd_fetch_compact", "mb_rpd_site_volumes", "mb_rpd_cell_volumes", "mb_rpd_device_buffers", "mb_rpd_emit",
    "mb_rpd_fetch_emit", "mb_rpd_fetch_facet_centroids", "mb_set_feature_edges", "mb_rpd_feature_edge_count",
    "mb_rpd_fetch_feature_edges", "mb_rpd_write_bgeo", "mb_bgeo_write_records", "mb_rpd_topology", "mb_rpd_fetch_topology", "mb_dist2mat", "mb_dist2mat_upload", "mb_dist2mat_run", "mb_dist2mat_fetch",
    "mb_dist2mat_set_medial_mesh", "mb_dist2mat_set_face_sites", "mb_dist2mat_set_face_sites_from_rpd", "mb_dist2mat_upload_by_face",
    "mb_dist2mat_by_face", "mb_dist2mat_fetch_closest_prims", "mb_dist2mat_face_list_size", "mb_dist2mat_fetch_face_lists",
]

# static-filter bounds (reference src/predicate_generator/main.cpp output; include/libmat_b200.h)
FILTER_BOUND_F64 = 1.2466136531027298e-13
FILTER_BOUND_F32 = 6.6876506e-05

RECORD_BYTES = 3456
# ConvexCellTransfer layout (reference src/rpd3d/convex_cell.h:189-217)
RECORD_DTYPE = np.dtype({
    "names": ["status", "thread_id", "voro_id", "tet_id", "weight", "is_active", "nb_v", "nb_p",
              "nb_e", "ver", "clip", "id2", "edge", "euler", "cell_vol", "id"],
    "formats": ["<i4", "<i4", "<i4", "<i4", "<f4", "u1", "u1", "u1", "u1", ("u1", (96, 4)),
                ("<f4", (64, 8)), ("<i4", (64, 2)), ("u1", (152, 3)), "<f4", "<f4", "<i4"],
    "offsets": [0, 4, 8, 12, 16, 20, 21, 22, 23, 24, 416, 2464, 2976, 3432, 3436, 3440],
    "itemsize": RECORD_BYTES,
})


class RpdOpts(C.Structure):
    _fields_ = [("lanes_per_cell", C.c_int), ("grid_k", C.c_int), ("want_volumes", C.c_int),
                ("grid_candidates", C.c_int), ("lean_records", C.c_int), ("security_radius", C.c_int)]


class EmitCounts(C.Structure):
    _fields_ = [("n_facets", C.c_long), ("n_vertices", C.c_long), ("n_edges", C.c_long)]


class TopoCounts(C.Structure):
    _fields_ = [("n_cells", C.c_long), ("n_facets", C.c_long), ("n_sites", C.c_long), ("n_halfplane_pairs", C.c_long),
                ("n_edges", C.c_long)]


class LibMatError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load libmat_b200.so; raises if it has not been built (never falls back)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibMatError(f"{LIB_PATH} is missing: run `python -m libmat_b200.build` (needs nvcc); "
                          "libmat_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    lib.mb_create.restype = C.c_void_p
    lib.mb_create.argtypes = [C.c_int, C.POINTER(C.c_int)]
    lib.mb_destroy.argtypes = [C.c_void_p]
    lib.mb_destroy.restype = None
    lib.mb_last_error.restype = C.c_char_p
    lib.mb_last_error.argtypes = [C.c_void_p]
    lib.mb_version.restype = C.c_char_p
    lib.mb_predicate_bounds.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_float)]
    lib.mb_predicate_bounds.restype = None
    lib.mb_rpd_free.argtypes = [C.c_void_p]
    lib.mb_rpd_free.restype = None
    vp = C.c_void_p
    lib.mb_set_tetmesh.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, vp]
    lib.mb_set_tet_range.argtypes = [vp, C.c_int, C.c_int]
    lib.mb_set_tet_id_base.argtypes = [vp, C.c_int]
    lib.mb_set_tet_subset.argtypes = [vp, vp, C.c_int]
    lib.mb_rpd3d.argtypes = [vp, vp, vp, vp, C.c_int, vp, C.c_int, vp, C.POINTER(vp)]
    lib.mb_rpd_upload_sites.argtypes = [vp, vp, vp, vp, C.c_int, vp, C.c_int]
    lib.mb_rpd_run.argtypes = [vp, vp, C.POINTER(vp)]
    lib.mb_rpd_sync.argtypes = [vp, vp]
    lib.mb_rpd_spans.argtypes = [vp]
    lib.mb_rpd_expand_compact.argtypes = [vp, vp, vp, C.c_long, C.c_long, vp]
    lib.mb_rpd_run_to_sink.argtypes = [vp, vp, C.c_int, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(vp)]
    lib.mb_sink_create.argtypes = [vp, C.c_size_t, C.POINTER(vp), vp]
    lib.mb_sink_destroy.argtypes = [vp, vp]
    lib.mb_sink_open.argtypes = [vp, vp, C.POINTER(vp)]
    lib.mb_sink_close.argtypes = [vp, vp]
    lib.mb_host_register.argtypes = [vp, vp, C.c_size_t]
    lib.mb_host_unregister.argtypes = [vp, vp]
    lib.mb_copy_to_host.argtypes = [vp, vp, vp, C.c_size_t]
    lib.mb_rpd_run_to_host.argtypes = [vp, vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.mb_rpd_run_incremental.argtypes = [vp, vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(C.c_long), C.POINTER(vp), C.POINTER(vp)]
    lib.mb_rpd_fetch_affected_tets.argtypes = [vp, vp]
    lib.mb_rpd_merge_compact.argtypes = [vp, vp, C.c_long, vp, vp, C.c_long, vp, C.c_long, vp, vp, C.POINTER(C.c_long), C.POINTER(C.c_long)]
    lib.mb_rpd_count.argtypes = [vp, C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(C.c_long)]
    lib.mb_rpd_status_histogram.argtypes = [vp, vp]
    lib.mb_rpd_clip_passes.argtypes = [vp, C.POINTER(C.c_long), C.POINTER(C.c_long)]
    lib.mb_debug_set_pair_hint.argtypes = [vp, C.c_double]
    lib.mb_rpd_flagged.argtypes = [vp, C.POINTER(C.c_long), C.POINTER(C.c_long)]
    lib.mb_rpd_fetch_flags.argtypes = [vp, vp, vp]
    lib.mb_rpd_stats.argtypes = [vp, vp]
    lib.mb_rpd_kernel_ms.argtypes = [vp, vp]
    lib.mb_rpd_fetch_records.argtypes = [vp, vp]
    lib.mb_rpd_fetch_pairs.argtypes = [vp, vp, vp, vp]
    lib.mb_set_stream.argtypes = [vp, vp]
    lib.mb_launch_count.argtypes = [vp, C.POINTER(C.c_ulonglong)]
    lib.mb_tet_adjacency.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, vp, vp, vp, vp, C.POINTER(C.c_int)]
    lib.mb_measure_peaks.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.mb_rpd_compact_bytes.argtypes = [vp, C.POINTER(C.c_long)]
    lib.mb_rpd_fetch_compact.argtypes = [vp, vp, vp]
    lib.mb_rpd_site_volumes.argtypes = [vp, vp, vp]
    lib.mb_rpd_cell_volumes.argtypes = [vp, vp]
    lib.mb_rpd_device_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_long), C.POINTER(vp), C.POINTER(C.c_long)]
    lib.mb_rpd_emit.argtypes = [vp, C.c_int, C.POINTER(EmitCounts)]
    lib.mb_rpd_fetch_emit.argtypes = [vp] + [vp] * 12
    lib.mb_rpd_fetch_facet_centroids.argtypes = [vp, vp]
    lib.mb_set_feature_edges.argtypes = [vp, vp, C.c_long]
    lib.mb_rpd_feature_edge_count.argtypes = [vp, C.POINTER(C.c_long)]
    lib.mb_rpd_fetch_feature_edges.argtypes = [vp, vp, vp, vp]
    lib.mb_rpd_write_bgeo.argtypes = [vp, vp, vp, C.c_long, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_long), C.POINTER(C.c_long)]
    lib.mb_bgeo_write_records.argtypes = [vp, C.c_long, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_long), C.POINTER(C.c_long)]
    lib.mb_rpd_topology.argtypes = [vp, C.POINTER(TopoCounts)]
    lib.mb_rpd_fetch_topology.argtypes = [vp] + [vp] * 9
    lib.mb_dist2mat.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, C.c_long, vp, vp, vp]
    lib.mb_dist2mat_upload.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, C.c_long]
    lib.mb_dist2mat_run.argtypes = [vp, C.POINTER(C.c_float)]
    lib.mb_dist2mat_fetch.argtypes = [vp, vp, vp, vp]
    lib.mb_dist2mat_set_medial_mesh.argtypes = [vp, vp, C.c_int, vp, C.c_int, vp, C.c_int]
    lib.mb_dist2mat_set_face_sites.argtypes = [vp, vp, C.c_long, C.c_int]
    lib.mb_dist2mat_set_face_sites_from_rpd.argtypes = [vp, vp, C.c_int]
    lib.mb_dist2mat_upload_by_face.argtypes = [vp, vp, vp, C.c_int]
    lib.mb_dist2mat_by_face.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp]
    lib.mb_dist2mat_fetch_closest_prims.argtypes = [vp, vp]
    lib.mb_dist2mat_face_list_size.argtypes = [vp, C.POINTER(C.c_long), C.POINTER(C.c_long)]
    lib.mb_dist2mat_fetch_face_lists.argtypes = [vp, vp, vp]
    _lib = lib
    return lib


def bgeo_write_records(records: np.ndarray, path: str, max_sf_fid: int, is_boundary_only: bool = False):
    """ConvexCellTransfer records -> Houdini .bgeo (the IO_CUDA result format); host code only, no GPU needed.
    Returns (n_points, n_polygons)."""
    lib = load()
    recs = np.ascontiguousarray(records)
    a, b = C.c_long(), C.c_long()
    rc = lib.mb_bgeo_write_records(ptr(recs), len(recs), int(max_sf_fid), int(is_boundary_only), path.encode(), C.byref(a), C.byref(b))
    if rc != 0:
        raise LibMatError(f"mb_bgeo_write_records failed ({rc})")
    return a.value, b.value


def tet_adjacency(indices, n_vert, boundary_sf_fids=None, n_sf_facets=None):
    """mb_tet_adjacency: (v_adjs, e_adj6 [n_tet,6], f_adjs [n_tet,4], f_ids [n_tet,4], n_boundary_faces) from the tet
    indices -- load_tet_adj_info (io.cxx:238-335) without the dense edge table; host code"""
    lib = load()
    idx = np.ascontiguousarray(indices, np.int32).reshape(-1, 4)
    n_tet = len(idx)
    v = np.zeros(n_vert, np.int32)
    e6 = np.zeros((n_tet, 6), np.int32)
    fa = np.zeros((n_tet, 4), np.int32)
    fi = np.zeros((n_tet, 4), np.int32)
    nb = C.c_int(0)
    bs = None if boundary_sf_fids is None else np.ascontiguousarray(boundary_sf_fids, np.int32)
    if n_sf_facets is None:
        # default numbering: boundary faces 0..n_b-1, interior faces after them
        rc = lib.mb_tet_adjacency(ptr(idx), n_tet, int(n_vert), None, 0, None, None, ptr(fa), None, C.byref(nb))
        if rc != 0:
            raise LibMatError(f"mb_tet_adjacency failed ({rc})")
        n_sf_facets = nb.value
    rc = lib.mb_tet_adjacency(ptr(idx), n_tet, int(n_vert), ptr(bs), int(n_sf_facets), ptr(v), ptr(e6), ptr(fa), ptr(fi), C.byref(nb))
    if rc != 0:
        raise LibMatError(f"mb_tet_adjacency failed ({rc})")
    return v, e6, fa, fi, nb.value


def merge_compact(prev_blob, prev_offs, patch_blob, patch_offs, affected_tets):
    """mb_rpd_merge_compact: (blob uint32, offsets int64) of the previous result with the affected tets' records replaced"""
    lib = load()
    po = np.ascontiguousarray(prev_offs, np.int64)
    qo = np.ascontiguousarray(patch_offs, np.int64)
    pb = np.ascontiguousarray(prev_blob)
    qb = np.ascontiguousarray(patch_blob)
    aff = np.ascontiguousarray(affected_tets, np.int32)
    n_prev, n_patch = len(po) - 1, len(qo) - 1
    out_off = np.zeros(n_prev + n_patch + 1, np.int64)
    out = np.zeros((int(po[-1]) + int(qo[-1])) // 4 + 1, np.uint32)
    n, nb = C.c_long(0), C.c_long(0)
    rc = lib.mb_rpd_merge_compact(ptr(pb), ptr(po), n_prev, ptr(qb), ptr(qo), n_patch, ptr(aff), len(aff), ptr(out), ptr(out_off),
                                  C.byref(n), C.byref(nb))
    if rc != 0:
        raise LibMatError(f"mb_rpd_merge_compact failed ({rc})")
    return out[: nb.value // 4], out_off[: n.value + 1]


def ptr(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)
