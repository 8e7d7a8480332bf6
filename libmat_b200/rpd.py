"""Host-side mirror of the reference interface for the hot path, in Python over the C ABI.

`Context.compute_clipped_voro_diagram` keeps the argument meaning of the reference's
compute_clipped_voro_diagram_GPU (src/rpd3d/voronoi.h:52-61) and `Context.compute_closest_dist2mat`
that of compute_closest_dist2mat (src/dist2mat/dist2mat.h:19-24); the C++ shims with the exact
C++ signatures are in include/libmat_b200_shim.hpp.  All compute happens in libmat_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import LibMatError, RECORD_DTYPE, ptr


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class RpdResult:
    """Handle on the device-resident result of one RPD run (mb_rpd_result)."""

    def __init__(self, ctx: "Context", handle, host_ptrs=None):
        self.ctx = ctx
        self._h = handle
        self._host_ptrs = host_ptrs  # streamed run: (blob address, offsets address) in pinned host memory
        lib = ctx.lib
        a, b, c = C.c_long(), C.c_long(), C.c_long()
        ctx._check(lib.mb_rpd_count(handle, C.byref(a), C.byref(b), C.byref(c)))
        self.n_cells, self.n_pairs, self.n_clips = a.value, b.value, c.value
        st = (C.c_long * 8)()
        ctx._check(lib.mb_rpd_stats(handle, st))
        self.n_culled, self.n_cand_overflow, self.n_big_pass_tets = st[3], st[4], st[6]
        self.n_exact = st[7]
        hist = (C.c_long * 10)()
        ctx._check(lib.mb_rpd_status_histogram(handle, hist))
        self.status_histogram = np.array(list(hist), dtype=np.int64)  # index = status + 1
        ms = (C.c_float * 4)()
        ctx._check(lib.mb_rpd_kernel_ms(handle, ms))
        self.kernel_ms = {"candidates": ms[0], "clip": ms[1], "order": ms[2], "total": ms[3]}
        nb = C.c_long()
        ctx._check(lib.mb_rpd_compact_bytes(handle, C.byref(nb)))
        self.compact_bytes = nb.value
        self.n_spans = int(lib.mb_rpd_spans(handle))
        a, b = C.c_long(), C.c_long()
        ctx._check(lib.mb_rpd_clip_passes(handle, C.byref(a), C.byref(b)))
        self.n_second_pass_cells, self.n_garbage_collections = a.value, b.value
        a, b = C.c_long(), C.c_long()
        ctx._check(lib.mb_rpd_flagged(handle, C.byref(a), C.byref(b)))
        self.n_flagged_cells, self.n_flagged_pairs = a.value, b.value

    def flags(self, pairs=False):
        """flagged class (a conflict |det| under the predicate_generator static-filter bound): uint8 per cell in record
        order, and with pairs=True also per candidate pair in the order of pairs() (one-shot runs only)"""
        cf = np.zeros(self.n_cells, np.uint8)
        pf = np.zeros(self.n_pairs, np.uint8) if pairs else None
        self.ctx._check(self.ctx.lib.mb_rpd_fetch_flags(self._h, ptr(cf), ptr(pf)))
        return (cf, pf) if pairs else cf

    def records(self) -> np.ndarray:
        """Cells sorted by (tet, site) in the ConvexCellTransfer layout (id = index)."""
        out = np.zeros(self.n_cells, dtype=RECORD_DTYPE)
        if self.n_cells:
            self.ctx._check(self.ctx.lib.mb_rpd_fetch_records(self._h, ptr(out)))
        return out

    def pairs(self):
        """candidate (tet, site) pairs of the run (the reference's tet_knn in pair form) + Status"""
        pt = np.zeros(self.n_pairs, np.int32)
        ps = np.zeros(self.n_pairs, np.int32)
        st = np.zeros(self.n_pairs, np.int8)
        self.ctx._check(self.ctx.lib.mb_rpd_fetch_pairs(self._h, ptr(pt), ptr(ps), ptr(st)))
        return pt, ps, st

    def compact(self):
        blob = np.zeros(max(1, self.compact_bytes // 4), dtype=np.uint32)
        offs = np.zeros(self.n_cells + 1, dtype=np.int64)
        self.ctx._check(self.ctx.lib.mb_rpd_fetch_compact(self._h, ptr(blob), ptr(offs)))
        return blob, offs

    def host_compact(self):
        """streamed run: zero-copy numpy views (blob uint32, offsets int64) of the library's pinned host
        result; valid until the next streamed run on the context"""
        assert self._host_ptrs is not None, "not a streamed result"
        bp, op = self._host_ptrs
        nw = self.compact_bytes // 4
        blob = np.ctypeslib.as_array(C.cast(bp, C.POINTER(C.c_uint32)), shape=(max(nw, 1),))[:nw]
        offs = np.ctypeslib.as_array(C.cast(op, C.POINTER(C.c_int64)), shape=(self.n_cells + 1,))
        return blob, offs

    def site_volumes(self):
        """per-site volume and barycentre sums (needs want_volumes=True); bary is SoA x|y|z"""
        n = self.ctx._n_site
        vol = np.zeros(n, np.float32)
        bary = np.zeros(3 * n, np.float32)
        self.ctx._check(self.ctx.lib.mb_rpd_site_volumes(self._h, ptr(vol), ptr(bary)))
        return vol, bary

    def cell_volumes(self):
        out = np.zeros(self.n_cells, np.float32)
        if self.n_cells:
            self.ctx._check(self.ctx.lib.mb_rpd_cell_volumes(self._h, ptr(out)))
        return out

    def device_buffers(self):
        b, o = C.c_void_p(), C.c_void_p()
        nb, nc = C.c_long(), C.c_long()
        self.ctx._check(self.ctx.lib.mb_rpd_device_buffers(self._h, C.byref(b), C.byref(nb), C.byref(o), C.byref(nc)))
        return b.value, nb.value, o.value, nc.value

    def emit(self, max_surf_fid: int) -> dict:
        """K4: facets / vertices / edges of every cell (get_all_voro_info keys)."""
        cnt = capi.EmitCounts()
        self.ctx._check(self.ctx.lib.mb_rpd_emit(self._h, int(max_surf_fid), C.byref(cnt)))
        nf, nv, ne = cnt.n_facets, cnt.n_vertices, cnt.n_edges
        out = {
            "facet_cell": np.zeros(nf, np.int32), "facet_key": np.zeros(nf, np.int32),
            "facet_is_tet": np.zeros(nf, np.uint8),
            "vert_cell": np.zeros(nv, np.int32), "vert_lvid": np.zeros(nv, np.int32),
            "vert_key": np.zeros((nv, 3), np.int32), "vert_pos": np.zeros((nv, 3), np.float32),
            "vert_surf_fid": np.zeros(nv, np.int32),
            "edge_cell": np.zeros(ne, np.int32), "edge_key": np.zeros((ne, 2), np.int32),
            "edge_lvid": np.zeros((ne, 2), np.int32),
            "cell_euler": np.zeros(self.n_cells, np.float32),
        }
        order = ["facet_cell", "facet_key", "facet_is_tet", "vert_cell", "vert_lvid", "vert_key",
                 "vert_pos", "vert_surf_fid", "edge_cell", "edge_key", "edge_lvid", "cell_euler"]
        self.ctx._check(self.ctx.lib.mb_rpd_fetch_emit(self._h, *[ptr(out[k]) for k in order]))
        # pc_face centroids of the surface facets (cell_to_surfv2fid) and, when the context holds a feature-edge map,
        # the covered sharp / concave edges of rpd_update.cxx:209-259
        out["facet_centroid"] = np.zeros((nf, 3), np.float32)
        if nf:
            self.ctx._check(self.ctx.lib.mb_rpd_fetch_facet_centroids(self._h, ptr(out["facet_centroid"])))
        nh = C.c_long(0)
        self.ctx._check(self.ctx.lib.mb_rpd_feature_edge_count(self._h, C.byref(nh)))
        out["fe_hit"] = np.zeros((nh.value, 6), np.int32)
        out["fe_end"] = np.zeros((2 * nh.value, 4), np.int32)
        out["fe_end_pos"] = np.zeros((2 * nh.value, 3), np.float32)
        if nh.value:
            self.ctx._check(self.ctx.lib.mb_rpd_fetch_feature_edges(self._h, ptr(out["fe_hit"]), ptr(out["fe_end"]), ptr(out["fe_end_pos"])))
        return out

    def topology(self) -> dict:
        """K6: cell / half-plane-facet connected components and Euler sums per power cell (call emit() first)"""
        cnt = capi.TopoCounts()
        self.ctx._check(self.ctx.lib.mb_rpd_topology(self._h, C.byref(cnt)))
        out = {"cell_cc": np.zeros(cnt.n_cells, np.int32), "facet_cc": np.zeros(cnt.n_facets, np.int32),
               "edge_cc": np.zeros(cnt.n_edges, np.int32),
               "site_n_cells": np.zeros(cnt.n_sites, np.int32), "site_n_cc": np.zeros(cnt.n_sites, np.int32),
               "site_euler_sum": np.zeros(cnt.n_sites, np.float64),
               "pair_site": np.zeros(cnt.n_halfplane_pairs, np.int32), "pair_neigh": np.zeros(cnt.n_halfplane_pairs, np.int32),
               "pair_n_cc": np.zeros(cnt.n_halfplane_pairs, np.int32)}
        order = ["cell_cc", "facet_cc", "edge_cc", "site_n_cells", "site_n_cc", "site_euler_sum", "pair_site", "pair_neigh", "pair_n_cc"]
        self.ctx._check(self.ctx.lib.mb_rpd_fetch_topology(self._h, *[ptr(out[k]) for k in order]))
        return out

    def free(self):
        if self._h:
            self.ctx.lib.mb_rpd_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One mb_ctx: one CUDA device, one stream, the tet mesh resident in HBM."""

    def __init__(self, device: int = -1):
        self.lib = capi.load()
        err = C.c_int(0)
        self._ctx = self.lib.mb_create(int(device), C.byref(err))
        if not self._ctx:
            raise LibMatError(f"mb_create failed (code {err.value}): no usable CUDA device; "
                              "libmat_b200 has no CPU fallback")
        self.mesh_n_tet = 0

    def _check(self, rc: int):
        if rc != 0:
            msg = self.lib.mb_last_error(self._ctx)
            raise LibMatError(f"libmat_b200 error {rc}: {msg.decode() if msg else ''}")

    def set_stream(self, cuda_stream: int | None):
        """run on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream)"""
        self._check(self.lib.mb_set_stream(self._ctx, C.c_void_p(cuda_stream) if cuda_stream else None))

    def measure_peaks(self):
        """(FP32 TFLOP/s, FP64 TFLOP/s) measured with FFMA / DFMA loops on this device"""
        a, b = C.c_double(0), C.c_double(0)
        self._check(self.lib.mb_measure_peaks(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def launch_count(self) -> int:
        n = C.c_ulonglong(0)
        self._check(self.lib.mb_launch_count(self._ctx, C.byref(n)))
        return n.value

    def close(self):
        if self._ctx:
            self.lib.mb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ mesh
    def set_tetmesh(self, vertices, indices, v_adjs, f_adjs, f_ids, e_adj6=None, e_adjs_dense=None):
        v = _c(vertices, np.float32).reshape(-1)
        idx = _c(indices, np.int32).reshape(-1)
        va = _c(v_adjs, np.int32)
        fa = _c(f_adjs, np.int32).reshape(-1)
        fi = _c(f_ids, np.int32).reshape(-1)
        e6 = None if e_adj6 is None else _c(e_adj6, np.int32).reshape(-1)
        ed = None if e_adjs_dense is None else _c(e_adjs_dense, np.int32)
        self._check(self.lib.mb_set_tetmesh(self._ctx, ptr(v), v.size // 3, ptr(idx), idx.size // 4,
                                            ptr(va), ptr(ed), ptr(e6), ptr(fa), ptr(fi)))
        self.mesh_n_tet = idx.size // 4

    def set_mesh(self, mesh):
        self.set_tetmesh(mesh.vertices, mesh.indices, mesh.v_adjs, mesh.f_adjs, mesh.f_ids, e_adj6=mesh.e_adj6)

    def set_feature_edges(self, rows6=None):
        """TetMesh::tet_es2fe_map of the resident mesh: rows (tet, lf_min, lf_max, fe_type, fe_id, fe_line_id); None clears"""
        if rows6 is None or len(rows6) == 0:
            self._check(self.lib.mb_set_feature_edges(self._ctx, None, 0))
        else:
            r = _c(rows6, np.int32).reshape(-1, 6)
            self._check(self.lib.mb_set_feature_edges(self._ctx, ptr(r), len(r)))

    def set_tet_range(self, first: int, count: int):
        self._check(self.lib.mb_set_tet_range(self._ctx, int(first), int(count)))

    def set_tet_id_base(self, base: int):
        """records carry (uploaded tet index + base): for ranks that upload only their shard of the tets"""
        self._check(self.lib.mb_set_tet_id_base(self._ctx, int(base)))

    def set_tet_subset(self, tet_ids=None):
        """process only the listed tets (ascending global ids); None / empty clears the subset"""
        if tet_ids is None or len(tet_ids) == 0:
            self._check(self.lib.mb_set_tet_subset(self._ctx, None, 0))
        else:
            ids = _c(tet_ids, np.int32)
            self._check(self.lib.mb_set_tet_subset(self._ctx, ptr(ids), ids.size))

    # ------------------------------------------------------------------ RPD
    def upload_sites(self, site_soa, site_weights, site_flags, site_knn=None, site_k=0):
        ss = _c(site_soa, np.float32).reshape(-1)
        sw = _c(site_weights, np.float32)
        sf = _c(site_flags, np.uint32)
        knn = None if site_knn is None else _c(site_knn, np.int32).reshape(-1)
        self._keep = (ss, sw, sf, knn)
        self._n_site = sw.size
        self._check(self.lib.mb_rpd_upload_sites(self._ctx, ptr(ss), ptr(sw), ptr(sf), sw.size, ptr(knn), int(site_k)))

    def run(self, lanes_per_cell=0, grid_k=0, want_volumes=False, grid_candidates=False, security_radius=False) -> RpdResult:
        opts = capi.RpdOpts(int(lanes_per_cell), int(grid_k), int(want_volumes), int(grid_candidates), 0, int(security_radius))
        h = C.c_void_p()
        self._check(self.lib.mb_rpd_run(self._ctx, C.byref(opts), C.byref(h)))
        self._check(self.lib.mb_rpd_sync(self._ctx, h))
        return RpdResult(self, h)

    def run_to_host(self, n_chunks=0, lanes_per_cell=0, grid_k=0, grid_candidates=False, lean=False) -> RpdResult:
        """streamed run (mb_rpd_run_to_host): D2H of span c overlaps the kernels of span c+1; the complete
        compact result is in pinned host memory on return (RpdResult.host_compact())"""
        opts = capi.RpdOpts(int(lanes_per_cell), int(grid_k), 0, int(grid_candidates), int(lean))
        h, bp, op = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self.lib.mb_rpd_run_to_host(self._ctx, C.byref(opts), int(n_chunks), C.byref(h), C.byref(bp), C.byref(op)))
        return RpdResult(self, h, host_ptrs=(bp.value, op.value))

    def run_incremental(self, to_host=False, n_chunks=0, lanes_per_cell=0, grid_k=0, lean=0):
        """mb_rpd_run_incremental: clips only the tets whose candidate list changed since the previous incremental
        run.  Returns (RpdResult of the affected tets, ascending affected tet ids)."""
        opts = capi.RpdOpts(int(lanes_per_cell), int(grid_k), 0, 0, int(lean) if to_host else 0)
        h, bp, op = C.c_void_p(), C.c_void_p(), C.c_void_p()
        na = C.c_long(0)
        self._check(self.lib.mb_rpd_run_incremental(self._ctx, C.byref(opts), int(bool(to_host)), int(n_chunks), C.byref(h),
                                                    C.byref(na), C.byref(bp), C.byref(op)))
        tets = np.zeros(na.value, np.int32)
        if na.value:
            self._check(self.lib.mb_rpd_fetch_affected_tets(self._ctx, ptr(tets)))
        return RpdResult(self, h, host_ptrs=(bp.value, op.value) if to_host else None), tets

    def run_to_sink(self, sink_blob_ptr: int, cap_bytes: int, sink_off_ptr: int, cap_cells: int, n_chunks=0,
                    lanes_per_cell=0, grid_k=0, grid_candidates=False, lean=False) -> RpdResult:
        """streamed run into caller memory (mb_rpd_run_to_sink): pinned / registered host memory, device memory
        of this GPU or of a peer GPU (mb_sink_open) -- the multi-GPU gather fused into the run"""
        opts = capi.RpdOpts(int(lanes_per_cell), int(grid_k), 0, int(grid_candidates), int(lean))
        h = C.c_void_p()
        self._check(self.lib.mb_rpd_run_to_sink(self._ctx, C.byref(opts), int(n_chunks), C.c_void_p(sink_blob_ptr),
                                                int(cap_bytes), C.c_void_p(sink_off_ptr), int(cap_cells), C.byref(h)))
        return RpdResult(self, h)

    # sink memory (multi-GPU gather): device buffer + CUDA IPC handle, peer mapping, host page-locking
    def sink_create(self, nbytes: int):
        p = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        self._check(self.lib.mb_sink_create(self._ctx, int(nbytes), C.byref(p), handle))
        return p.value, bytes(handle)

    def sink_destroy(self, ptr_: int):
        self._check(self.lib.mb_sink_destroy(self._ctx, C.c_void_p(ptr_)))

    def sink_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        self._check(self.lib.mb_sink_open(self._ctx, buf, C.byref(p)))
        return p.value

    def sink_close(self, ptr_: int):
        self._check(self.lib.mb_sink_close(self._ctx, C.c_void_p(ptr_)))

    def host_register(self, ptr_: int, nbytes: int):
        self._check(self.lib.mb_host_register(self._ctx, C.c_void_p(ptr_), int(nbytes)))

    def host_unregister(self, ptr_: int):
        self._check(self.lib.mb_host_unregister(self._ctx, C.c_void_p(ptr_)))

    def copy_to_host(self, host_array: np.ndarray, d_src: int, nbytes: int):
        self._check(self.lib.mb_copy_to_host(self._ctx, ptr(host_array), C.c_void_p(d_src), int(nbytes)))

    def write_bgeo(self, blob: np.ndarray, offsets: np.ndarray, path: str, max_sf_fid: int, is_boundary_only=False):
        """compact records (full or lean) in host memory -> Houdini .bgeo, the IO_CUDA result format
        (save_convex_cells_houdini, io_cuda.cxx:152-187); returns (n_points, n_polygons)"""
        offsets = np.ascontiguousarray(offsets, np.int64)
        a, b = C.c_long(), C.c_long()
        self._check(self.lib.mb_rpd_write_bgeo(self._ctx, ptr(np.ascontiguousarray(blob)), ptr(offsets), len(offsets) - 1,
                                               int(max_sf_fid), int(is_boundary_only), path.encode(), C.byref(a), C.byref(b)))
        return a.value, b.value

    def expand_compact(self, blob: np.ndarray, offsets: np.ndarray, first_id: int = 0) -> np.ndarray:
        """compact records (full or lean, e.g. gathered from several ranks) -> ConvexCellTransfer records; the
        context must hold the mesh and sites they were computed from"""
        offsets = np.ascontiguousarray(offsets, np.int64)
        n = len(offsets) - 1
        out = np.zeros(n, dtype=RECORD_DTYPE)
        if n:
            self._check(self.lib.mb_rpd_expand_compact(self._ctx, ptr(np.ascontiguousarray(blob)), ptr(offsets), n, int(first_id), ptr(out)))
        return out

    def compute_clipped_voro_diagram(self, site, site_weights, site_flags, site_knn=None, site_k=0,
                                     **opts) -> RpdResult:
        """site: SoA x|y|z float[3*n_site]; site_weights r^2; site_knn (site_k+1) x n_site or None
        (None = grid-kNN mode).  Returns the device-resident result (cells sorted by (tet, site))."""
        self.upload_sites(site, site_weights, site_flags, site_knn, site_k)
        return self.run(**opts)

    # ------------------------------------------------------------------ dist2mat
    def dist2mat_upload(self, spheres, samples, offset, count, prims):
        sp = _c(spheres, np.float32).reshape(-1)
        sm = _c(samples, np.float32).reshape(-1)
        of = _c(offset, np.uint32)
        cn = _c(count, np.uint32)
        pr = _c(prims, np.int32).reshape(-1)
        self._d2m_n = of.size
        self._check(self.lib.mb_dist2mat_upload(self._ctx, ptr(sp), sp.size // 4, ptr(sm), of.size, ptr(of),
                                                ptr(cn), ptr(pr), pr.size // 3))

    def dist2mat_run(self) -> float:
        ms = C.c_float(0)
        self._check(self.lib.mb_dist2mat_run(self._ctx, C.byref(ms)))
        return ms.value

    def dist2mat_fetch(self, want_tie=True):
        n = self._d2m_n
        res = np.zeros(n, np.float32)
        cid = np.zeros(n, np.int32)
        tie = np.zeros(n, np.uint8) if want_tie else None
        self._check(self.lib.mb_dist2mat_fetch(self._ctx, ptr(res), ptr(cid), ptr(tie)))
        return res, cid, tie

    # ---- f3: candidate lists built on the device (fix_geo_error.cxx:149-215 semantics) --------------------
    def dist2mat_set_medial_mesh(self, spheres, mm_faces, mm_edges):
        sp = _c(spheres, np.float32).reshape(-1)
        mf = _c(mm_faces, np.int32).reshape(-1)
        me = _c(mm_edges, np.int32).reshape(-1)
        self._check(self.lib.mb_dist2mat_set_medial_mesh(self._ctx, ptr(sp), sp.size // 4, ptr(mf), mf.size // 3, ptr(me), me.size // 2))

    def dist2mat_set_face_sites(self, fid_site_rows, n_fid):
        r = _c(fid_site_rows, np.int32).reshape(-1, 2)
        self._check(self.lib.mb_dist2mat_set_face_sites(self._ctx, ptr(r), len(r), int(n_fid)))

    def dist2mat_set_face_sites_from_rpd(self, res: "RpdResult", max_surf_fid: int):
        self._check(self.lib.mb_dist2mat_set_face_sites_from_rpd(self._ctx, res._h, int(max_surf_fid)))

    def dist2mat_upload_by_face(self, samples, sample_fid):
        sm = _c(samples, np.float32).reshape(-1)
        fi = _c(sample_fid, np.int32)
        self._d2m_n = fi.size
        self._check(self.lib.mb_dist2mat_upload_by_face(self._ctx, ptr(sm), ptr(fi), fi.size))

    def dist2mat_closest_prims(self):
        out = np.zeros((self._d2m_n, 3), np.int32)
        if self._d2m_n:
            self._check(self.lib.mb_dist2mat_fetch_closest_prims(self._ctx, ptr(out)))
        return out

    def dist2mat_by_face(self, samples, sample_fid, want_tie=True):
        """samples + surface-face ids against the device-built per-face lists: (result, closest_id[, tie], closest_prim3)"""
        self.dist2mat_upload_by_face(samples, sample_fid)
        self.dist2mat_run()
        return (*self.dist2mat_fetch(want_tie), self.dist2mat_closest_prims())

    def dist2mat_face_lists(self):
        a, b = C.c_long(), C.c_long()
        self._check(self.lib.mb_dist2mat_face_list_size(self._ctx, C.byref(a), C.byref(b)))
        off = np.zeros(a.value + 1, np.int64)
        prims = np.zeros((b.value, 3), np.int32)
        self._check(self.lib.mb_dist2mat_fetch_face_lists(self._ctx, ptr(off), ptr(prims)))
        return off, prims

    def compute_closest_dist2mat(self, spheres, samples, offset, count, prims, want_tie=True):
        """Argument meaning of compute_closest_dist2mat (reference dist2mat.h:19-24):
        returns (results, closest_mat_id[, tie_flag])."""
        self.dist2mat_upload(spheres, samples, offset, count, prims)
        self.dist2mat_run()
        return self.dist2mat_fetch(want_tie)
