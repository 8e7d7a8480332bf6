/* libmat_b200 -- C ABI of the B200-native RPD3D + dist2mat hot path.
 *
 * Plain C, POD only, every size explicit.  These entry points are what LibMAT's host code binds
 * in place of its own CUDA layer:
 *
 *   reference interface                                              replaced by
 *   ---------------------------------------------------------------  ---------------------------
 *   compute_clipped_voro_diagram_GPU  src/rpd3d/voronoi.h:52-61      mb_set_tetmesh + mb_rpd3d +
 *     (body src/rpd3d/voronoi.cu:455-795)                            mb_rpd_count/mb_rpd_fetch_*
 *   copy_tet_data / load_num_adjacent_cells_and_ids                  mb_set_tetmesh
 *     src/rpd3d/voronoi.cu:324-362, 379-415
 *   compute_tet_sphere_relation  src/rpd3d/voronoi.cu:198-322        inside mb_rpd3d (K2)
 *   clipped_voro_cell_test_GPU_param_tet                             inside mb_rpd3d (K3)
 *     src/rpd3d/convex_cell.cu:1166-1337
 *   ConvexCellTransfer D2H + copy_cc  voronoi.cu:433-449, 717-769    mb_rpd_fetch_records
 *   reload_active / get_all_voro_info                                mb_rpd_fetch_emit (K4)
 *     src/rpd3d_base/voronoi_defs.cxx:76-106, rpd_update.cxx:112-301
 *   compute_closest_dist2mat  src/dist2mat/dist2mat.h:19-24          mb_dist2mat
 *     (body src/dist2mat/dist2mat.cu:280-315)
 *
 * The C++ shims with the reference's exact signatures live in include/libmat_b200_shim.hpp.
 *
 * Conventions: every function returns 0 on success or a negative mb_status; the message is
 * available from mb_last_error().  Nothing calls exit().  Inputs are caller-owned host memory,
 * read-only, and may be freed as soon as the call returns (the upload calls synchronise their copies).  Results are library-owned handles
 * with explicit free; bulk results use the two-call count -> fetch pattern and the caller
 * allocates the destination.  One mb_ctx per host thread / per GPU; no global state, no files.
 * There is NO CPU fallback: without a usable CUDA device mb_create() fails.
 */
#ifndef LIBMAT_B200_H
#define LIBMAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mb_ctx mb_ctx;
typedef struct mb_rpd_result mb_rpd_result;

enum mb_status {
  MB_OK = 0,
  MB_ERR_CUDA = -1,     /* a CUDA runtime call failed                      */
  MB_ERR_ARG = -2,      /* invalid argument                                */
  MB_ERR_STATE = -3,    /* call order (e.g. mb_rpd3d before mb_set_tetmesh) */
  MB_ERR_NOMEM = -4,    /* host or device allocation failed                */
  MB_ERR_NODEVICE = -5  /* no CUDA device                                  */
};

/* Per-cell status values, identical to the reference enum Status
 * (src/rpd3d_base/voronoi_common.h:10-25). */
enum mb_cell_status {
  MB_CELL_early_return = -1,
  MB_CELL_triangle_overflow = 0,
  MB_CELL_vertex_overflow = 1,
  MB_CELL_inconsistent_boundary = 2,
  MB_CELL_security_radius_not_reached = 3,
  MB_CELL_success = 4,
  MB_CELL_needs_exact_predicates = 5,
  MB_CELL_no_intersection = 6,
  MB_CELL_edge_overflow = 7,
  MB_CELL_needs_perturb = 8
};

#define MB_MAX_P 64  /* _MAX_P_ */
#define MB_MAX_T 96  /* _MAX_T_ */
#define MB_MAX_E 152 /* _MAX_E_ */
#define MB_RECORD_BYTES 3456 /* sizeof(ConvexCellTransfer), src/rpd3d/convex_cell.h:189-217 */

/* ---------------------------------------------------------------- context */
/* device < 0 -> current device.  Returns NULL without a CUDA device (err may receive the code). */
mb_ctx* mb_create(int device, int* err);
void mb_destroy(mb_ctx* ctx);
/* run all subsequent work of this context on the caller's CUDA stream (cudaStream_t passed as
 * void*; NULL = back to the context's own stream).  Lets a host framework (e.g. torch + NCCL) order
 * its own work and events against the library's kernels. */
int mb_set_stream(mb_ctx* ctx, void* cuda_stream);
const char* mb_last_error(const mb_ctx* ctx);
const char* mb_version(void);
/* the static-filter error bounds of the conflict predicate, as printed by the reference's
 * src/predicate_generator/main.cpp:52-78 and consumed at convex_cell.cu:421,486 */
#define MB_FILTER_BOUND_F64 1.2466136531027298e-13
#define MB_FILTER_BOUND_F32 6.6876506e-05f
void mb_predicate_bounds(double* bound_f64, float* bound_f32);
/* measured FFMA / DFMA throughput of the device in TFLOP/s (instrumentation: the compute roofline bench.py reports
 * for the issue-bound kernels next to the HBM one) */
int mb_measure_peaks(mb_ctx* ctx, double* fp32_tflops, double* fp64_tflops);
/* number of CUDA kernels this context has launched since mb_create (instrumentation) */
int mb_launch_count(const mb_ctx* ctx, unsigned long long* n_launches);

/* ---------------------------------------------------------------- tet mesh (resident in HBM) */
/* verts_aos float[3*n_vert], idx_aos int[4*n_tet] (positively oriented), v_adjs int[n_vert],
 * f_adjs / f_ids int[4*n_tet] in tet_faces_lvid order (convex_cell.h:30-31).
 * Edge adjacency: either e_adjs_dense (the reference's triangular table of
 * n_vert(n_vert+1)/2+1 ints indexed by get_edge_idx, convex_cell.h:46-66) or e_adj6
 * (6 ints per tet in the (0,1)(0,2)(0,3)(1,2)(1,3)(2,3) local-vertex order of
 * convex_cell.cu:194-207); exactly one may be NULL.  Partial-tet calls (rpd_api.cxx:254-281)
 * pass the subset as idx_aos/f_adjs/f_ids with global vertices. */
int mb_set_tetmesh(mb_ctx* ctx, const float* verts_aos, int n_vert, const int* idx_aos, int n_tet,
                   const int* v_adjs, const int* e_adjs_dense, const int* e_adj6,
                   const int* f_adjs, const int* f_ids);
/* The mesh adjacency arrays mb_set_tetmesh takes, computed from the tet indices alone: the sparse counterpart of
 * load_tet_adj_info (src/IO/IO_CXX/io.cxx:238-335), whose dense n_vert (n_vert + 1) / 2 + 1 edge table (io.cxx:264)
 * is 2.6 GB at 36 k vertices and overflows int at 46 341.  Host code, no context, no GPU.
 *   v_adjs int[n_vert], e_adj6 int[6 * n_tet] (the e_adj6 form of mb_set_tetmesh), f_adjs / f_ids int[4 * n_tet] in
 *   tet_faces_lvid order; any output may be NULL.  f_ids: a boundary face gets boundary_sf_fids[k], k = its rank in
 *   (tet, local face) scan order (the surface-mesh facet id the reference finds through tet_vs2sf_fids, :306-313),
 *   or k itself when boundary_sf_fids is NULL; an interior face gets n_sf_facets + the rank of its first visit in
 *   scan order (:314-325).  *n_boundary_faces receives the number of boundary faces. */
int mb_tet_adjacency(const int* idx_aos, int n_tet, int n_vert, const int* boundary_sf_fids, int n_sf_facets, int* v_adjs,
                     int* e_adj6, int* f_adjs, int* f_ids, int* n_boundary_faces);
/* restrict subsequent mb_rpd3d calls to tets [first, first+count) (multi-GPU tet shards). */
int mb_set_tet_range(mb_ctx* ctx, int first, int count);
/* the tet ids stored in the result become (index in the uploaded arrays + base): a rank of a multi-GPU job
 * uploads only ITS contiguous shard of idx_aos / f_adjs / f_ids / e_adj6 (with the global vertex array, like
 * the partial-tet calls of rpd_api.cxx:254-281) and still returns global tet ids.  Reset by mb_set_tetmesh. */
int mb_set_tet_id_base(mb_ctx* ctx, int base);
/* restrict subsequent mb_rpd3d calls to the listed tets (strictly ascending ids of the resident mesh),
 * the device-side counterpart of load_partial_tet_given_spheres (rpd_api.cxx:482-535): a partial
 * recompute without re-uploading a sub-mesh; cells keep the GLOBAL tet id.  n = 0 clears the subset. */
int mb_set_tet_subset(mb_ctx* ctx, const int* tet_ids, int n);

/* ---------------------------------------------------------------- RPD */
typedef struct {
  int lanes_per_cell;   /* 0 = default; 8, 16 or 32 lanes cooperate on one (tet, site) cell */
  int grid_k;           /* grid-kNN mode: expected candidate sites per tet (0 = default 96): <=32 / <=96 /
                           >96 pick the fast-path list capacity 32 / 96 / 256; longer lists are redone
                           by a 2048-entry pass; at most 96 (256 if grid_k > 96) are kept per tet */
  int want_volumes;     /* accumulate per-site volume / barycentre sums                       */
  int grid_candidates;  /* given-neighbours mode only: 1 = find the candidate (tet, site) pairs with the
                           uniform-grid search instead of the reference's dense relation predicate
                           (voronoi.cu:154-193, O(n_tet*n_site)); each cell is still clipped by exactly
                           the listed neighbours in list order, so valid cells are byte-identical
                           whenever the lists contain every true power neighbour (regular-triangulation
                           lists do); only the set of empty no_intersection pairs differs */
  int lean_records;     /* streamed runs only: 0 = full compact records; 2 = SLIM: like 1, and the three id words per
                           plane shrink to one neighbour site id per bisector (tet-face ids and adjacency counts are
                           functions of the tet id): 16 + 4 nb_v + 4 (nb_p - 4) + 3 nb_e bytes, ~2.6x fewer than
                           full.  Bits 31 and 29 of record word 2 mark it.
                           (mb_rpd_run_to_host / _to_sink): 1 = the records travel WITHOUT their
                           plane equations (16 of ~28 bytes per plane, ~38 % of a record): a tet-face plane is a
                           function of (tet, face), a power bisector of (seed, neighbour) -- the ids stay.  Bit 31
                           of record word 2 marks the format; mb_rpd_fetch_records / mb_rpd_expand_compact
                           recompute the equations on the host with the reference's literal arithmetic
                           (tri2plane common_cuda.h:248-254, new_plane convex_cell.cu:561-592): expanded records
                           are byte-identical to the full format's */
  int security_radius;  /* given-neighbours mode only, for neighbour lists sorted by distance (the reference's kNN lists,
                           not its regular-triangulation lists): 1 = the security-radius early exit the live reference
                           comments out (is_security_radius_reached convex_cell.cu:240-268, used at :1285-1296): a cell
                           stops clipping at the first listed neighbour whose bisector lies beyond twice its farthest
                           vertex, and a cell whose LAST listed neighbour does not reach that radius is dropped as
                           security_radius_not_reached (:1304-1316; counted in mb_rpd_status_histogram).  grid-kNN
                           mode needs no such exit: its per-tet candidate lists already hold only the sites that can
                           reach the tet */
} mb_rpd_opts;

/* site_soa float[3*n_site] = x.. | y.. | z.. (rpd_api.cxx:363-365), site_w float[n_site] = r^2,
 * site_flags unsigned[n_site] (SiteFlag), site_knn int[(site_k+1)*n_site] row = slot, -1 padded
 * (triangulation.cxx:237-258).
 *   site_knn != NULL : given-neighbours mode = the reference's semantics: candidate (tet, site)
 *                      pairs by the tet_sphere_relations_dev predicate, each cell clipped by
 *                      exactly the listed neighbours in list order.
 *   site_knn == NULL : grid-kNN mode: uniform-grid per-tet nearest-sphere search; each cell is
 *                      clipped by the other candidates of its tet.
 * opts may be NULL.  *out receives a result handle (free with mb_rpd_free). */
int mb_rpd3d(mb_ctx* ctx, const float* site_soa, const float* site_w, const unsigned* site_flags,
             int n_site, const int* site_knn, int site_k, const mb_rpd_opts* opts,
             mb_rpd_result** out);

/* The same call split in three so that a caller can keep inputs resident and time / overlap the
 * phases: upload the sites (H2D), run all kernels on the context's stream, synchronise. */
int mb_rpd_upload_sites(mb_ctx* ctx, const float* site_soa, const float* site_w,
                        const unsigned* site_flags, int n_site, const int* site_knn, int site_k);
int mb_rpd_run(mb_ctx* ctx, const mb_rpd_opts* opts, mb_rpd_result** out); /* async launch */
int mb_rpd_sync(mb_ctx* ctx, mb_rpd_result* res); /* waits, reads back the counters */

/* Streamed variant of mb_rpd_run + mb_rpd_sync + mb_rpd_fetch_compact for callers that want the result in
 * HOST memory (what compute_clipped_voro_diagram_GPU returns, voronoi.cu:717-769): the processed tets are
 * cut into n_chunks contiguous spans (0 = automatic: ~40k tets each for the lean / slim transport formats, ~32k for
 * full records, first and last span half as long) and the ordered compact records of span c are copied to
 * library-owned pinned host memory on a second stream while span c+1 is clipped, so the PCIe transfer hides behind
 * the kernels.  With grid candidates the neighbour search runs once for all tets up front and only the clipping and
 * ordering are cut into spans.  On return the complete result -- identical to
 * the one-shot run's mb_rpd_fetch_compact output -- is at *host_blob / *host_cell_offsets (n_cells+1 byte
 * offsets), valid until the next streamed run on this context or mb_destroy.  The result handle carries the
 * counters (mb_rpd_count / _stats / _kernel_ms) and serves mb_rpd_fetch_records / mb_rpd_fetch_compact from
 * the host copy; it keeps nothing on the device (mb_rpd_emit, mb_rpd_device_buffers, mb_rpd_fetch_pairs
 * return MB_ERR_STATE). */
int mb_rpd_run_to_host(mb_ctx* ctx, const mb_rpd_opts* opts, int n_chunks, mb_rpd_result** out,
                       const void** host_blob, const long** host_cell_offsets);

/* Incremental recompute for the MATTopo loop (BASELINE configs[4]; replaces RPD3D_GPU::calculate_partial's ring
 * selection, rpd_api.cxx:147-313, triangulation.cxx:442-549, load_partial_tet_given_spheres :482-535, which need the
 * CGAL regular triangulation).  Call it after mb_rpd_upload_sites (grid-kNN mode, whole mesh) instead of mb_rpd_run:
 * the candidate lists of ALL tets are recomputed (K1 + K2) and compared with the previous incremental run's; only the
 * tets whose list changed or lists a changed site -- the AFFECTED tets, an exact set: every other tet's records are
 * bit-for-bit those of the previous run -- are clipped.  *out holds the records of the affected tets only (global tet
 * ids, (tet, site) order); the caller replaces those tets' records in its previous result (merge_convex_cells,
 * rpd_api.cxx:432-479).  The first call (or the first after mb_set_tetmesh / a change of grid_k) affects every tet.
 * to_host = 1 streams the records to pinned host memory like mb_rpd_run_to_host (host_blob / host_cell_offsets as
 * there); 0 keeps them on the device like mb_rpd_run.  Existing sites keep their ids across calls; new sites are
 * appended.  mb_rpd_fetch_affected_tets: the ascending ids of the tets of the last incremental run. */
int mb_rpd_run_incremental(mb_ctx* ctx, const mb_rpd_opts* opts, int to_host, int n_chunks, mb_rpd_result** out,
                           long* n_affected_tets, const void** host_blob, const long** host_cell_offsets);
int mb_rpd_fetch_affected_tets(mb_ctx* ctx, int* tet_ids);
/* host-side merge of compact records (no GPU, no context): prev = the caller's previous result, patch = the records of
 * an incremental run, affected_tets = its ascending tet ids.  out_offs needs n_prev + n_patch + 1 entries, out_blob
 * prev bytes + patch bytes (NULL: only count).  Works on full and transport-format (lean / slim) records alike, as
 * long as prev and patch use the same format. */
int mb_rpd_merge_compact(const void* prev_blob, const long* prev_offs, long n_prev, const void* patch_blob,
                         const long* patch_offs, long n_patch, const int* affected_tets, long n_affected, void* out_blob,
                         long* out_offs, long* n_out, long* out_bytes);

/* The same streamed run into CALLER memory of fixed capacity (MB_ERR_NOMEM if it does not fit): sink_blob /
 * sink_cell_offsets may be pinned or registered host memory (mb_host_register: e.g. a shared-memory segment
 * that every rank of a multi-GPU job writes its tet shard into, all PCIe links in parallel), device memory of
 * this GPU, or device memory of a PEER GPU opened with mb_sink_open.  In the last case span c crosses NVLink
 * by copy-engine DMA while span c+1 is clipped: the multi-GPU gather of SURVEY 8e happens inside the run
 * instead of as a collective after it.  Offsets start at 0 (rebase by the preceding shards' sizes). */
int mb_rpd_run_to_sink(mb_ctx* ctx, const mb_rpd_opts* opts, int n_chunks, void* sink_blob, size_t sink_cap_bytes,
                       long* sink_cell_offsets, size_t sink_cap_cells, mb_rpd_result** out);
/* sink memory: a device buffer on this context's GPU that other processes / GPUs of the box can write into
 * (64-byte CUDA IPC handle, ship it to the peers by any means), the peer-side mapping, and page-locking of
 * caller host memory.  mb_copy_to_host: blocking D2H of any device-visible range (sink readout). */
int mb_sink_create(mb_ctx* ctx, size_t bytes, void** d_ptr, unsigned char ipc_handle[64]);
int mb_sink_destroy(mb_ctx* ctx, void* d_ptr);
int mb_sink_open(mb_ctx* ctx, const unsigned char ipc_handle[64], void** d_peer_ptr);
int mb_sink_close(mb_ctx* ctx, void* d_peer_ptr);
int mb_host_register(mb_ctx* ctx, void* host_ptr, size_t bytes);
int mb_host_unregister(mb_ctx* ctx, void* host_ptr);
int mb_copy_to_host(mb_ctx* ctx, void* host_dst, const void* d_src, size_t bytes);

/* expand n_cells compact records (full or lean format, as delivered by a streamed run or gathered from several
 * ranks' sinks) into n_cells * MB_RECORD_BYTES ConvexCellTransfer records; id = first_id + index.  The context
 * must hold the tet mesh and the sites the records were computed from. */
int mb_rpd_expand_compact(mb_ctx* ctx, const void* blob, const long* cell_offsets, long n_cells, long first_id, void* dst);

/* number of tet spans the run was cut into (1 for mb_rpd_run); negative mb_status on a NULL handle */
int mb_rpd_spans(const mb_rpd_result* res);

void mb_rpd_free(mb_rpd_result* res);

/* number of valid cells (status success), candidate pairs clipped, clip_by_plane calls */
int mb_rpd_count(const mb_rpd_result* res, long* n_cells, long* n_pairs, long* n_clips);
/* stats[8]: [0] cells  [1] candidate pairs  [2] clip_by_plane calls that reached the exact predicate
 * [3] listed neighbours rejected by the conservative bounding filter (the reference would have
 * clipped and popped them)  [4] tets whose candidate list overflowed grid_k (grid mode)
 * (truncated: should be 0)  [5] compact result bytes  [6] tets redone by the big-list candidate pass
 * [7] conflict tests the FP32 filter could not decide (evaluated with the FP64 determinant) */
int mb_rpd_stats(const mb_rpd_result* res, long stats[8]);
/* The flagged class: cells / candidate pairs with at least one conflict test whose FP64 determinant fell
 * under the static-filter bound of src/predicate_generator (main.cpp:52-78),
 *     |det| < 1.2466136531027298e-13 * maxx * maxy * maxz * max(maxx, maxy, maxz)^2
 * -- exactly the test of the reference's USE_ARITHMETIC_FILTER branch (convex_cell.cu:479-497), which drops such a
 * cell as needs_exact_predicates.  The live reference compiles that branch out (voronoi_common.h:32) and so does
 * this library: the decision stays the FP64 one, the cell is kept, and it carries the flag (bit 30 of word 2 of its
 * compact record).  Combinatorial differences against another clip order are only legitimate on flagged cells.
 *   mb_rpd_flagged      counts
 *   mb_rpd_fetch_flags  cell_flag[n_cells] (order of the records) and / or pair_flag[n_pairs] (order of
 *                       mb_rpd_fetch_pairs; device-resident results only); either may be NULL */
int mb_rpd_flagged(const mb_rpd_result* res, long* n_flagged_cells, long* n_flagged_pairs);
int mb_rpd_fetch_flags(mb_rpd_result* res, unsigned char* cell_flag, unsigned char* pair_flag);
/* test hook: overrides the pairs-per-tet estimate the speculative span launches size their arrays with (0 = learn
 * it again with a synchronising span); a too small value must only cost a redone span, never change a result */
int mb_debug_set_pair_hint(mb_ctx* ctx, double pairs_per_tet);
/* grid-kNN mode internals: cells that outgrew the compact caps of K3's first pass (48 planes / 72 vertices /
 * 120 edges) and were recomputed by the second pass at the reference's caps (64 / 96 / 152), and dead plane /
 * edge garbage collections */
int mb_rpd_clip_passes(const mb_rpd_result* res, long* n_second_pass_cells, long* n_garbage_collections);
/* int[10]: index = status+1 (early_return .. needs_perturb), over all candidate pairs */
int mb_rpd_status_histogram(const mb_rpd_result* res, long hist[10]);
/* kernel milliseconds of the last run: [0]=candidates (K1+K2) [1]=clip (K3) [2]=emit (K4)
 * [3]=total device time, measured with CUDA events on the context's stream */
int mb_rpd_kernel_ms(const mb_rpd_result* res, float ms[4]);

/* the candidate (tet, site) pairs the run clipped, sorted by (tet, site) -- the reference's tet_knn
 * (voronoi.cu:266-317) in pair form -- and the per-pair Status; n_pairs entries each, any may be
 * NULL.  Valid until the next mb_rpd_run on the same context. */
int mb_rpd_fetch_pairs(mb_rpd_result* res, int* pair_tet, int* pair_site, signed char* pair_status);

/* cells sorted by (tet, site), dst = n_cells * MB_RECORD_BYTES in the ConvexCellTransfer
 * layout; only entries < nb_v/nb_p/nb_e are defined (others zero); id = index. */
int mb_rpd_fetch_records(mb_rpd_result* res, void* dst);
/* the compact form: bytes needed, then the blob + per-cell byte offsets (n_cells+1 longs). */
int mb_rpd_compact_bytes(const mb_rpd_result* res, long* n_bytes);
int mb_rpd_fetch_compact(mb_rpd_result* res, void* blob, long* cell_offsets);
/* per-site volume and barycentre sums (float[n_site], float[3*n_site] SoA); needs want_volumes */
int mb_rpd_site_volumes(mb_rpd_result* res, float* vol, float* bary_sum_soa);
/* per-cell volume (the value the reference accumulates per site, convex_cell.cu:1008-1069); float[n_cells] */
int mb_rpd_cell_volumes(mb_rpd_result* res, float* cell_vol);
/* raw device pointers of the compact result (for NCCL gathers): blob, cell_offsets(long) */
int mb_rpd_device_buffers(mb_rpd_result* res, void** d_blob, long* n_bytes, void** d_offsets,
                          long* n_cells);

/* K4 emission (get_all_voro_info, rpd_update.cxx:112-301): counts, then SoA arrays.
 *   facets   : cell_id, key (neighbour site id, or tet-face id), is_tet_face
 *   vertices : cell_id, lvid, key[3] (sorted neighbour ids, -1 = surface), pos[3], surf_fid
 *   edges    : cell_id, key[2] (sorted neighbour ids), end vertices lvid[2]
 * max_surf_fid = sf_mesh.facets.nb()-1 (rpd_update.cxx:606). */
typedef struct {
  long n_facets, n_vertices, n_edges;
} mb_emit_counts;
int mb_rpd_emit(mb_rpd_result* res, int max_surf_fid, mb_emit_counts* counts);
int mb_rpd_fetch_emit(mb_rpd_result* res, int* facet_cell, int* facet_key, unsigned char* facet_is_tet,
                      int* vert_cell, int* vert_lvid, int* vert_key3, float* vert_pos3,
                      int* vert_surf_fid, int* edge_cell, int* edge_key2, int* edge_lvid2,
                      float* cell_euler);

/* The rest of get_all_voro_info's per-cell emission:
 *   mb_rpd_fetch_facet_centroids  float[3 * n_facets]: for a tet-face facet with id <= max_surf_fid the pc_face centroid
 *       of cell_to_surfv2fid (get_cell_v2surffid, rpd_update.cxx:20-42, 129-133: mean of the face loop's vertices in
 *       the loop order of reload_pc_explicit, float); zero for every other facet.  Needs mb_rpd_emit.
 *   mb_set_feature_edges  TetMesh::tet_es2fe_map for the resident mesh: n rows (tet, lf_min, lf_max, fe_type, fe_id,
 *       fe_line_id), lf = local face 0..3, fe_type 1 = sharp (SE) / 2 = concave (CE) (input_types.h:13-17).  Reset by
 *       mb_set_tetmesh; n = 0 clears.  Subsequent mb_rpd_emit calls then also produce the covered feature edges
 *       (rpd_update.cxx:209-259):
 *   mb_rpd_feature_edge_count / mb_rpd_fetch_feature_edges
 *       hit6   int[6 * n_hits]  (cell, fe_type, lv1 < lv2, fe_line_id, fe_id)   = se_covered_lvids / ce_covered_lvids
 *       end4   int[8 * n_hits]  two rows (cell, lvid, neigh, fe_line_id) per hit: the end vertex's FIRST half-plane
 *                               neighbour (se_line_endpos key), neigh = -1 when the vertex lies on no half-plane
 *       end_pos3 float[6 * n_hits]  the end vertices' positions (zero where neigh = -1) */
int mb_rpd_fetch_facet_centroids(mb_rpd_result* res, float* centroid3);
int mb_set_feature_edges(mb_ctx* ctx, const int* rows6, long n);
int mb_rpd_feature_edge_count(const mb_rpd_result* res, long* n_hits);
int mb_rpd_fetch_feature_edges(mb_rpd_result* res, int* hit6, int* end4, float* end_pos3);

/* K6 power-cell topology summary = the rest of update_power_cells after get_all_voro_info
 * (rpd_update.cxx:303-316 cell_neighbors, :497-503 cc_cells, :439-470 facet_cc_cells) and the sums
 * check_cc_and_euler consumes (fix_topo.cxx:81-144, 150-235, 310-380); needs mb_rpd_emit first.
 *   cell_cc[n_cells]     smallest cell id of the cell's connected component inside its power cell (two cells
 *                        of a power cell are neighbours iff they share a tet-face id)
 *   facet_cc[n_facets]   for a half-plane facet (site, neigh) of a cell: smallest FACET index of its component
 *                        among the cells carrying that half-plane; -1 for tet-face facets
 *   edge_cc[n_edges]     for an emitted edge (between two half-planes, key = sorted neighbour pair) of a cell:
 *                        smallest EDGE index of its component among the cells carrying that key
 *                        (edge_cc_cells, update_pc_edge_cc_info rpd_update.cxx:507-521)
 *   site_n_cells / site_n_cc / site_euler_sum [n_site]   cells, cell components, double sum of the per-cell
 *                        Euler values in ascending cell id (euler = sum - n_cells, fix_topo.cxx:128-131)
 *   pair_site / pair_neigh / pair_n_cc [n_halfplane_pairs]   one entry per half-plane, sorted by (site, neigh):
 *                        number of components of that facet (is_to_fix_facet_cc: > 1 needs fixing) */
typedef struct {
  long n_cells, n_facets, n_sites, n_halfplane_pairs, n_edges;
} mb_topo_counts;
int mb_rpd_topology(mb_rpd_result* res, mb_topo_counts* counts);
int mb_rpd_fetch_topology(mb_rpd_result* res, int* cell_cc, int* facet_cc, int* edge_cc, int* site_n_cells, int* site_n_cc,
                          double* site_euler_sum, int* pair_site, int* pair_neigh, int* pair_n_cc);

/* The IO_CUDA result format: Houdini .bgeo V5 (big-endian) of the cell polygons, byte-identical to the
 * reference's save_convex_cells_houdini (src/IO/IO_CUDA/io_cuda.cxx:152-187 with is_slice_plane = false;
 * facet loops io_cuda.cxx:21-148; container io_utils.cpp:106-231): every vertex of every cell as a point,
 * one polygon per active facet (all of them, or with is_boundary_only only half-planes and surface faces
 * with id < max_sf_fid), primitive attribute "PrimAttr" = site id; 16-bit point indices up to 65 536 points.
 * `path` is the file to write (the reference derives "../out/<name>/rpd/rpd_<name>_<timestamp>.bgeo").
 *   mb_rpd_write_bgeo      from compact records in host memory (full or lean; ctx = the context they came from)
 *   mb_bgeo_write_records  from records in the ConvexCellTransfer layout (no context, no GPU needed) */
int mb_rpd_write_bgeo(mb_ctx* ctx, const void* blob, const long* cell_offsets, long n_cells, int max_sf_fid,
                      int is_boundary_only, const char* path, long* n_points, long* n_polygons);
int mb_bgeo_write_records(const void* records, long n_cells, int max_sf_fid, int is_boundary_only, const char* path,
                          long* n_points, long* n_polygons);

/* ---------------------------------------------------------------- dist2mat */
/* spheres float[4*n_sph] = (cx,cy,cz,r) -- r, not r^2 (fix_geo_error.cxx:324-328);
 * samples float[3*n_samples]; offset/count unsigned[n_samples]; prims int[3*n_prims]:
 * (-1,-1,s) sphere, (-1,a,b) cone, (a,b,c) slab (dist2mat.cu:233-246).
 * result float[n_samples], closest_id int[n_samples] = index within the sample's list
 * (empty list -> 1e16f, -1).  tie_flag (nullable) unsigned char[n_samples]: 1 where the two best
 * distances differ by < 1e-6 relative to max(|distance|, |sample coordinates|) (the flagged-tie class: closer than
 * the float resolution of |p - c| - r, where the FMA-contracted device build and an IEEE evaluation of the
 * reference's expressions can rank them differently).
 * Arithmetic: the reference's expressions evaluated like its DEVICE build (nvcc FMA contraction, clamp(t,0,1) as a
 * saturate: NaN -> 0): bit-identical to the reference's own CUDA kernel on 99.97 % of 200 000 samples, the rest
 * within 1e-6 (tests/test_gpu_reference_build.py). */
int mb_dist2mat(mb_ctx* ctx, const float* spheres, int n_sph, const float* samples, int n_samples,
                const unsigned* offset, const unsigned* count, const int* prims, long n_prims,
                float* result, int* closest_id, unsigned char* tie_flag);
/* resident variant: upload once, run many times (device-only timing), fetch */
int mb_dist2mat_upload(mb_ctx* ctx, const float* spheres, int n_sph, const float* samples,
                       int n_samples, const unsigned* offset, const unsigned* count,
                       const int* prims, long n_prims);
int mb_dist2mat_run(mb_ctx* ctx, float* kernel_ms);
int mb_dist2mat_fetch(mb_ctx* ctx, float* result, int* closest_id, unsigned char* tie_flag);

/* ---------------------------------------------------------------- dist2mat with device-built candidate lists
 * Replaces, for callers that have the medial mesh and the samples' surface-face ids, the CPU list construction and
 * the replicated upload of the reference (gather_point_to_sites fix_geo_error.cxx:149-178, gather_point_to_slab_and_cone
 * :180-215, load_and_compute_sample_dist2mat_gpubuffer :300-366: one private int3 list per sample, ~290 B / sample):
 * the per-surface-face lists are built once on the device and a sample costs 16 bytes (position + face id).
 *   mb_dist2mat_set_medial_mesh   spheres (cx,cy,cz,r) + medial faces (3 sphere ids each, MedialMesh::faces) + medial
 *                                 edges (2 sphere ids each, MedialMesh::edges); a sphere's faces_ / edges_ sets are the
 *                                 faces / edges that list it
 *   mb_dist2mat_set_face_sites    (surface fid, site) rows in any order, duplicates allowed = the union of every power
 *                                 cell's cell_to_surfv2fid; n_fid = number of surface faces.  Site ids index spheres.
 *   mb_dist2mat_set_face_sites_from_rpd   the same, taken on the device from an RPD result after mb_rpd_emit (K4's
 *                                 surface facets): RPD -> K4 -> dist2mat lists without a host round trip
 *   mb_dist2mat_by_face           samples + their surface-face id -> result / closest_id exactly as mb_dist2mat would
 *                                 return them for the reference's lists (per site in ascending id: its slabs in
 *                                 ascending face id, its cones in ascending edge id, the sphere; no de-duplication;
 *                                 a face no cell touches has an empty list -> 1e16f, -1), and closest_prim3 (nullable):
 *                                 the winning primitive's int3 (samples_clostprim, fix_geo_error.cxx:368-380)
 *   mb_dist2mat_upload_by_face / mb_dist2mat_run / mb_dist2mat_fetch: the resident split of the same call
 *   mb_dist2mat_fetch_face_lists  the device-built CSR (n_fid + 1 offsets in prims, then the int3 prims) for inspection */
int mb_dist2mat_set_medial_mesh(mb_ctx* ctx, const float* spheres, int n_sph, const int* mm_faces, int n_faces,
                                const int* mm_edges, int n_edges);
int mb_dist2mat_set_face_sites(mb_ctx* ctx, const int* fid_site_rows, long n_rows, int n_fid);
int mb_dist2mat_set_face_sites_from_rpd(mb_ctx* ctx, mb_rpd_result* res, int max_surf_fid);
int mb_dist2mat_upload_by_face(mb_ctx* ctx, const float* samples, const int* sample_fid, int n_samples);
int mb_dist2mat_by_face(mb_ctx* ctx, const float* samples, const int* sample_fid, int n_samples, float* result,
                        int* closest_id, int* closest_prim3, unsigned char* tie_flag);
int mb_dist2mat_fetch_closest_prims(mb_ctx* ctx, int* closest_prim3);
int mb_dist2mat_face_list_size(mb_ctx* ctx, long* n_fid, long* n_prims);
int mb_dist2mat_fetch_face_lists(mb_ctx* ctx, long long* list_off, int* prims3);

#ifdef __cplusplus
}
#endif
#endif /* LIBMAT_B200_H */
