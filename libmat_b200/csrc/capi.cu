// extern "C" entry points of libmat_b200 (see include/libmat_b200.h).  Every call is wrapped so that
// no exception and no exit() crosses the boundary (the reference exits the process on any CUDA
// error: include/common_cuda.h:6-7, src/rpd3d/voronoi.cu:31-37, include/cuda_utils.h:18-25).
#include <omp.h>

#include <algorithm>
#include <cstdlib>
#include <new>

#include "mb_internal.h"
#include "rpd_device.cuh"

#define MB_TRY(ctx_) \
  mb_ctx* ctx__ = (ctx_); \
  try {
#define MB_CATCH                                       \
  }                                                    \
  catch (const MbError& e) {                           \
    if (ctx__) ctx__->err = e.msg;                     \
    return e.code;                                     \
  }                                                    \
  catch (const std::bad_alloc&) {                      \
    if (ctx__) ctx__->err = "host allocation failed";  \
    return MB_ERR_NOMEM;                               \
  }                                                    \
  catch (...) {                                        \
    if (ctx__) ctx__->err = "unknown error";           \
    return MB_ERR_CUDA;                                \
  }                                                    \
  return MB_OK;

extern "C" {

const char* mb_version(void) { return "libmat_b200 0.1 (sm_100a)"; }

void mb_predicate_bounds(double* bound_f64, float* bound_f32) {
  if (bound_f64) *bound_f64 = MB_FILTER_BOUND_F64;
  if (bound_f32) *bound_f32 = MB_FILTER_BOUND_F32;
}

mb_ctx* mb_create(int device, int* err) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    if (err) *err = MB_ERR_NODEVICE;
    return nullptr;  // no CPU fallback
  }
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) device = 0;
  }
  if (device >= ndev || cudaSetDevice(device) != cudaSuccess) {
    if (err) *err = MB_ERR_ARG;
    return nullptr;
  }
  mb_ctx* ctx = new (std::nothrow) mb_ctx();
  if (!ctx) {
    if (err) *err = MB_ERR_NOMEM;
    return nullptr;
  }
  ctx->device = device;
  if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete ctx;
    if (err) *err = MB_ERR_CUDA;
    return nullptr;
  }
  ctx->stream = ctx->own_stream;
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (const char* v = getenv("MB_D2M_VARIANT")) ctx->d2m_variant = atoi(v);
  if (const char* v = getenv("MB_NO_CULL")) ctx->no_cull = atoi(v) != 0;
  if (const char* v = getenv("MB_CLIP_VARIANT")) ctx->clip_variant = atoi(v);
  if (const char* v = getenv("MB_K2_VARIANT")) ctx->k2_variant = atoi(v);
  if (const char* v = getenv("MB_STREAM_VARIANT")) ctx->stream_variant = atoi(v);
  if (const char* v = getenv("MB_DEBUG_SMALL_SCRATCH")) ctx->debug_small_scratch = atoi(v);
  if (const char* v = getenv("MB_TRACE")) {
    ctx->trace_level = atoi(v);
    ctx->trace_on = ctx->trace_level != 0;
  }
  if (err) *err = MB_OK;
  return ctx;
}

void mb_destroy(mb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->trace_level == 1)
    fprintf(stderr, "[libmat_b200 trace] host us: launchK2 %.0f waitK2 %.0f launchK3+scans %.0f waitK3 %.0f launchGather %.0f\n",
            ctx->trace_us[0], ctx->trace_us[1], ctx->trace_us[2], ctx->trace_us[3], ctx->trace_us[6]);
  for (mb_rpd_result* r : ctx->live_results) r->ctx = nullptr;  // results outlive the context safely
  ctx->spare_blob.release();
  ctx->spare_cell_off.release();
  TetMeshDev& M = ctx->mesh;
  M.fe_table.release(); M.fe_rows.release();
  M.vert4.release(); M.tet_idx.release(); M.tet_fadj.release(); M.tet_fid.release(); M.tet_e6.release(); M.tet_geo.release(); M.tet_vadj.release(); M.tet_sel.release();
  SitesDev& S = ctx->sites;
  S.site4.release(); S.flags.release(); S.nbr.release(); S.knn_staging.release(); S.soa_staging.release();
  D2MDev& D = ctx->d2m;
  D.spheres.release(); D.samples.release(); D.offset.release(); D.count.release(); D.prims.release();
  D.result.release(); D.closest.release(); D.tie.release();
  ctx->tet_cnt.release(); ctx->tet_off.release(); ctx->pair_tet.release(); ctx->pair_site.release(); ctx->pair_local.release();
  ctx->cand_pad.release(); ctx->redo_list.release(); ctx->cand_cnt.release(); ctx->ovf_list.release(); ctx->word_off.release(); ctx->pair_valid.release();
  ctx->pair_cell.release(); ctx->pair_status.release(); ctx->pair_blob.release(); ctx->pair_words.release();
  ctx->scratch.release(); ctx->counters.release(); ctx->cub_tmp.release();
  ctx->grid_cnt.release(); ctx->grid_off.release(); ctx->grid_sorted_id.release(); ctx->grid_cell_of.release();
  ctx->grid_site4.release(); ctx->grid_wmax0.release(); ctx->grid_wmax1.release();
  ctx->pin_in.release(); ctx->pin_out.release(); ctx->pin_blob.release(); ctx->pin_off.release();
  for (int b = 0; b < 2; b++) {
    ctx->span_blob[b].release();
    ctx->span_off[b].release();
    if (ctx->ev_gathered[b]) cudaEventDestroy(ctx->ev_gathered[b]);
    if (ctx->ev_copied[b]) cudaEventDestroy(ctx->ev_copied[b]);
  }
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->hs) cudaFreeHost(ctx->hs);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

int mb_set_stream(mb_ctx* ctx, void* cuda_stream) {
  if (!ctx) return MB_ERR_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
  return MB_OK;
}

const char* mb_last_error(const mb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int mb_set_tetmesh(mb_ctx* ctx, const float* verts_aos, int n_vert, const int* idx_aos, int n_tet,
                   const int* v_adjs, const int* e_adjs_dense, const int* e_adj6,
                   const int* f_adjs, const int* f_ids) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx, MB_ERR_ARG, "null context");
  MB_REQUIRE(verts_aos && idx_aos && v_adjs && f_adjs && f_ids, MB_ERR_ARG, "null mesh array");
  MB_REQUIRE(n_vert > 0 && n_tet > 0, MB_ERR_ARG, "empty mesh");
  MB_REQUIRE((e_adjs_dense != nullptr) != (e_adj6 != nullptr), MB_ERR_ARG,
             "exactly one of e_adjs_dense / e_adj6 must be given");
  MB_CUDA(cudaSetDevice(ctx->device));
  rpd_upload_mesh(ctx, verts_aos, n_vert, idx_aos, n_tet, v_adjs, e_adjs_dense, e_adj6, f_adjs, f_ids);
  MB_CATCH
}

int mb_set_tet_range(mb_ctx* ctx, int first, int count) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx, MB_ERR_ARG, "null context");
  MB_REQUIRE(first >= 0 && (count < 0 || first + count <= ctx->mesh.n_tet), MB_ERR_ARG, "bad tet range");
  ctx->mesh.range_first = first;
  ctx->mesh.range_count = count;
  MB_CATCH
}

int mb_set_tet_id_base(mb_ctx* ctx, int base) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx, MB_ERR_ARG, "null context");
  MB_REQUIRE(ctx->mesh.n_tet > 0, MB_ERR_STATE, "mb_set_tetmesh first (it resets the base to 0)");
  ctx->mesh.tet_id_base = base;
  MB_CATCH
}

int mb_set_tet_subset(mb_ctx* ctx, const int* tet_ids, int n) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx, MB_ERR_ARG, "null context");
  MB_REQUIRE(n >= 0 && (n == 0 || tet_ids), MB_ERR_ARG, "bad subset");
  TetMeshDev& M = ctx->mesh;
  for (int i = 0; i < n; i++)
    MB_REQUIRE(tet_ids[i] >= 0 && tet_ids[i] < M.n_tet && (i == 0 || tet_ids[i] > tet_ids[i - 1]), MB_ERR_ARG,
               "tet subset must be strictly ascending ids of the resident mesh");
  MB_CUDA(cudaSetDevice(ctx->device));
  if (n > 0) {
    M.tet_sel.reserve((size_t)n);
    MB_CUDA(cudaMemcpyAsync(M.tet_sel.p, tet_ids, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  M.n_sel = n;
  MB_CATCH
}

int mb_rpd_upload_sites(mb_ctx* ctx, const float* site_soa, const float* site_w,
                        const unsigned* site_flags, int n_site, const int* site_knn, int site_k) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx, MB_ERR_ARG, "null context");
  MB_REQUIRE(site_soa && site_w && site_flags && n_site > 0, MB_ERR_ARG, "bad site arrays");
  MB_CUDA(cudaSetDevice(ctx->device));
  rpd_upload_sites(ctx, site_soa, site_w, site_flags, n_site, site_knn, site_k);
  MB_CATCH
}

int mb_rpd_run(mb_ctx* ctx, const mb_rpd_opts* opts, mb_rpd_result** out) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx && out, MB_ERR_ARG, "null argument");
  MB_CUDA(cudaSetDevice(ctx->device));
  mb_rpd_result* res = new mb_rpd_result();
  *out = res;
  res->ctx = ctx;
  ctx->live_results.push_back(res);
  ctx->spare_blob.move_to(res->blob);
  ctx->spare_cell_off.move_to(res->cell_off);
  try {
    rpd_run(ctx, opts, res);
  } catch (...) {
    mb_rpd_free(res);
    *out = nullptr;
    throw;
  }
  MB_CATCH
}

int mb_rpd_sync(mb_ctx* ctx, mb_rpd_result* res) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx && res, MB_ERR_ARG, "null argument");
  rpd_sync(ctx, res);
  MB_CATCH
}

static void run_streamed(mb_ctx* ctx, const mb_rpd_opts* opts, int n_chunks, mb_rpd_result** out, void* dst_blob,
                         size_t cap_bytes, long long* dst_off, size_t cap_cells) {
  MB_REQUIRE(ctx && out, MB_ERR_ARG, "null argument");
  MB_CUDA(cudaSetDevice(ctx->device));
  mb_rpd_result* res = new mb_rpd_result();
  *out = res;
  res->ctx = ctx;
  ctx->live_results.push_back(res);
  try {
    rpd_run_to_host(ctx, opts, n_chunks, res, dst_blob, cap_bytes, dst_off, cap_cells);
    rpd_sync(ctx, res);
  } catch (...) {
    mb_rpd_free(res);
    *out = nullptr;
    throw;
  }
}

int mb_rpd_run_to_host(mb_ctx* ctx, const mb_rpd_opts* opts, int n_chunks, mb_rpd_result** out,
                       const void** host_blob, const long** host_cell_offsets) {
  MB_TRY(ctx)
  run_streamed(ctx, opts, n_chunks, out, nullptr, 0, nullptr, 0);
  if (host_blob) *host_blob = (*out)->host_blob;
  if (host_cell_offsets) *host_cell_offsets = reinterpret_cast<const long*>((*out)->host_off);
  MB_CATCH
}

int mb_rpd_run_incremental(mb_ctx* ctx, const mb_rpd_opts* opts, int to_host, int n_chunks, mb_rpd_result** out,
                           long* n_affected_tets, const void** host_blob, const long** host_cell_offsets) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx && out, MB_ERR_ARG, "null argument");
  MB_CUDA(cudaSetDevice(ctx->device));
  const int n_aff = rpd_incremental_select(ctx, opts);  // leaves the affected tets as the context's subset
  if (n_affected_tets) *n_affected_tets = n_aff;
  TetMeshDev& M = ctx->mesh;
  if (n_aff == 0) {  // nothing changed: an empty result (n_sel = 0 would mean "all tets")
    M.range_first = 0;
    M.range_count = 0;
  }
  struct Restore {  // whatever happens, later runs see the whole mesh again
    TetMeshDev& M;
    ~Restore() {
      M.n_sel = 0;
      M.range_first = 0;
      M.range_count = -1;
    }
  } restore{M};
  if (to_host) {
    run_streamed(ctx, opts, n_chunks, out, nullptr, 0, nullptr, 0);
    if (host_blob) *host_blob = (*out)->host_blob;
    if (host_cell_offsets) *host_cell_offsets = reinterpret_cast<const long*>((*out)->host_off);
  } else {
    mb_rpd_result* res = new mb_rpd_result();
    *out = res;
    res->ctx = ctx;
    ctx->live_results.push_back(res);
    ctx->spare_blob.move_to(res->blob);
    ctx->spare_cell_off.move_to(res->cell_off);
    try {
      rpd_run(ctx, opts, res);
      rpd_sync(ctx, res);
    } catch (...) {
      mb_rpd_free(res);
      *out = nullptr;
      throw;
    }
  }
  MB_CATCH
}

// Host-side counterpart of merge_convex_cells (rpd_api.cxx:432-479) on compact records: the previous result with the
// records of the affected tets replaced by the patch's.  Both inputs are in (tet, site) order, so is the output.
int mb_rpd_merge_compact(const void* prev_blob, const long* prev_offs, long n_prev, const void* patch_blob,
                         const long* patch_offs, long n_patch, const int* affected_tets, long n_affected, void* out_blob,
                         long* out_offs, long* n_out, long* out_bytes) {
  if (!prev_offs || !patch_offs || !out_offs || (n_prev && !prev_blob) || (n_patch && !patch_blob) || (n_affected && !affected_tets))
    return MB_ERR_ARG;
  const unsigned char* pb = static_cast<const unsigned char*>(prev_blob);
  const unsigned char* qb = static_cast<const unsigned char*>(patch_blob);
  unsigned char* ob = static_cast<unsigned char*>(out_blob);
  auto tet_of = [](const unsigned char* b, long off) {
    int t;
    memcpy(&t, b + off, 4);
    return t;
  };
  long i = 0, j = 0, a = 0, n = 0, at = 0;
  out_offs[0] = 0;
  while (i < n_prev || j < n_patch) {
    const int tp = i < n_prev ? tet_of(pb, prev_offs[i]) : 0x7fffffff;
    const int tq = j < n_patch ? tet_of(qb, patch_offs[j]) : 0x7fffffff;
    while (a < n_affected && affected_tets[a] < tp) a++;
    if (i < n_prev && a < n_affected && affected_tets[a] == tp) {  // superseded record: skip the whole tet run
      i++;
      continue;
    }
    if (tq <= tp && j < n_patch) {  // the patch's tets are affected ones: they never tie with a surviving record
      long j1 = j;
      while (j1 < n_patch && tet_of(qb, patch_offs[j1]) == tq) j1++;
      const long bytes = patch_offs[j1] - patch_offs[j];
      if (ob) memcpy(ob + at, qb + patch_offs[j], (size_t)bytes);
      for (long k = j; k < j1; k++) out_offs[++n] = at + (patch_offs[k + 1] - patch_offs[j]);
      at += bytes;
      j = j1;
    } else {
      // a run of surviving records: up to the next affected tet / the next patch tet
      long i1 = i;
      const int stop = std::min(a < n_affected ? affected_tets[a] : 0x7fffffff, tq);
      while (i1 < n_prev && tet_of(pb, prev_offs[i1]) < stop) i1++;
      if (i1 == i) i1 = i + 1;
      const long bytes = prev_offs[i1] - prev_offs[i];
      if (ob) memcpy(ob + at, pb + prev_offs[i], (size_t)bytes);
      for (long k = i; k < i1; k++) out_offs[++n] = at + (prev_offs[k + 1] - prev_offs[i]);
      at += bytes;
      i = i1;
    }
  }
  if (n_out) *n_out = n;
  if (out_bytes) *out_bytes = at;
  return MB_OK;
}

int mb_rpd_fetch_affected_tets(mb_ctx* ctx, int* tet_ids) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx && tet_ids, MB_ERR_ARG, "null argument");
  MB_REQUIRE(ctx->inc_valid, MB_ERR_STATE, "mb_rpd_run_incremental first");
  MB_CUDA(cudaSetDevice(ctx->device));
  const int n = ctx->inc_n_affected;
  if (n == ctx->inc_n_tet) {
    for (int i = 0; i < n; i++) tet_ids[i] = i;  // first run / after a reset: every tet
  } else if (n > 0) {
    MB_CUDA(cudaMemcpyAsync(tet_ids, ctx->inc_affected.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    MB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  MB_CATCH
}

int mb_rpd_run_to_sink(mb_ctx* ctx, const mb_rpd_opts* opts, int n_chunks, void* sink_blob, size_t sink_cap_bytes,
                       long* sink_cell_offsets, size_t sink_cap_cells, mb_rpd_result** out) {
  MB_TRY(ctx)
  MB_REQUIRE(sink_blob && sink_cell_offsets, MB_ERR_ARG, "null sink");
  run_streamed(ctx, opts, n_chunks, out, sink_blob, sink_cap_bytes, reinterpret_cast<long long*>(sink_cell_offsets),
               sink_cap_cells);
  MB_CATCH
}

// ---- memory a sink can live in ---------------------------------------------------------------------------
int mb_sink_create(mb_ctx* ctx, size_t bytes, void** d_ptr, unsigned char ipc_handle[64]) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx && d_ptr && bytes > 0, MB_ERR_ARG, "bad sink arguments");
  MB_CUDA(cudaSetDevice(ctx->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  MB_CUDA(cudaMalloc(&p, bytes));
  if (ipc_handle) {
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
      cudaFree(p);
      throw MbError{MB_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)};
    }
    memcpy(ipc_handle, &h, 64);
  }
  *d_ptr = p;
  MB_CATCH
}

int mb_sink_destroy(mb_ctx* ctx, void* d_ptr) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx, MB_ERR_ARG, "null context");
  MB_CUDA(cudaSetDevice(ctx->device));
  if (d_ptr) MB_CUDA(cudaFree(d_ptr));
  MB_CATCH
}

int mb_sink_open(mb_ctx* ctx, const unsigned char ipc_handle[64], void** d_peer_ptr) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx && ipc_handle && d_peer_ptr, MB_ERR_ARG, "null argument");
  MB_CUDA(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, 64);
  MB_CUDA(cudaIpcOpenMemHandle(d_peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  MB_CATCH
}

int mb_sink_close(mb_ctx* ctx, void* d_peer_ptr) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx, MB_ERR_ARG, "null context");
  MB_CUDA(cudaSetDevice(ctx->device));
  if (d_peer_ptr) MB_CUDA(cudaIpcCloseMemHandle(d_peer_ptr));
  MB_CATCH
}

int mb_host_register(mb_ctx* ctx, void* host_ptr, size_t bytes) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx && host_ptr && bytes > 0, MB_ERR_ARG, "bad arguments");
  MB_CUDA(cudaSetDevice(ctx->device));
  MB_CUDA(cudaHostRegister(host_ptr, bytes, cudaHostRegisterPortable));
  MB_CATCH
}

int mb_host_unregister(mb_ctx* ctx, void* host_ptr) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx && host_ptr, MB_ERR_ARG, "bad arguments");
  MB_CUDA(cudaSetDevice(ctx->device));
  MB_CUDA(cudaHostUnregister(host_ptr));
  MB_CATCH
}

int mb_copy_to_host(mb_ctx* ctx, void* host_dst, const void* d_src, size_t bytes) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx && host_dst && d_src, MB_ERR_ARG, "bad arguments");
  MB_CUDA(cudaSetDevice(ctx->device));
  MB_CUDA(cudaMemcpyAsync(host_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  MB_CUDA(cudaStreamSynchronize(ctx->stream));
  MB_CATCH
}

int mb_rpd3d(mb_ctx* ctx, const float* site_soa, const float* site_w, const unsigned* site_flags,
             int n_site, const int* site_knn, int site_k, const mb_rpd_opts* opts,
             mb_rpd_result** out) {
  int rc = mb_rpd_upload_sites(ctx, site_soa, site_w, site_flags, n_site, site_knn, site_k);
  if (rc) return rc;
  rc = mb_rpd_run(ctx, opts, out);
  if (rc) return rc;
  return mb_rpd_sync(ctx, *out);
}

void mb_rpd_free(mb_rpd_result* res) {
  if (!res) return;
  if (mb_ctx* ctx = res->ctx) {
    cudaSetDevice(ctx->device);
    for (size_t i = 0; i < ctx->live_results.size(); i++)
      if (ctx->live_results[i] == res) {
        ctx->live_results[i] = ctx->live_results.back();
        ctx->live_results.pop_back();
        break;
      }
    // park the two big buffers in the context for the next run (keeps the larger of the two)
    if (res->blob.cap > ctx->spare_blob.cap) {
      ctx->spare_blob.release();
      res->blob.move_to(ctx->spare_blob);
    }
    if (res->cell_off.cap > ctx->spare_cell_off.cap) {
      ctx->spare_cell_off.release();
      res->cell_off.move_to(ctx->spare_cell_off);
    }
  }
  res->blob.release(); res->cell_off.release(); res->site_vol.release(); res->site_bary.release(); res->cell_vol.release();
  res->f_cell.release(); res->f_key.release(); res->v_cell.release(); res->v_lvid.release();
  res->f_centroid3.release(); res->fe_hit6.release(); res->fe_end4.release(); res->fe_end_pos3.release();
  res->v_key3.release(); res->v_surf.release(); res->e_cell.release(); res->e_key2.release();
  res->e_lvid2.release(); res->f_istet.release(); res->v_pos3.release(); res->c_euler.release();
  res->t_cell_cc.release(); res->t_facet_cc.release(); res->t_edge_cc.release(); res->t_site_n_cells.release(); res->t_site_n_cc.release();
  res->t_pair_site.release(); res->t_pair_neigh.release(); res->t_pair_ncc.release(); res->t_site_euler.release();
  for (cudaEvent_t e : res->evs) {
    if (res->ctx)
      res->ctx->ev_pool.push_back(e);  // recycled by the next run
    else
      cudaEventDestroy(e);
  }
  delete res;
}

int mb_rpd_count(const mb_rpd_result* res, long* n_cells, long* n_pairs, long* n_clips) {
  if (!res) return MB_ERR_ARG;
  if (n_cells) *n_cells = res->n_cells;
  if (n_pairs) *n_pairs = res->n_pairs;
  if (n_clips) *n_clips = res->n_clips;
  return MB_OK;
}

int mb_rpd_spans(const mb_rpd_result* res) { return res ? res->n_spans : MB_ERR_ARG; }

int mb_rpd_stats(const mb_rpd_result* res, long stats[8]) {
  if (!res || !stats) return MB_ERR_ARG;
  stats[0] = res->n_cells;
  stats[1] = res->n_pairs;
  stats[2] = res->n_clips;
  stats[3] = res->n_culled;
  stats[4] = res->n_cand_overflow;
  stats[5] = res->compact_bytes;
  stats[6] = res->n_ovf_tets;
  stats[7] = res->n_exact;
  return MB_OK;
}

int mb_rpd_flagged(const mb_rpd_result* res, long* n_flagged_cells, long* n_flagged_pairs) {
  if (!res) return MB_ERR_ARG;
  if (n_flagged_cells) *n_flagged_cells = res->n_flag_cells;
  if (n_flagged_pairs) *n_flagged_pairs = res->n_flag_pairs;
  return MB_OK;
}

int mb_rpd_fetch_flags(mb_rpd_result* res, unsigned char* cell_flag, unsigned char* pair_flag) {
  mb_ctx* ctx = res ? res->ctx : nullptr;
  MB_TRY(ctx)
  MB_REQUIRE(res && ctx, MB_ERR_ARG, "null result");
  MB_CUDA(cudaSetDevice(ctx->device));
  if (res->host_only) {
    MB_REQUIRE(!pair_flag, MB_ERR_STATE, "pairs are not kept by a streamed run (mb_rpd_run_to_host)");
    MB_REQUIRE(res->host_blob, MB_ERR_STATE, "the result lives in a device sink");
    MB_REQUIRE(!res->sink_owned || res->generation == ctx->stream_generation, MB_ERR_STATE,
               "streamed result superseded by a later run");
    if (cell_flag)
      for (long i = 0; i < res->n_cells; i++)
        cell_flag[i] = (res->host_blob[res->host_off[i] / 4 + 2] & 0x40000000u) ? 1 : 0;
    return MB_OK;
  }
  rpd_fetch_flags(ctx, res, cell_flag, pair_flag);
  MB_CATCH
}

int mb_debug_set_pair_hint(mb_ctx* ctx, double pairs_per_tet) {
  if (!ctx) return MB_ERR_ARG;
  ctx->pairs_per_tet_hint = pairs_per_tet;
  return MB_OK;
}

int mb_rpd_clip_passes(const mb_rpd_result* res, long* n_second_pass_cells, long* n_garbage_collections) {
  if (!res) return MB_ERR_ARG;
  if (n_second_pass_cells) *n_second_pass_cells = res->n_redo;
  if (n_garbage_collections) *n_garbage_collections = res->n_gc;
  return MB_OK;
}

int mb_rpd_status_histogram(const mb_rpd_result* res, long hist[10]) {
  if (!res || !hist) return MB_ERR_ARG;
  for (int i = 0; i < 10; i++) hist[i] = res->hist[i];
  return MB_OK;
}

int mb_rpd_kernel_ms(const mb_rpd_result* res, float ms[4]) {
  if (!res || !ms) return MB_ERR_ARG;
  for (int i = 0; i < 4; i++) ms[i] = res->ms[i];
  return MB_OK;
}

int mb_tet_adjacency(const int* idx_aos, int n_tet, int n_vert, const int* boundary_sf_fids, int n_sf_facets, int* v_adjs,
                     int* e_adj6, int* f_adjs, int* f_ids, int* n_boundary_faces) {
  try {
    if (!idx_aos || n_tet <= 0 || n_vert <= 0) return MB_ERR_ARG;
    tet_adjacency(idx_aos, n_tet, n_vert, boundary_sf_fids, n_sf_facets, v_adjs, e_adj6, f_adjs, f_ids, n_boundary_faces);
  } catch (const MbError& e) {
    fprintf(stderr, "[libmat_b200] mb_tet_adjacency: %s\n", e.msg.c_str());
    return e.code;
  } catch (...) {
    return MB_ERR_NOMEM;
  }
  return MB_OK;
}

int mb_measure_peaks(mb_ctx* ctx, double* fp32_tflops, double* fp64_tflops) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx, MB_ERR_ARG, "null context");
  MB_CUDA(cudaSetDevice(ctx->device));
  peaks_measure(ctx, fp32_tflops, fp64_tflops);
  MB_CATCH
}

int mb_launch_count(const mb_ctx* ctx, unsigned long long* n_launches) {
  if (!ctx || !n_launches) return MB_ERR_ARG;
  *n_launches = ctx->n_launches;
  return MB_OK;
}

int mb_rpd_fetch_pairs(mb_rpd_result* res, int* pair_tet, int* pair_site, signed char* pair_status) {
  mb_ctx* ctx = res ? res->ctx : nullptr;
  MB_TRY(ctx)
  MB_REQUIRE(res && ctx, MB_ERR_ARG, "null result");
  MB_REQUIRE(!res->host_only, MB_ERR_STATE, "pairs are not kept by a streamed run (mb_rpd_run_to_host)");
  MB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const size_t n = (size_t)res->n_pairs;
  if (n > 0) {
    if (pair_tet) MB_CUDA(cudaMemcpyAsync(pair_tet, ctx->pair_tet.p, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
    if (pair_site) MB_CUDA(cudaMemcpyAsync(pair_site, ctx->pair_site.p, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
    if (pair_status) MB_CUDA(cudaMemcpyAsync(pair_status, ctx->pair_status.p, n, cudaMemcpyDeviceToHost, s));
  }
  MB_CUDA(cudaStreamSynchronize(s));
  if (pair_tet && ctx->mesh.tet_id_base)
    for (size_t i = 0; i < n; i++) pair_tet[i] += ctx->mesh.tet_id_base;
  MB_CATCH
}

int mb_rpd_compact_bytes(const mb_rpd_result* res, long* n_bytes) {
  if (!res || !n_bytes) return MB_ERR_ARG;
  *n_bytes = res->compact_bytes;
  return MB_OK;
}

int mb_rpd_fetch_compact(mb_rpd_result* res, void* blob, long* cell_offsets) {
  mb_ctx* ctx = res ? res->ctx : nullptr;
  MB_TRY(ctx)
  MB_REQUIRE(res && ctx, MB_ERR_ARG, "null result");
  MB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  if (res->host_only) {  // streamed run: the records are already in pinned host memory
    MB_REQUIRE(res->host_blob, MB_ERR_STATE, "the result lives in a device sink");
    MB_REQUIRE(!res->sink_owned || res->generation == ctx->stream_generation, MB_ERR_STATE, "streamed result superseded by a later run");
    if (blob && res->compact_bytes > 0) memcpy(blob, res->host_blob, (size_t)res->compact_bytes);
    if (cell_offsets) memcpy(cell_offsets, res->host_off, sizeof(long long) * ((size_t)res->n_cells + 1));
    return MB_OK;
  }
  if (blob && res->compact_bytes > 0)
    MB_CUDA(cudaMemcpyAsync(blob, res->blob.p, (size_t)res->compact_bytes, cudaMemcpyDeviceToHost, s));
  if (cell_offsets)
    MB_CUDA(cudaMemcpyAsync(cell_offsets, res->cell_off.p, sizeof(long long) * ((size_t)res->n_cells + 1),
                            cudaMemcpyDeviceToHost, s));
  MB_CUDA(cudaStreamSynchronize(s));
  MB_CATCH
}

// ---- host side of the lean transport format: plane equations from the ids ---------------------------------
// Same operations in the same order as the device code (rpd_device.cuh: tri2plane_exact is evaluated once per tet
// on the device and fetched; bisector_exact is restated here).  This TU's host code is compiled without FMA
// contraction (-ffp-contract=off), so every operation rounds like the device's __f*_rn intrinsics.
static inline float4 bisector_host(const float4& A, const float4& B) {
  const float dx = A.x - B.x, dy = A.y - B.y, dz = A.z - B.z;
  const float sx = A.x + B.x, sy = A.y + B.y, sz = A.z + B.z;
  const float dot = ((sx * dx + sy * dy) + sz * dz) + (B.w - A.w);
  return make_float4(dx, dy, dz, -dot / 2.f);
}

struct LeanTables {
  const float4* tet_planes = nullptr;  // 4 per tet of the resident mesh
  const int4* tet_fid = nullptr;       // f_ids / f_adjs per tet (slim records)
  const int4* tet_fadj = nullptr;
  const float4* site4 = nullptr;
  int n_tet = 0, n_site = 0, tet_id_base = 0;
};

// the face planes live on the device (k_tet_geometry); fetched once per mesh upload, on first use
static LeanTables lean_tables(mb_ctx* ctx) {
  TetMeshDev& M = ctx->mesh;
  if (!ctx->h_tet_planes_valid) {
    ctx->h_tet_planes.resize((size_t)M.n_tet * 4);
    if (M.n_tet > 0) {
      ctx->h_tet_fid.resize((size_t)M.n_tet);
      ctx->h_tet_fadj.resize((size_t)M.n_tet);
      MB_CUDA(cudaMemcpy2DAsync(ctx->h_tet_planes.data(), 4 * sizeof(float4), M.tet_geo.p, 8 * sizeof(float4),
                                4 * sizeof(float4), (size_t)M.n_tet, cudaMemcpyDeviceToHost, ctx->stream));
      MB_CUDA(cudaMemcpyAsync(ctx->h_tet_fid.data(), M.tet_fid.p, sizeof(int4) * (size_t)M.n_tet, cudaMemcpyDeviceToHost, ctx->stream));
      MB_CUDA(cudaMemcpyAsync(ctx->h_tet_fadj.data(), M.tet_fadj.p, sizeof(int4) * (size_t)M.n_tet, cudaMemcpyDeviceToHost, ctx->stream));
      MB_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    ctx->h_tet_planes_valid = true;
  }
  LeanTables T;
  T.tet_planes = ctx->h_tet_planes.data();
  T.tet_fid = ctx->h_tet_fid.data();
  T.tet_fadj = ctx->h_tet_fadj.data();
  T.site4 = ctx->h_site4.data();
  T.n_tet = M.n_tet;
  T.n_site = (int)ctx->h_site4.size();
  T.tet_id_base = M.tet_id_base;
  return T;
}

// host expansion of one compact record into the ConvexCellTransfer layout
// (reference src/rpd3d/convex_cell.h:189-217; copy() convex_cell.cu:933-949)
static bool expand_record(const uint32_t* w, unsigned char* dst, int id, const LeanTables* T) {
  memset(dst, 0, MB_RECORD_BYTES);
  const int tet = (int)w[0], site = (int)w[1];
  const int nb_v = w[2] & 0xff, nb_p = (w[2] >> 8) & 0xff, nb_e = (w[2] >> 16) & 0xff;
  const bool lean = (w[2] & MB_LEAN_FLAG) != 0;
  const bool slim = lean && (w[2] & MB_SLIM_FLAG) != 0;
  const int status = (int)((w[2] >> 24) & 0xf);  // bit 29 = slim format, bit 30 = flagged class, bit 31 = lean format
  int32_t* di = reinterpret_cast<int32_t*>(dst);
  di[0] = status;
  di[1] = id;        // thread_id: debug-only in the reference; the cell index here
  di[2] = site;      // voro_id
  di[3] = tet;       // tet_id
  memcpy(dst + 16, &w[3], 4);  // weight
  dst[20] = 1;       // is_active
  dst[21] = (unsigned char)nb_v;
  dst[22] = (unsigned char)nb_p;
  dst[23] = (unsigned char)nb_e;
  const uint32_t* p = w + 4;
  memcpy(dst + 24, p, 4 * (size_t)nb_v);
  p += nb_v;
  const uint32_t* pl = p;
  if (!lean) p += 4 * nb_p;
  const uint32_t* meta = p;
  p += slim ? (nb_p - 4) : 3 * nb_p;
  if (lean) {
    if (!T) return false;
    const int tl = tet - T->tet_id_base;
    if (tl < 0 || tl >= T->n_tet || site < 0 || site >= T->n_site) return false;
  }
  for (int i = 0; i < nb_p; i++) {
    uint32_t m3[3];  // (id2.x, id2.y, h) of plane i
    if (slim) {
      // tet faces: (f_id, -1, (uchar) f_adj); bisectors: (min, max) of (seed, neighbour), h = 1 -- what K3 stored
      if (i < 4) {
        const int4 fi = T->tet_fid[tet - T->tet_id_base], fa = T->tet_fadj[tet - T->tet_id_base];
        const int id = i == 0 ? fi.x : (i == 1 ? fi.y : (i == 2 ? fi.z : fi.w));
        const int ad = i == 0 ? fa.x : (i == 1 ? fa.y : (i == 2 ? fa.z : fa.w));
        const float h = (float)(unsigned)(unsigned char)ad;
        m3[0] = (uint32_t)id;
        m3[1] = (uint32_t)-1;
        memcpy(&m3[2], &h, 4);
      } else {
        const int nb = (int)meta[i - 4];
        const float h = 1.f;
        m3[0] = (uint32_t)(site < nb ? site : nb);
        m3[1] = (uint32_t)(site < nb ? nb : site);
        memcpy(&m3[2], &h, 4);
      }
    } else {
      memcpy(m3, meta + 3 * i, 12);
    }
    if (lean) {
      float4 eq;
      if (i < 4) {
        eq = T->tet_planes[(size_t)(tet - T->tet_id_base) * 4 + i];
      } else {
        const int a = (int)m3[0], b = (int)m3[1];
        const int nb = (a == site) ? b : a;  // id2 = (min, max) of (seed, neighbour)
        if (nb < 0 || nb >= T->n_site) return false;
        eq = bisector_host(T->site4[site], T->site4[nb]);
      }
      memcpy(dst + 416 + 32 * i, &eq, 16);
    } else {
      memcpy(dst + 416 + 32 * i, pl + 4 * i, 16);
    }
    memcpy(dst + 416 + 32 * i + 16, &m3[2], 4);  // h
    memcpy(dst + 2464 + 8 * i, m3, 8);            // id2
  }
  memcpy(dst + 2976, p, 3 * (size_t)nb_e);
  const float m1 = -1.f;
  memcpy(dst + 3432, &m1, 4);  // euler
  memcpy(dst + 3436, &m1, 4);  // cell_vol
  di[3440 / 4] = id;
  return true;
}

static void expand_all(mb_ctx* ctx, const uint32_t* blob, const long long* offs, long n, long first_id, unsigned char* out) {
  // the plane tables are only needed (and only fetched from the device) for lean records
  const bool any_lean = n > 0 && (blob[offs[0] / 4 + 2] & MB_LEAN_FLAG) != 0;
  LeanTables T;
  if (any_lean) T = lean_tables(ctx);
  long bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (long i = 0; i < n; i++)
    bad += expand_record(blob + offs[i] / 4, out + (size_t)i * MB_RECORD_BYTES, (int)(first_id + i), any_lean ? &T : nullptr) ? 0 : 1;
  MB_REQUIRE(bad == 0, MB_ERR_STATE,
             "lean records refer to tets / sites outside the context's resident mesh and sites: expand them with the "
             "context (mesh, tet id base, sites) they were computed from");
}

int mb_rpd_expand_compact(mb_ctx* ctx, const void* blob, const long* cell_offsets, long n_cells, long first_id, void* dst) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx && blob && cell_offsets && dst && n_cells >= 0, MB_ERR_ARG, "bad arguments");
  MB_CUDA(cudaSetDevice(ctx->device));
  expand_all(ctx, static_cast<const uint32_t*>(blob), reinterpret_cast<const long long*>(cell_offsets), n_cells, first_id,
             static_cast<unsigned char*>(dst));
  MB_CATCH
}

int mb_rpd_fetch_records(mb_rpd_result* res, void* dst) {
  mb_ctx* ctx = res ? res->ctx : nullptr;
  MB_TRY(ctx)
  MB_REQUIRE(res && ctx && dst, MB_ERR_ARG, "null argument");
  MB_CUDA(cudaSetDevice(ctx->device));
  const long n = res->n_cells;
  if (n == 0) return MB_OK;
  const size_t off_bytes = sizeof(long long) * ((size_t)n + 1);
  const long long* offs;
  const uint32_t* blob;
  if (res->host_only) {
    MB_REQUIRE(res->host_blob, MB_ERR_STATE, "the result lives in a device sink");
    MB_REQUIRE(!res->sink_owned || res->generation == ctx->stream_generation, MB_ERR_STATE, "streamed result superseded by a later run");
    offs = res->host_off;
    blob = res->host_blob;
  } else {
    unsigned char* pin = (unsigned char*)ctx->pin_out.reserve((size_t)res->compact_bytes + off_bytes + 16);
    long long* o = reinterpret_cast<long long*>(pin);
    uint32_t* b = reinterpret_cast<uint32_t*>(pin + ((off_bytes + 15) / 16) * 16);
    cudaStream_t s = ctx->stream;
    MB_CUDA(cudaMemcpyAsync(b, res->blob.p, (size_t)res->compact_bytes, cudaMemcpyDeviceToHost, s));
    MB_CUDA(cudaMemcpyAsync(o, res->cell_off.p, off_bytes, cudaMemcpyDeviceToHost, s));
    MB_CUDA(cudaStreamSynchronize(s));
    offs = o;
    blob = b;
  }
  expand_all(ctx, blob, offs, n, 0, reinterpret_cast<unsigned char*>(dst));
  MB_CATCH
}

int mb_rpd_site_volumes(mb_rpd_result* res, float* vol, float* bary_sum_soa) {
  mb_ctx* ctx = res ? res->ctx : nullptr;
  MB_TRY(ctx)
  MB_REQUIRE(res && ctx, MB_ERR_ARG, "null result");
  MB_REQUIRE(res->want_volumes && res->site_vol.p, MB_ERR_STATE, "run with opts.want_volumes = 1");
  cudaStream_t s = ctx->stream;
  if (vol) MB_CUDA(cudaMemcpyAsync(vol, res->site_vol.p, sizeof(float) * (size_t)res->n_site, cudaMemcpyDeviceToHost, s));
  if (bary_sum_soa)
    MB_CUDA(cudaMemcpyAsync(bary_sum_soa, res->site_bary.p, sizeof(float) * 3 * (size_t)res->n_site, cudaMemcpyDeviceToHost, s));
  MB_CUDA(cudaStreamSynchronize(s));
  MB_CATCH
}

int mb_rpd_cell_volumes(mb_rpd_result* res, float* cell_vol) {
  mb_ctx* ctx = res ? res->ctx : nullptr;
  MB_TRY(ctx)
  MB_REQUIRE(res && ctx && cell_vol, MB_ERR_ARG, "null argument");
  MB_REQUIRE(res->want_volumes && res->cell_vol.p, MB_ERR_STATE, "run with opts.want_volumes = 1");
  if (res->n_cells > 0)
    MB_CUDA(cudaMemcpyAsync(cell_vol, res->cell_vol.p, sizeof(float) * (size_t)res->n_cells, cudaMemcpyDeviceToHost, ctx->stream));
  MB_CUDA(cudaStreamSynchronize(ctx->stream));
  MB_CATCH
}

int mb_rpd_device_buffers(mb_rpd_result* res, void** d_blob, long* n_bytes, void** d_offsets,
                          long* n_cells) {
  if (!res) return MB_ERR_ARG;
  if (res->host_only) return MB_ERR_STATE;  // a streamed run keeps nothing on the device
  if (d_blob) *d_blob = res->blob.p;
  if (n_bytes) *n_bytes = res->compact_bytes;
  if (d_offsets) *d_offsets = res->cell_off.p;
  if (n_cells) *n_cells = res->n_cells;
  return MB_OK;
}

int mb_rpd_emit(mb_rpd_result* res, int max_surf_fid, mb_emit_counts* counts) {
  mb_ctx* ctx = res ? res->ctx : nullptr;
  MB_TRY(ctx)
  MB_REQUIRE(res && ctx, MB_ERR_ARG, "null result");
  MB_REQUIRE(!res->host_only, MB_ERR_STATE, "K4 emission needs a device-resident result (mb_rpd_run)");
  MB_CUDA(cudaSetDevice(ctx->device));
  rpd_emit(ctx, res, max_surf_fid);
  if (counts) *counts = res->emit_counts;
  MB_CATCH
}

#define FETCH(dst, buf, count, T)                                                                   \
  if ((dst) && (count) > 0)                                                                         \
  MB_CUDA(cudaMemcpyAsync((dst), (buf).p, sizeof(T) * (size_t)(count), cudaMemcpyDeviceToHost, s))

int mb_rpd_fetch_emit(mb_rpd_result* res, int* facet_cell, int* facet_key, unsigned char* facet_is_tet,
                      int* vert_cell, int* vert_lvid, int* vert_key3, float* vert_pos3,
                      int* vert_surf_fid, int* edge_cell, int* edge_key2, int* edge_lvid2,
                      float* cell_euler) {
  mb_ctx* ctx = res ? res->ctx : nullptr;
  MB_TRY(ctx)
  MB_REQUIRE(res && ctx, MB_ERR_ARG, "null result");
  MB_REQUIRE(res->emitted, MB_ERR_STATE, "mb_rpd_emit must be called first");
  MB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const mb_emit_counts& c = res->emit_counts;
  FETCH(facet_cell, res->f_cell, c.n_facets, int);
  FETCH(facet_key, res->f_key, c.n_facets, int);
  FETCH(facet_is_tet, res->f_istet, c.n_facets, unsigned char);
  FETCH(vert_cell, res->v_cell, c.n_vertices, int);
  FETCH(vert_lvid, res->v_lvid, c.n_vertices, int);
  FETCH(vert_key3, res->v_key3, 3 * c.n_vertices, int);
  FETCH(vert_pos3, res->v_pos3, 3 * c.n_vertices, float);
  FETCH(vert_surf_fid, res->v_surf, c.n_vertices, int);
  FETCH(edge_cell, res->e_cell, c.n_edges, int);
  FETCH(edge_key2, res->e_key2, 2 * c.n_edges, int);
  FETCH(edge_lvid2, res->e_lvid2, 2 * c.n_edges, int);
  FETCH(cell_euler, res->c_euler, res->n_cells, float);
  MB_CUDA(cudaStreamSynchronize(s));
  MB_CATCH
}

int mb_rpd_fetch_facet_centroids(mb_rpd_result* res, float* centroid3) {
  mb_ctx* ctx = res ? res->ctx : nullptr;
  MB_TRY(ctx)
  MB_REQUIRE(res && ctx && centroid3, MB_ERR_ARG, "null argument");
  MB_REQUIRE(res->emitted, MB_ERR_STATE, "mb_rpd_emit must be called first");
  MB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  FETCH(centroid3, res->f_centroid3, 3 * res->emit_counts.n_facets, float);
  MB_CUDA(cudaStreamSynchronize(s));
  MB_CATCH
}

int mb_set_feature_edges(mb_ctx* ctx, const int* rows6, long n) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx && n >= 0 && (n == 0 || rows6), MB_ERR_ARG, "bad arguments");
  TetMeshDev& M = ctx->mesh;
  MB_REQUIRE(M.n_tet > 0, MB_ERR_STATE, "mb_set_tetmesh first (it clears the feature-edge map)");
  MB_CUDA(cudaSetDevice(ctx->device));
  M.n_fe = 0;
  if (n == 0) return MB_OK;
  std::vector<int> table((size_t)M.n_tet * 6, -1);
  for (long i = 0; i < n; i++) {
    const int t = rows6[6 * i], a = rows6[6 * i + 1], b = rows6[6 * i + 2];
    MB_REQUIRE(t >= 0 && t < M.n_tet && a >= 0 && a < b && b < 4, MB_ERR_ARG,
               "feature-edge rows are (tet, lf_min < lf_max in 0..3, fe_type, fe_id, fe_line_id)");
    table[(size_t)t * 6 + (a == 0 ? b - 1 : (a == 1 ? b + 1 : 5))] = (int)i;
  }
  M.fe_table.reserve(table.size());
  M.fe_rows.reserve(6 * (size_t)n);
  MB_CUDA(cudaMemcpyAsync(M.fe_table.p, table.data(), sizeof(int) * table.size(), cudaMemcpyHostToDevice, ctx->stream));
  MB_CUDA(cudaMemcpyAsync(M.fe_rows.p, rows6, sizeof(int) * 6 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  MB_CUDA(cudaStreamSynchronize(ctx->stream));
  M.n_fe = n;
  MB_CATCH
}

int mb_rpd_feature_edge_count(const mb_rpd_result* res, long* n_hits) {
  if (!res || !n_hits) return MB_ERR_ARG;
  if (!res->emitted) return MB_ERR_STATE;
  *n_hits = res->n_fe_hits;
  return MB_OK;
}

int mb_rpd_fetch_feature_edges(mb_rpd_result* res, int* hit6, int* end4, float* end_pos3) {
  mb_ctx* ctx = res ? res->ctx : nullptr;
  MB_TRY(ctx)
  MB_REQUIRE(res && ctx, MB_ERR_ARG, "null result");
  MB_REQUIRE(res->emitted, MB_ERR_STATE, "mb_rpd_emit must be called first");
  MB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  FETCH(hit6, res->fe_hit6, 6 * res->n_fe_hits, int);
  FETCH(end4, res->fe_end4, 8 * res->n_fe_hits, int);
  FETCH(end_pos3, res->fe_end_pos3, 6 * res->n_fe_hits, float);
  MB_CUDA(cudaStreamSynchronize(s));
  MB_CATCH
}

int mb_rpd_topology(mb_rpd_result* res, mb_topo_counts* counts) {
  mb_ctx* ctx = res ? res->ctx : nullptr;
  MB_TRY(ctx)
  MB_REQUIRE(res && ctx, MB_ERR_ARG, "null result");
  MB_REQUIRE(res->emitted, MB_ERR_STATE, "mb_rpd_emit must be called first");
  MB_CUDA(cudaSetDevice(ctx->device));
  rpd_topology(ctx, res);
  if (counts) {
    counts->n_cells = res->n_cells;
    counts->n_facets = res->emit_counts.n_facets;
    counts->n_edges = res->emit_counts.n_edges;
    counts->n_sites = res->n_site;
    counts->n_halfplane_pairs = res->topo_pairs;
  }
  MB_CATCH
}

int mb_rpd_fetch_topology(mb_rpd_result* res, int* cell_cc, int* facet_cc, int* edge_cc, int* site_n_cells, int* site_n_cc,
                          double* site_euler_sum, int* pair_site, int* pair_neigh, int* pair_n_cc) {
  mb_ctx* ctx = res ? res->ctx : nullptr;
  MB_TRY(ctx)
  MB_REQUIRE(res && ctx, MB_ERR_ARG, "null result");
  MB_REQUIRE(res->topo_done, MB_ERR_STATE, "mb_rpd_topology must be called first");
  MB_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const bool have = res->n_cells > 0 && res->emit_counts.n_facets > 0;
  if (have) {
    FETCH(cell_cc, res->t_cell_cc, res->n_cells, int);
    FETCH(facet_cc, res->t_facet_cc, res->emit_counts.n_facets, int);
    FETCH(edge_cc, res->t_edge_cc, res->emit_counts.n_edges, int);
    FETCH(pair_site, res->t_pair_site, res->topo_pairs, int);
    FETCH(pair_neigh, res->t_pair_neigh, res->topo_pairs, int);
    FETCH(pair_n_cc, res->t_pair_ncc, res->topo_pairs, int);
  }
  FETCH(site_n_cells, res->t_site_n_cells, res->n_site, int);
  FETCH(site_n_cc, res->t_site_n_cc, res->n_site, int);
  FETCH(site_euler_sum, res->t_site_euler, res->n_site, double);
  MB_CUDA(cudaStreamSynchronize(s));
  MB_CATCH
}

// ---- IO_CUDA result format (bgeo.cu) -----------------------------------------------------------------------
int mb_bgeo_write_records(const void* records, long n_cells, int max_sf_fid, int is_boundary_only, const char* path,
                          long* n_points, long* n_polygons) {
  try {
    if (!records || n_cells < 0 || !path) return MB_ERR_ARG;
    bgeo_write_records(static_cast<const unsigned char*>(records), n_cells, max_sf_fid, is_boundary_only != 0, path,
                       n_points, n_polygons);
  } catch (const MbError& e) {
    fprintf(stderr, "[libmat_b200] mb_bgeo_write_records: %s\n", e.msg.c_str());
    return e.code;
  } catch (...) {
    return MB_ERR_NOMEM;
  }
  return MB_OK;
}

int mb_rpd_write_bgeo(mb_ctx* ctx, const void* blob, const long* cell_offsets, long n_cells, int max_sf_fid,
                      int is_boundary_only, const char* path, long* n_points, long* n_polygons) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx && blob && cell_offsets && path && n_cells >= 0, MB_ERR_ARG, "bad arguments");
  MB_CUDA(cudaSetDevice(ctx->device));
  // expand slice by slice (full or lean records -> ConvexCellTransfer layout) and feed the writer's cell loop
  const uint32_t* b = static_cast<const uint32_t*>(blob);
  const long long* offs = reinterpret_cast<const long long*>(cell_offsets);
  bgeo_write_sliced(
      n_cells, [&](long first, long count, unsigned char* dst) { expand_all(ctx, b, offs + first, count, first, dst); },
      max_sf_fid, is_boundary_only != 0, path, n_points, n_polygons);
  MB_CATCH
}

int mb_dist2mat_upload(mb_ctx* ctx, const float* spheres, int n_sph, const float* samples,
                       int n_samples, const unsigned* offset, const unsigned* count,
                       const int* prims, long n_prims) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx, MB_ERR_ARG, "null context");
  MB_REQUIRE(spheres && n_sph > 0 && n_samples >= 0 && n_prims >= 0, MB_ERR_ARG, "bad dist2mat arguments");
  MB_REQUIRE(n_samples == 0 || (samples && offset && count), MB_ERR_ARG, "null sample arrays");
  MB_CUDA(cudaSetDevice(ctx->device));
  d2m_upload(ctx, spheres, n_sph, samples, n_samples, offset, count, prims, n_prims);
  MB_CATCH
}

int mb_dist2mat_run(mb_ctx* ctx, float* kernel_ms) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx, MB_ERR_ARG, "null context");
  MB_CUDA(cudaSetDevice(ctx->device));
  d2m_run(ctx, kernel_ms);
  MB_CATCH
}

int mb_dist2mat_fetch(mb_ctx* ctx, float* result, int* closest_id, unsigned char* tie_flag) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx, MB_ERR_ARG, "null context");
  MB_CUDA(cudaSetDevice(ctx->device));
  d2m_fetch(ctx, result, closest_id, tie_flag);
  MB_CATCH
}

int mb_dist2mat(mb_ctx* ctx, const float* spheres, int n_sph, const float* samples, int n_samples,
                const unsigned* offset, const unsigned* count, const int* prims, long n_prims,
                float* result, int* closest_id, unsigned char* tie_flag) {
  int rc = mb_dist2mat_upload(ctx, spheres, n_sph, samples, n_samples, offset, count, prims, n_prims);
  if (rc) return rc;
  rc = mb_dist2mat_run(ctx, nullptr);
  if (rc) return rc;
  return mb_dist2mat_fetch(ctx, result, closest_id, tie_flag);
}

int mb_dist2mat_set_medial_mesh(mb_ctx* ctx, const float* spheres, int n_sph, const int* mm_faces, int n_faces,
                                const int* mm_edges, int n_edges) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx && spheres && n_sph > 0 && n_faces >= 0 && n_edges >= 0, MB_ERR_ARG, "bad medial mesh");
  MB_REQUIRE((n_faces == 0 || mm_faces) && (n_edges == 0 || mm_edges), MB_ERR_ARG, "null medial faces / edges");
  for (long i = 0; i < 3L * n_faces; i++) MB_REQUIRE(mm_faces[i] >= 0 && mm_faces[i] < n_sph, MB_ERR_ARG, "medial face refers to a sphere out of range");
  for (long i = 0; i < 2L * n_edges; i++) MB_REQUIRE(mm_edges[i] >= 0 && mm_edges[i] < n_sph, MB_ERR_ARG, "medial edge refers to a sphere out of range");
  MB_CUDA(cudaSetDevice(ctx->device));
  d2m_set_medial_mesh(ctx, spheres, n_sph, mm_faces, n_faces, mm_edges, n_edges);
  MB_CATCH
}

int mb_dist2mat_set_face_sites(mb_ctx* ctx, const int* fid_site_rows, long n_rows, int n_fid) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx && n_rows >= 0 && n_fid >= 0 && (n_rows == 0 || fid_site_rows), MB_ERR_ARG, "bad arguments");
  for (long i = 0; i < n_rows; i++)
    MB_REQUIRE(fid_site_rows[2 * i] >= 0 && fid_site_rows[2 * i] < n_fid && fid_site_rows[2 * i + 1] >= 0 &&
                   fid_site_rows[2 * i + 1] < ctx->d2m.n_sph, MB_ERR_ARG, "(fid, site) row out of range");
  MB_CUDA(cudaSetDevice(ctx->device));
  d2m_set_face_sites(ctx, fid_site_rows, n_rows, n_fid);
  MB_CATCH
}

int mb_dist2mat_set_face_sites_from_rpd(mb_ctx* ctx, mb_rpd_result* res, int max_surf_fid) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx && res && res->ctx == ctx && max_surf_fid >= 0, MB_ERR_ARG, "bad arguments");
  MB_REQUIRE(!res->host_only, MB_ERR_STATE, "needs a device-resident result (mb_rpd_run)");
  MB_REQUIRE(res->n_site <= ctx->d2m.n_sph, MB_ERR_ARG, "RPD site ids must index the medial spheres");
  MB_CUDA(cudaSetDevice(ctx->device));
  d2m_set_face_sites_from_rpd(ctx, res, max_surf_fid);
  MB_CATCH
}

int mb_dist2mat_upload_by_face(mb_ctx* ctx, const float* samples, const int* sample_fid, int n_samples) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx && n_samples >= 0 && (n_samples == 0 || (samples && sample_fid)), MB_ERR_ARG, "bad arguments");
  MB_CUDA(cudaSetDevice(ctx->device));
  d2m_upload_by_face(ctx, samples, sample_fid, n_samples);
  MB_CATCH
}

int mb_dist2mat_fetch_closest_prims(mb_ctx* ctx, int* closest_prim3) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx && closest_prim3, MB_ERR_ARG, "null argument");
  MB_CUDA(cudaSetDevice(ctx->device));
  d2m_fetch_closest_prims(ctx, closest_prim3);
  MB_CATCH
}

int mb_dist2mat_by_face(mb_ctx* ctx, const float* samples, const int* sample_fid, int n_samples, float* result,
                        int* closest_id, int* closest_prim3, unsigned char* tie_flag) {
  int rc = mb_dist2mat_upload_by_face(ctx, samples, sample_fid, n_samples);
  if (rc) return rc;
  rc = mb_dist2mat_run(ctx, nullptr);
  if (rc) return rc;
  rc = mb_dist2mat_fetch(ctx, result, closest_id, tie_flag);
  if (rc || !closest_prim3) return rc;
  return mb_dist2mat_fetch_closest_prims(ctx, closest_prim3);
}

int mb_dist2mat_face_list_size(mb_ctx* ctx, long* n_fid, long* n_prims) {
  if (!ctx) return MB_ERR_ARG;
  if (!ctx->d2m_lists.have_lists) return MB_ERR_STATE;
  if (n_fid) *n_fid = ctx->d2m_lists.n_fid;
  if (n_prims) *n_prims = ctx->d2m.n_prims;
  return MB_OK;
}

int mb_dist2mat_fetch_face_lists(mb_ctx* ctx, long long* list_off, int* prims3) {
  MB_TRY(ctx)
  MB_REQUIRE(ctx, MB_ERR_ARG, "null context");
  MB_CUDA(cudaSetDevice(ctx->device));
  d2m_fetch_face_lists(ctx, list_off, prims3);
  MB_CATCH
}

}  // extern "C"
