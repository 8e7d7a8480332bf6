// Internal declarations shared by the libmat_b200 translation units (not installed).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "../../include/libmat_b200.h"

struct MbError {
  int code;
  std::string msg;
};

#define MB_CUDA(call)                                                                  \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      throw MbError{MB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + \
                                     " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"}; \
    }                                                                                  \
  } while (0)

#define MB_REQUIRE(cond, code, text)        \
  do {                                      \
    if (!(cond)) throw MbError{code, text}; \
  } while (0)

// grow-only device buffer; contents are not preserved on growth
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }  // locals of a launcher that throws (e.g. out of memory) free their allocations
  void reserve(size_t n) {
    if (n <= cap) return;
    if (p) MB_CUDA(cudaFree(p));
    p = nullptr;
    cap = 0;
    size_t want = n + n / 8 + 16;
    MB_CUDA(cudaMalloc(&p, want * sizeof(T)));
    cap = want;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  // hand the allocation over to `o` (which must be empty)
  void move_to(DevBuf<T>& o) {
    o.p = p;
    o.cap = cap;
    p = nullptr;
    cap = 0;
  }
  size_t bytes() const { return cap * sizeof(T); }
};

// pinned host staging buffer (grow-only)
struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  void* reserve(size_t bytes) {
    if (bytes <= cap) return p;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    MB_CUDA(cudaMallocHost(&p, bytes + bytes / 8 + 256));
    cap = bytes + bytes / 8 + 256;
    return p;
  }
  // grow while keeping the first `keep` bytes (streamed results whose total size is only known at the end)
  void* reserve_keep(size_t bytes, size_t keep) {
    if (bytes <= cap) return p;
    void* q = nullptr;
    const size_t want = bytes + bytes / 4 + 256;
    MB_CUDA(cudaMallocHost(&q, want));
    if (p && keep) memcpy(q, p, keep);
    if (p) cudaFreeHost(p);
    p = q;
    cap = want;
    return p;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

// ---- device-side mesh layout (HBM-resident between calls) --------------------------------
struct TetMeshDev {
  int n_vert = 0, n_tet = 0;
  DevBuf<float4> vert4;    // x,y,z, w = __int_as_float(v_adj)            16 B / vertex
  DevBuf<int4> tet_idx;    // 4 vertex ids                                16 B / tet
  DevBuf<int4> tet_fadj;   // f_adjs (int, exact copy)                    16 B / tet
  DevBuf<int4> tet_fid;    // f_ids                                       16 B / tet
  DevBuf<uint2> tet_e6;    // 6 edge adjacency counts as bytes (+2 pad)    8 B / tet
  DevBuf<unsigned> tet_vadj;  // (uchar) v_adjs of the tet's 4 vertices, packed                    4 B / tet
  DevBuf<float4> tet_geo;  // per tet: 4 face planes + 4 vertex cofactor vectors (k_tet_geometry)  128 B / tet
  int range_first = 0, range_count = -1;
  int tet_id_base = 0;     // records carry tet id + base (a rank that uploaded only its shard of the tets)
  DevBuf<int> fe_table;    // optional: 6 row indices per tet into fe_rows (-1 = none): TetMesh::tet_es2fe_map
  DevBuf<int> fe_rows;     // (tet, lf_min, lf_max, fe_type, fe_id, fe_line_id) per feature edge (mb_set_feature_edges)
  long n_fe = 0;
  DevBuf<int> tet_sel;     // optional ascending list of tet ids to process (mb_set_tet_subset)
  int n_sel = 0;
  const int* sel_ptr() const { return n_sel > 0 ? tet_sel.p : nullptr; }
};

// K1 uniform grid over the sites (rpd_grid.cuh)
struct GridDev {
  const float4* site4;      // cell-sorted sites (x,y,z,w)
  const int* sorted_id;     // original site id of sorted slot
  const int* cell_off;      // R^3 + 1
  const float* wmax0;       // R^3   max weight per fine cell (-inf if empty)
  const float* wmax1;       // R1^3  max weight per coarse node (4^3 fine cells)
  int R, R1;
  float minx, miny, minz, h, inv_h;
  float wmax_all;           // max weight over all sites
};

struct SitesDev {
  int n_site = 0, site_k = 0;
  bool given = false;      // site_knn supplied
  DevBuf<float4> site4;    // x,y,z,w=r^2
  DevBuf<unsigned> flags;
  DevBuf<int> nbr;         // given mode: row-major [n_site][site_k] (transposed on device)
  DevBuf<int> knn_staging; // raw (site_k+1) x n_site upload
  DevBuf<float> soa_staging;
  float w_max = 0.f;
};

// counters written by the kernels (one cache line)
struct RpdCounters {
  unsigned long long blob_words;   // bump cursor of the compact scratch, in 4-byte words
  unsigned long long n_clips;      // clip_by_plane calls that reached the exact predicate
  unsigned long long n_culled;     // neighbours rejected by the bounding filter
  unsigned long long n_valid;      // cells with status success
  unsigned long long n_cand_overflow;
  unsigned long long hist[10];
  unsigned long long pad[1];       // [15] conflict tests that needed the FP64 determinant
  unsigned long long n_ovf_tets;   // [16] grid mode: tets handed to the big-list candidate pass
  unsigned long long work_cursor;  // [17] K3 dynamic work distribution
  unsigned long long n_gc;         // [18] K3 per-tet mode: dead plane / edge garbage collections
  unsigned long long n_redo;       // [19] grid mode: cells recomputed at the reference's caps by K3's second pass
  unsigned long long work_cursor2; // [20] second pass work distribution
  unsigned long long reserved[3];  // [21] flagged pairs, [22] flagged valid cells (static-filter class), [23] CNT_FB_TETS
};
#define MB_LEAN_FLAG 0x80000000u  // bit 31 of a record's word 2: lean transport format (no plane equations)
#define MB_SLIM_FLAG 0x20000000u  // bit 29 (with bit 31): slim transport format (plane ids reduced to the neighbour id)
#define CNT_OVF_TETS 16
#define CNT_WORK_CURSOR 17
#define CNT_FB_TETS 23  // grid mode: tets the cluster search handed back to the per-tet search

// mapped pinned host memory the kernels publish stage scalars into (rpd_kernels.cu: publish)
struct HostScalars {
  int n_pairs;
  int pad_;
  long long total_words;
  RpdCounters counters;
  long long zero;
  long long cut_off[40];   // staged streamed run: pair offsets at the span cuts (k_publish_cuts)
  unsigned long long seq;  // written last by every publish kernel; the host spins on it
};

struct mb_rpd_result {
  mb_ctx* ctx = nullptr;
  long n_pairs = 0, n_cells = 0, n_clips = 0, n_culled = 0, n_cand_overflow = 0, n_ovf_tets = 0, n_exact = 0;
  long n_redo = 0, n_gc = 0;
  long n_flag_pairs = 0, n_flag_cells = 0;  // flagged class: a conflict |det| under the predicate_generator bound
  unsigned long long generation = 0;        // streamed results: the context's streamed-run counter when it was produced
  long hist[10] = {0};
  long compact_bytes = 0;
  float ms[4] = {0, 0, 0, 0};
  int n_site = 0;
  bool synced = false, want_volumes = false;
  // device results (owned)
  DevBuf<uint32_t> blob;       // ordered compact records
  DevBuf<long long> cell_off;  // n_cells+1 byte offsets into blob
  DevBuf<float> site_vol, site_bary, cell_vol;
  std::vector<cudaEvent_t> evs;  // 4 per processed tet span: start, after K2, after K3, after ordering
  // streamed run (mb_rpd_run_to_host): the ordered records live in the context's pinned host buffers only
  bool host_only = false;
  const uint32_t* host_blob = nullptr;
  const long long* host_off = nullptr;
  int n_spans = 1;
  int lean = 0;            // transport format of a streamed run: 0 full, 1 lean, 2 slim (mb_rpd_opts.lean_records)
  bool sink_owned = true;  // host_blob points into the context's own pinned buffer (else caller memory)
  // emission (K4)
  bool emitted = false;
  mb_emit_counts emit_counts = {0, 0, 0};
  DevBuf<int> f_cell, f_key, v_cell, v_lvid, v_key3, v_surf, e_cell, e_key2, e_lvid2;
  DevBuf<unsigned char> f_istet;
  DevBuf<float> v_pos3, c_euler, f_centroid3;
  // feature-edge hits of the emission (rpd_update.cxx:209-259), when mb_set_feature_edges supplied the map
  long n_fe_hits = 0;
  DevBuf<int> fe_hit6, fe_end4;
  DevBuf<float> fe_end_pos3;
  // topology summary (K6)
  bool topo_done = false;
  long topo_pairs = 0;
  DevBuf<int> t_cell_cc, t_facet_cc, t_edge_cc, t_site_n_cells, t_site_n_cc, t_pair_site, t_pair_neigh, t_pair_ncc;
  DevBuf<double> t_site_euler;
};

struct D2MDev {
  int n_sph = 0, n_samples = 0;
  long n_prims = 0;
  DevBuf<float4> spheres;
  DevBuf<float> samples;  // 3 floats per sample (float3 is not 16-byte aligned; read as scalars)
  DevBuf<unsigned> offset, count;
  DevBuf<int> prims;
  DevBuf<float> result;
  DevBuf<int> closest;
  DevBuf<unsigned char> tie;
};

// f3: device-built dist2mat candidate lists (dist2mat_lists.cu)
struct D2MLists {
  bool have_mesh = false, have_lists = false;
  int n_faces = 0, n_edges = 0, n_fid = 0;
  long n_sf = 0, n_se = 0, n_fs = 0;
  DevBuf<int> mm_faces, mm_edges;              // medial faces (3 sphere ids) / edges (2 sphere ids)
  DevBuf<unsigned long long> sf_keys, se_keys; // sorted (sphere << 32 | face id) / (sphere << 32 | edge id)
  DevBuf<int> sf_first, se_first;              // CSR row starts per sphere
  DevBuf<unsigned long long> fs_keys;          // sorted unique (surface fid << 32 | site)
  DevBuf<int> fs_first;                        // CSR row starts per surface fid
  DevBuf<long long> list_off;                  // per surface fid: start of its primitive list in D2MDev::prims
  DevBuf<int> sample_fid;
  // scratch kept between calls (cudaMalloc / cudaFree synchronise the device and cost far more than the kernels here)
  DevBuf<unsigned long long> tmp_keys, tmp_out;
  DevBuf<int> tmp_flag, tmp_pos, tmp_len, tmp_rows, tmp_prim3;
};

struct mb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;      // the stream in use (own_stream unless mb_set_stream)
  cudaStream_t own_stream = nullptr;
  std::string err;
  int sm_count = 148;
  unsigned long long n_launches = 0;  // kernels launched by this context (bench: gpu_launches)
  TetMeshDev mesh;
  SitesDev sites;
  D2MDev d2m;
  D2MLists d2m_lists;
  int d2m_variant = 0;  // 0 = queue-compacted kernel, 1 = warp-per-sample kernel (MB_D2M_VARIANT, A/B only)
  PinBuf pin_in, pin_out;
  // streamed runs: second stream + double-buffered span results + pinned destination
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_gathered[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  DevBuf<uint32_t> span_blob[2];
  DevBuf<long long> span_off[2];
  PinBuf pin_blob, pin_off;
  HostScalars* hs = nullptr;
  unsigned long long publish_seq = 0;
  std::vector<float4> h_site4;      // host copy of the sites (lean records: bisectors are recomputed on expansion)
  std::vector<float4> h_tet_planes; // host copy of the 4 face planes per tet, fetched on first use
  std::vector<int4> h_tet_fid, h_tet_fadj;  // host copies of f_ids / f_adjs (slim records: tet-face plane ids)
  bool h_tet_planes_valid = false;
  double pairs_per_tet_hint = 0.0;  // grid mode: 1.5 x the largest pairs-per-tet seen (speculative span launches)
  int debug_small_scratch = 0;      // MB_DEBUG_SMALL_SCRATCH=1 (tests): the pipelined ranges get a scratch bound they overflow
  int stream_variant = 0;           // MB_STREAM_VARIANT (A/B tests): 1 = streamed runs redo K2 per span instead of once up front, 3 = staged with host-pipelined ranges
  int k2_variant = 0;               // MB_K2_VARIANT=1 (A/B tests): per-tet candidate search instead of the cluster search
  int clip_variant = 0;             // MB_CLIP_VARIANT=1 (A/B tests): grid-kNN first pass with the state-machine kernel k_clip
  bool no_cull = false;             // MB_NO_CULL=1 (debug / parity tests): no conservative cull of listed neighbours
  unsigned long long stream_generation = 0;  // bumped by every streamed run into the context's own pinned buffers
  int trace_level = 0;
  bool trace_on = false;       // MB_TRACE=1: host-side stage timers, printed by mb_destroy
  double trace_us[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<cudaEvent_t> ev_pool;   // timing events recycled between runs
  // rpd scratch (reused across calls)
  DevBuf<int> tet_cnt, tet_off, pair_tet, pair_site, pair_local, cand_pad;
  DevBuf<int> redo_list;           // grid mode: pairs whose cell outgrew the compact caps of K3's first pass
  DevBuf<int> cand_cnt;            // grid mode: #candidates per tet (cand_pad holds the lists)
  DevBuf<unsigned long long> acc_words;  // pipelined streamed run: record words of the run's earlier ranges
  DevBuf<int> fb_list;             // grid mode: tets the cluster search handed back to the per-tet search
  DevBuf<int> ovf_list;            // grid mode: tets whose survivor list overflowed the fast pass
  int cand_kcap = 0;               // grid mode: row stride of cand_pad
  DevBuf<long long> word_off;      // ordering: exclusive scan of pair_words
  DevBuf<int> pair_valid, pair_cell;
  DevBuf<signed char> pair_status;
  DevBuf<long long> pair_blob;     // per pair: word offset into scratch (or -1)
  DevBuf<int> pair_words;          // per pair: compact size in words (0 if invalid)
  DevBuf<uint32_t> scratch;        // unordered compact records
  DevBuf<RpdCounters> counters;
  DevBuf<unsigned char> cub_tmp;
  // grid (K1)
  DevBuf<int> grid_cnt, grid_off, grid_sorted_id, grid_cell_of;
  DevBuf<float4> grid_site4;
  DevBuf<float> grid_wmax0, grid_wmax1;
  float site_bbox[6] = {0, 0, 0, 0, 0, 0};  // min xyz, max xyz of the site centres (host-computed)
  // incremental recompute (mb_rpd_run_incremental): the candidate lists and sites of the previous run
  DevBuf<int> inc_cand_pad, inc_cand_cnt, inc_affected, inc_flag, inc_pos;
  DevBuf<float4> inc_site4;
  DevBuf<unsigned> inc_flags;
  int inc_n_tet = 0, inc_n_site = 0, inc_kcap = 0, inc_n_affected = 0;
  bool inc_valid = false;
  // result buffers recycled between runs (cudaMalloc / cudaFree synchronise the device): a freed
  // result parks its blob / offset buffers here and the next run takes them back
  DevBuf<uint32_t> spare_blob;
  DevBuf<long long> spare_cell_off;
  std::vector<mb_rpd_result*> live_results;  // orphaned (ctx = nullptr) by mb_destroy
};

// ---- launchers implemented in rpd_kernels.cu ---------------------------------------------
void rpd_upload_mesh(mb_ctx* ctx, const float* verts_aos, int n_vert, const int* idx_aos,
                     int n_tet, const int* v_adjs, const int* e_adjs_dense, const int* e_adj6,
                     const int* f_adjs, const int* f_ids);
void rpd_upload_sites(mb_ctx* ctx, const float* site_soa, const float* site_w,
                      const unsigned* site_flags, int n_site, const int* site_knn, int site_k);
void rpd_run(mb_ctx* ctx, const mb_rpd_opts* opts, mb_rpd_result* res);
void rpd_run_to_host(mb_ctx* ctx, const mb_rpd_opts* opts, int n_chunks, mb_rpd_result* res, void* dst_blob,
                     size_t dst_cap_bytes, long long* dst_off, size_t dst_cap_cells);
void rpd_sync(mb_ctx* ctx, mb_rpd_result* res);
// f2 / config 5: finds the tets whose cells can have changed since the previous incremental run and leaves them as the
// context's tet subset (returns their number; all tets on the first call or after a mesh / capacity change)
int rpd_incremental_select(mb_ctx* ctx, const mb_rpd_opts* opts);
void rpd_fetch_flags(mb_ctx* ctx, mb_rpd_result* res, unsigned char* cell_flag, unsigned char* pair_flag);
void rpd_emit(mb_ctx* ctx, mb_rpd_result* res, int max_surf_fid);
void rpd_topology(mb_ctx* ctx, mb_rpd_result* res);  // K6: cell / facet components + Euler sums per power cell
void rpd_volumes(mb_ctx* ctx, mb_rpd_result* res);  // a12: per-cell / per-site volume + barycentre sums

// ---- bgeo.cu: the IO_CUDA result format (host code) ---------------------------------------------------
void bgeo_write_records(const unsigned char* recs, long n, int max_sf_fid, bool boundary_only, const char* path,
                        long* n_points, long* n_polys);
void bgeo_write_sliced(long n, const std::function<void(long first, long count, unsigned char* dst)>& expand,
                       int max_sf_fid, bool boundary_only, const char* path, long* n_points, long* n_polys);

// ---- dist2mat_kernels.cu -------------------------------------------------------------------
void d2m_upload(mb_ctx* ctx, const float* spheres, int n_sph, const float* samples, int n_samples,
                const unsigned* offset, const unsigned* count, const int* prims, long n_prims);
void d2m_run(mb_ctx* ctx, float* kernel_ms);
void d2m_fetch(mb_ctx* ctx, float* result, int* closest_id, unsigned char* tie_flag);
void peaks_measure(mb_ctx* ctx, double* fp32_tflops, double* fp64_tflops);  // peaks.cu
void tet_adjacency(const int* idx, int n_tet, int n_vert, const int* boundary_sf_fids, int n_sf_facets, int* v_adjs,
                   int* e_adj6, int* f_adjs, int* f_ids, int* n_boundary);  // adjacency.cu (host code)
// ---- dist2mat_lists.cu (f3) ----------------------------------------------------------------------
void d2m_set_medial_mesh(mb_ctx* ctx, const float* spheres, int n_sph, const int* faces, int n_faces, const int* edges, int n_edges);
void d2m_set_face_sites(mb_ctx* ctx, const int* fid_site_rows, long n_rows, int n_fid);
void d2m_set_face_sites_from_rpd(mb_ctx* ctx, mb_rpd_result* res, int max_surf_fid);
void d2m_upload_by_face(mb_ctx* ctx, const float* samples, const int* sample_fid, int n_samples);
void d2m_fetch_closest_prims(mb_ctx* ctx, int* prim3);
void d2m_fetch_face_lists(mb_ctx* ctx, long long* list_off, int* prims3);
