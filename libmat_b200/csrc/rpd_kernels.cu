// RPD3D device pipeline: mesh / site upload, candidate generation (K1/K2), clipping (K3),
// ordering + compaction (first half of K4).  Host orchestration on one CUDA stream.
#include <cub/cub.cuh>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>

#include "mb_internal.h"
#include "rpd_clip.cuh"
#include "rpd_clip2.cuh"
#include "rpd_grid.cuh"

// =============================================================================================
// uploads
// =============================================================================================
// closed form of get_edge_idx (reference src/rpd3d/convex_cell.h:46-59)
static inline long long edge_idx_closed(long long v1, long long v2, long long n) {
  long long vmin = std::min(v1, v2), vmax = std::max(v1, v2);
  return (vmin + 1) * n - vmin * (vmin + 1) / 2 - (n - vmax);
}

// Per-tet geometry that every cell of the tet starts from (ConvexCell ctor, convex_cell.cu:116-214):
// the 4 un-normalised face planes tri2plane(face i) in tet_faces_lvid order (convex_cell.h:30-31) and,
// for the 4 initial vertices (1,3,2) (0,2,3) (0,3,1) (0,1,2), the FP32 cofactor vector of the FP64
// minors of their three planes (the filter data of the clip kernel).  Computed once per mesh upload
// instead of once per (tet, site) cell.  One thread per (tet, face).
__global__ void k_tet_geometry(const float4* __restrict__ vert4, const int4* __restrict__ tet_idx, int n_tet,
                               float4* __restrict__ tet_geo, unsigned* __restrict__ tet_vadj) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = i >> 2, f = i & 3;
  if (t >= n_tet) return;
  const int4 vi = tet_idx[t];
  const float4 q0 = vert4[vi.x], q1 = vert4[vi.y], q2 = vert4[vi.z], q3 = vert4[vi.w];
  const float3 p0 = make_float3(q0.x, q0.y, q0.z), p1 = make_float3(q1.x, q1.y, q1.z),
               p2 = make_float3(q2.x, q2.y, q2.z), p3 = make_float3(q3.x, q3.y, q3.z);
  // faces {2,1,3},{0,2,3},{1,0,3},{0,1,2}
  const float4 pl0 = tri2plane_exact(p2, p1, p3), pl1 = tri2plane_exact(p0, p2, p3),
               pl2 = tri2plane_exact(p1, p0, p3), pl3 = tri2plane_exact(p0, p1, p2);
  const float4 mine = f == 0 ? pl0 : (f == 1 ? pl1 : (f == 2 ? pl2 : pl3));
  // vertex f = dual triangle (1,3,2) (0,2,3) (0,3,1) (0,1,2)
  const Minors m = f == 0 ? minors_exact(pl1, pl3, pl2)
                          : (f == 1 ? minors_exact(pl0, pl2, pl3)
                                    : (f == 2 ? minors_exact(pl0, pl3, pl1) : minors_exact(pl0, pl1, pl2)));
  if (f == 0)  // (uchar) v_adjs of the 4 vertices = w of the 4 initial cell vertices (convex_cell.cu:186-189)
    tet_vadj[t] = ((unsigned)__float_as_int(q0.w) & 0xffu) | (((unsigned)__float_as_int(q1.w) & 0xffu) << 8) |
                  (((unsigned)__float_as_int(q2.w) & 0xffu) << 16) | (((unsigned)__float_as_int(q3.w) & 0xffu) << 24);
  tet_geo[(size_t)t * 8 + f] = mine;
  tet_geo[(size_t)t * 8 + 4 + f] = make_float4((float)m.m234, (float)(-m.m134), (float)m.m124, (float)(-m.m123));
}

void rpd_upload_mesh(mb_ctx* ctx, const float* verts_aos, int n_vert, const int* idx_aos,
                     int n_tet, const int* v_adjs, const int* e_adjs_dense, const int* e_adj6,
                     const int* f_adjs, const int* f_ids) {
  TetMeshDev& M = ctx->mesh;
  M.vert4.reserve(n_vert);
  M.tet_idx.reserve(n_tet);
  M.tet_fadj.reserve(n_tet);
  M.tet_fid.reserve(n_tet);
  M.tet_e6.reserve(n_tet);
  // pack on the host into pinned memory (replaces copy_tet_data voronoi.cu:324-362 and
  // load_num_adjacent_cells_and_ids :379-415: no pitched SoA, no dense e_adjs on the device)
  const size_t bytes_v = sizeof(float4) * (size_t)n_vert;
  const size_t bytes_e = sizeof(uint2) * (size_t)n_tet;
  unsigned char* pin = (unsigned char*)ctx->pin_in.reserve(bytes_v + bytes_e);
  float4* hv = (float4*)pin;
  uint2* he = (uint2*)(pin + bytes_v);
#pragma omp parallel for schedule(static) num_threads(4) if (n_vert > 100000)
  for (int v = 0; v < n_vert; v++) {
    float w;
    int a = v_adjs[v];
    memcpy(&w, &a, 4);
    hv[v] = make_float4(verts_aos[3 * (size_t)v], verts_aos[3 * (size_t)v + 1],
                        verts_aos[3 * (size_t)v + 2], w);
  }
  static const int ep[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
#pragma omp parallel for schedule(static) num_threads(4) if (n_tet > 100000)
  for (int t = 0; t < n_tet; t++) {
    unsigned long long pk = 0;
    for (int e = 0; e < 6; e++) {
      int val;
      if (e_adj6)
        val = e_adj6[6 * (size_t)t + e];
      else
        val = e_adjs_dense[edge_idx_closed(idx_aos[4 * (size_t)t + ep[e][0]],
                                           idx_aos[4 * (size_t)t + ep[e][1]], n_vert)];
      pk |= (unsigned long long)((unsigned char)val) << (8 * e);  // make_uchar3(.., e_adj) truncation
    }
    he[t] = make_uint2((unsigned)(pk & 0xffffffffull), (unsigned)(pk >> 32));
  }
  cudaStream_t s = ctx->stream;
  MB_CUDA(cudaMemcpyAsync(M.vert4.p, hv, bytes_v, cudaMemcpyHostToDevice, s));
  MB_CUDA(cudaMemcpyAsync(M.tet_e6.p, he, bytes_e, cudaMemcpyHostToDevice, s));
  MB_CUDA(cudaMemcpyAsync(M.tet_idx.p, idx_aos, sizeof(int4) * (size_t)n_tet, cudaMemcpyHostToDevice, s));
  MB_CUDA(cudaMemcpyAsync(M.tet_fadj.p, f_adjs, sizeof(int4) * (size_t)n_tet, cudaMemcpyHostToDevice, s));
  MB_CUDA(cudaMemcpyAsync(M.tet_fid.p, f_ids, sizeof(int4) * (size_t)n_tet, cudaMemcpyHostToDevice, s));
  M.tet_geo.reserve((size_t)n_tet * 8);
  M.tet_vadj.reserve((size_t)n_tet);
  ctx->n_launches++;
  k_tet_geometry<<<(unsigned)(((size_t)n_tet * 4 + 255) / 256), 256, 0, s>>>(M.vert4.p, M.tet_idx.p, n_tet, M.tet_geo.p,
                                                                              M.tet_vadj.p);
  MB_CUDA(cudaGetLastError());
  MB_CUDA(cudaStreamSynchronize(s));
  M.n_vert = n_vert;
  M.n_tet = n_tet;
  M.range_first = 0;
  M.range_count = -1;
  M.n_sel = 0;
  M.tet_id_base = 0;
  M.n_fe = 0;  // the feature-edge map belongs to a mesh
  ctx->h_tet_planes_valid = false;
}

// SoA x|y|z + w -> float4; (site_k+1) x n_site slot-major knn -> row-major [n_site][site_k]
__global__ void k_prep_sites(const float* __restrict__ soa, const float* __restrict__ w, int n_site,
                             float4* __restrict__ site4) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_site) site4[s] = make_float4(soa[s], soa[s + n_site], soa[s + 2 * (size_t)n_site], w[s]);
}

__global__ void k_transpose_knn(const int* __restrict__ knn, int n_site, int site_k,
                                int* __restrict__ nbr) {
  __shared__ int tile[32][33];
  // in: [slot][site], out: [site][slot]
  int s0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int k = k0 + r, s = s0 + threadIdx.x;
    tile[r][threadIdx.x] = (k < site_k && s < n_site) ? knn[(size_t)k * n_site + s] : -1;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int s = s0 + r, k = k0 + threadIdx.x;
    if (s < n_site && k < site_k) nbr[(size_t)s * site_k + k] = tile[threadIdx.x][r];
  }
}

void rpd_upload_sites(mb_ctx* ctx, const float* site_soa, const float* site_w,
                      const unsigned* site_flags, int n_site, const int* site_knn, int site_k) {
  SitesDev& S = ctx->sites;
  cudaStream_t s = ctx->stream;
  S.site4.reserve(n_site);
  S.flags.reserve(n_site);
  S.soa_staging.reserve(4 * (size_t)n_site);
  MB_CUDA(cudaMemcpyAsync(S.soa_staging.p, site_soa, sizeof(float) * 3 * (size_t)n_site,
                          cudaMemcpyHostToDevice, s));
  MB_CUDA(cudaMemcpyAsync(S.soa_staging.p + 3 * (size_t)n_site, site_w, sizeof(float) * (size_t)n_site,
                          cudaMemcpyHostToDevice, s));
  MB_CUDA(cudaMemcpyAsync(S.flags.p, site_flags, sizeof(unsigned) * (size_t)n_site,
                          cudaMemcpyHostToDevice, s));
  ctx->n_launches++;
  k_prep_sites<<<(n_site + 255) / 256, 256, 0, s>>>(S.soa_staging.p, S.soa_staging.p + 3 * (size_t)n_site,
                                                    n_site, S.site4.p);
  S.given = site_knn != nullptr;
  S.site_k = site_k;
  if (S.given) {
    MB_REQUIRE(site_k > 0, MB_ERR_ARG, "site_k must be > 0 when site_knn is given");
    S.knn_staging.reserve((size_t)(site_k + 1) * n_site);
    S.nbr.reserve((size_t)site_k * n_site);
    // the last row (slot site_k) is always -1 and never read (triangulation.cxx:245-256,
    // convex_cell.cu:1225,1255): only rows 0..site_k-1 are uploaded
    MB_CUDA(cudaMemcpyAsync(S.knn_staging.p, site_knn, sizeof(int) * (size_t)site_k * n_site,
                            cudaMemcpyHostToDevice, s));
    dim3 grid((n_site + 31) / 32, (site_k + 31) / 32), block(32, 8);
    ctx->n_launches++;
    k_transpose_knn<<<grid, block, 0, s>>>(S.knn_staging.p, n_site, site_k, S.nbr.p);
  }
  ctx->h_site4.resize((size_t)n_site);
  for (int i = 0; i < n_site; i++)
    ctx->h_site4[(size_t)i] = make_float4(site_soa[i], site_soa[(size_t)n_site + i], site_soa[2 * (size_t)n_site + i], site_w[i]);
  float wmax = 0.f;
  float* bb = ctx->site_bbox;
  bb[0] = bb[1] = bb[2] = INFINITY;
  bb[3] = bb[4] = bb[5] = -INFINITY;
  for (int i = 0; i < n_site; i++) {
    wmax = std::max(wmax, site_w[i]);
    for (int c = 0; c < 3; c++) {
      const float v = site_soa[(size_t)c * n_site + i];
      bb[c] = std::min(bb[c], v);
      bb[3 + c] = std::max(bb[3 + c], v);
    }
  }
  S.w_max = wmax;
  S.n_site = n_site;
  MB_CUDA(cudaGetLastError());
  // the caller may free or overwrite its arrays as soon as this returns (header contract): a copy from pinned
  // or registered host memory is truly asynchronous, so wait for it (by now the host loops above have hidden it)
  MB_CUDA(cudaStreamSynchronize(s));
}

// =============================================================================================
// K2 (given-neighbours mode): the tet-sphere relation of the reference, evaluated sparsely.
//   reference: compute_distances + dist_minus_weight (kNN-CUDA/knncuda.cu:21-94,130-160) build a
//   dense n_site x n_vert matrix, tet_sphere_relations_dev (voronoi.cu:154-193) a dense
//   n_site x n_tet matrix, both copied to the host and compacted there (:266-317).
//   here: one warp per tet, lanes over sites, early exit over the neighbour list, ballot-ordered
//   compaction -> per-tet candidate lists in ascending site id (the reference's tet_knn order).
// =============================================================================================
__device__ __forceinline__ float pdist_exact(float4 S, float4 p) {
  // ssd = ((dx*dx + dy*dy) + dz*dz) - w with dx = site - vertex (knncuda.cu:78-81, :153-156)
  const float dx = xfsub(S.x, p.x), dy = xfsub(S.y, p.y), dz = xfsub(S.z, p.z);
  const float ssd = xfadd(xfadd(xfadd(0.f, xfmul(dx, dx)), xfmul(dy, dy)), xfmul(dz, dz));
  return xfsub(ssd, S.w);
}

__device__ __forceinline__ bool relate_exact(const float4* __restrict__ site4,
                                             const unsigned* __restrict__ flags,
                                             const int* __restrict__ nbr, int site_k, int s,
                                             const float4* p) {
  if (flags[s] != 1u) return false;  // SiteFlag::is_selected, voronoi.cu:165-169
  const float4 S = site4[s];
  float pd_i[4];
#pragma unroll
  for (int l = 0; l < 4; l++) pd_i[l] = pdist_exact(S, p[l]);
  const int* row = nbr + (size_t)s * site_k;
  for (int sm = 0; sm < site_k; sm++) {
    const int m = row[sm];
    if (m == -1) continue;  // voronoi.cu:175
    const float4 Mq = site4[m];
    bool any = false;
#pragma unroll
    for (int l = 0; l < 4; l++) any = any || (pdist_exact(Mq, p[l]) > pd_i[l]);
    if (!any) return false;
  }
  return true;
}

#define CAND_PAD 32

template <bool FILL>
__global__ void __launch_bounds__(256) k_cand_given(const float4* __restrict__ vert4,
                                                    const int4* __restrict__ tet_idx, int tet_first,
                                                    int tet_count, const int* __restrict__ tet_sel, const float4* __restrict__ site4,
                                                    const unsigned* __restrict__ flags, int n_site,
                                                    const int* __restrict__ nbr, int site_k,
                                                    int* __restrict__ tet_cnt, int* __restrict__ cand_pad,
                                                    const int* __restrict__ tet_off,
                                                    int* __restrict__ pair_tet, int* __restrict__ pair_site) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= tet_count) return;
  const int t = tet_sel ? tet_sel[warp] : tet_first + warp;
  int count = 0;
  if (FILL) {
    // second pass: copy the padded list, or recompute when it overflowed CAND_PAD
    const int n = tet_cnt[warp];
    const int off = tet_off[warp];
    if (n <= CAND_PAD) {
      if (lane < n) {
        pair_tet[off + lane] = t;
        pair_site[off + lane] = cand_pad[(size_t)warp * CAND_PAD + lane];
      }
      return;
    }
  }
  const int4 vi = tet_idx[t];
  const float4 p[4] = {vert4[vi.x], vert4[vi.y], vert4[vi.z], vert4[vi.w]};
  for (int base = 0; base < n_site; base += 32) {
    const int s = base + lane;
    const bool ok = (s < n_site) && relate_exact(site4, flags, nbr, site_k, s, p);
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const int pos = count + __popc(m & ((1u << lane) - 1u));
      if (FILL) {
        const int off = tet_off[warp];
        pair_tet[off + pos] = t;
        pair_site[off + pos] = s;
      } else if (pos < CAND_PAD) {
        cand_pad[(size_t)warp * CAND_PAD + pos] = s;
      }
    }
    count += __popc(m);
  }
  if (!FILL && lane == 0) tet_cnt[warp] = count;
}

// =============================================================================================
// ordering + compaction: records leave K3 in arbitrary order in the scratch; two exclusive scans
// (record words, valid flags) give each valid cell its slot in (tet, site) order.
// =============================================================================================
// packed scan element: valid-cell count << 36 | record words.  36 bits of words = 256 GiB of records (more than a
// B200 holds), 28 bits of cells = 268 M valid cells per span (84 GB of records at ~314 B / cell); rpd_run_span
// refuses a span beyond either limit instead of wrapping.
#define PACK_SHIFT 36
#define PACK_MASK ((1ull << PACK_SHIFT) - 1ull)
#define PACK_MAX_CELLS ((1ull << (64 - PACK_SHIFT)) - 1ull)
// pair_words = record words | nb_p << 16 (0 = no record).  LEAN: the transport format without the 4 * nb_p
// plane-equation words (recomputable from the ids: tet face planes, power bisectors)
// SLIM (LEAN == 2): additionally the 3 id words per plane shrink to ONE word per bisector (the neighbour site id;
// seed, tet-face ids and adjacency counts are functions of the header): words - 7 nb_p + (nb_p - 4)
template <int LEAN>
struct PackWords {
  __host__ __device__ unsigned long long operator()(int pw) const {
    const int words = pw & 0xffff;
    const int nb_p = (pw >> 16) & 0xff;
    const int out = words == 0 ? 0 : (LEAN == 2 ? words - 6 * nb_p - 4 : (LEAN == 1 ? words - 4 * nb_p : words));
    return (unsigned long long)out | ((unsigned long long)(words > 0) << PACK_SHIFT);
  }
};

// lanes per record in k_gather (measured at config 2: 8 lanes 0.162 ms for the ordering stage, 16 lanes 0.230 ms --
// a sixth of the pairs has no record and short records leave wide groups idle)
#define GATHER_LANES 8
template <int LEAN>
__global__ void k_gather(const uint32_t* __restrict__ scratch, const long long* __restrict__ pair_blob,
                         const int* __restrict__ pair_words, const unsigned long long* __restrict__ packed_off,
                         long long n_pairs, uint32_t* __restrict__ blob, long long* __restrict__ cell_off,
                         long long total_words, long long n_cells, long long base_bytes,
                         const unsigned long long* __restrict__ total_dev = nullptr,
                         const unsigned long long* __restrict__ base_words_dev = nullptr) {
  // total_dev: the packed scan total still on the device (the gather is launched before the host has seen it);
  // base_words_dev: the record words of the run's earlier ranges, accumulated on the device by k_publish_range
  if (total_dev) {
    const unsigned long long v = *total_dev;
    total_words = (long long)(v & PACK_MASK);
    n_cells = (long long)(v >> PACK_SHIFT);
  }
  if (base_words_dev) base_bytes = (long long)(*base_words_dev * 4ull);
  const long long g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / GATHER_LANES;
  const int lane = threadIdx.x % GATHER_LANES;
  if (g >= n_pairs) return;
  const int pw = pair_words[g];
  const int words = pw & 0xffff;
  if (words == 0) return;
  const unsigned long long pk = packed_off[g];
  const long long dst = (long long)(pk & PACK_MASK);
  const uint32_t* src = scratch + pair_blob[g];
  if (LEAN == 2) {
    // [4 header | nb_v vertices] [plane equations: dropped] [ids: one neighbour id per bisector] [edges]
    const int nb_p = (pw >> 16) & 0xff;
    const int head = 4 + (int)(src[2] & 0xffu);
    const uint32_t seed = src[1];
    for (int i = lane; i < head; i += GATHER_LANES) blob[dst + i] = (i == 2) ? (src[2] | MB_LEAN_FLAG | MB_SLIM_FLAG) : src[i];
    const uint32_t* meta = src + head + 4 * nb_p;
    for (int i = 4 + lane; i < nb_p; i += GATHER_LANES) blob[dst + head + i - 4] = (meta[3 * i] == seed) ? meta[3 * i + 1] : meta[3 * i];
    const int e0 = head + 7 * nb_p, shift = 6 * nb_p + 4;
    for (int i = e0 + lane; i < words; i += GATHER_LANES) blob[dst + i - shift] = src[i];
  } else if (LEAN == 1) {
    // [4 header | nb_v vertices] [4*nb_p plane equations: dropped] [3*nb_p ids | edges]
    const int nb_p = (pw >> 16) & 0xff;
    const int head = 4 + (int)(src[2] & 0xffu);
    const int skip = 4 * nb_p;
    for (int i = lane; i < head; i += GATHER_LANES) blob[dst + i] = (i == 2) ? (src[2] | MB_LEAN_FLAG) : src[i];
    for (int i = head + skip + lane; i < words; i += GATHER_LANES) blob[dst + i - skip] = src[i];
  } else {
    for (int i = lane; i < words; i += GATHER_LANES) blob[dst + i] = src[i];
  }
  if (lane == 0) {
    const long long c = (long long)(pk >> PACK_SHIFT);
    cell_off[c] = base_bytes + dst * 4;
    if (c == n_cells - 1) cell_off[n_cells] = base_bytes + total_words * 4;
  }
}

// K3's counters and the packed scan total reach the host through mapped memory in one launch
__global__ void k_publish_k3(const uint32_t* __restrict__ counters, const unsigned long long* __restrict__ total,
                             const int* __restrict__ n_pairs_dev, HostScalars* __restrict__ hs, unsigned long long seq) {
  uint32_t* dst = reinterpret_cast<uint32_t*>(&hs->counters);
  for (int i = threadIdx.x; i < (int)(sizeof(RpdCounters) / 4); i += blockDim.x) dst[i] = counters[i];
  if (threadIdx.x == 0) hs->total_words = (long long)*total;
  if (threadIdx.x == 1 && n_pairs_dev) hs->n_pairs = *n_pairs_dev;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(&hs->seq) = seq;
}

// staged streamed run: the pair offsets at the span cuts and the candidate-stage counters, one launch
struct CutList {
  int n;
  int tet[36];
};
__global__ void k_publish_cuts(const int* __restrict__ tet_off, CutList cuts, const uint32_t* __restrict__ counters,
                               HostScalars* __restrict__ hs, unsigned long long seq) {
  if ((int)threadIdx.x < cuts.n) hs->cut_off[threadIdx.x] = tet_off[cuts.tet[threadIdx.x]];
  uint32_t* dst = reinterpret_cast<uint32_t*>(&hs->counters);
  for (int i = threadIdx.x; i < (int)(sizeof(RpdCounters) / 4); i += blockDim.x) dst[i] = counters[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(&hs->seq) = seq;
}

// pipelined ranges: K3's counters and the scan total to the range's host slot, the total added to the run's device
// accumulator (the next range's gather reads it as its base offset)
__global__ void k_publish_range(const uint32_t* __restrict__ counters, const unsigned long long* __restrict__ total,
                                unsigned long long* __restrict__ acc_words, HostScalars* __restrict__ hs,
                                unsigned long long seq) {
  uint32_t* dst = reinterpret_cast<uint32_t*>(&hs->counters);
  for (int i = threadIdx.x; i < (int)(sizeof(RpdCounters) / 4); i += blockDim.x) dst[i] = counters[i];
  if (threadIdx.x == 0) {
    const unsigned long long v = total ? *total : 0ull;
    hs->total_words = (long long)v;
    *acc_words += (v & PACK_MASK);
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned long long*>(&hs->seq) = seq;
}
__global__ void k_base_offset(const unsigned long long* __restrict__ acc_words, long long* __restrict__ cell_off) {
  cell_off[0] = (long long)(*acc_words * 4ull);
}

// =============================================================================================
// K1 host side: grid build (counting sort by cell + max-weight pyramid), K2 launch
// =============================================================================================
// Zero-fill by a KERNEL, not cudaMemsetAsync: the driver runs small memsets on a copy engine, where they queue behind
// the bulk D2H copy of the previous tet span (measured: every span that overlapped a 27 MB host copy ran 0.2-0.4 ms
// late -- its counter / tail-word memsets were waiting on the copy engine).
__global__ void k_zero_u32(unsigned* __restrict__ p, size_t n_words) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x) p[i] = 0u;
}
static void dev_zero(mb_ctx* ctx, void* p, size_t bytes) {
  if (bytes == 0) return;
  const size_t n = bytes / 4;  // every caller zeroes whole 4-byte words
  const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)ctx->sm_count * 8);
  ctx->n_launches++;
  k_zero_u32<<<std::max(1, blocks), 256, 0, ctx->stream>>>(reinterpret_cast<unsigned*>(p), n);
  MB_CUDA(cudaGetLastError());
}

static GridDev grid_build(mb_ctx* ctx) {
  SitesDev& S = ctx->sites;
  cudaStream_t s = ctx->stream;
  GridDev G;
  int R = (int)std::ceil(std::cbrt(std::max(1.0, S.n_site / 2.0)));
  R = ((R + 3) / 4) * 4;
  R = std::max(4, std::min(128, R));
  G.R = R;
  G.R1 = R / 4;
  const float* bb = ctx->site_bbox;
  const float ext = std::max(std::max(bb[3] - bb[0], bb[4] - bb[1]), std::max(bb[5] - bb[2], 1e-3f));
  G.h = ext * 1.0001f / R;
  G.inv_h = 1.f / G.h;
  G.minx = bb[0];
  G.miny = bb[1];
  G.minz = bb[2];
  G.wmax_all = S.w_max;
  const int nc = R * R * R, n1 = G.R1 * G.R1 * G.R1;
  ctx->grid_cnt.reserve((size_t)nc + 1);
  ctx->grid_off.reserve((size_t)nc + 1);
  ctx->grid_cell_of.reserve(S.n_site);
  ctx->grid_sorted_id.reserve(S.n_site);
  ctx->grid_site4.reserve(S.n_site);
  ctx->grid_wmax0.reserve(nc);
  ctx->grid_wmax1.reserve(n1);
  dev_zero(ctx, ctx->grid_cnt.p, sizeof(int) * ((size_t)nc + 1));
  ctx->n_launches++;
  k_grid_count<<<(S.n_site + 255) / 256, 256, 0, s>>>(S.site4.p, S.n_site, G, ctx->grid_cnt.p, ctx->grid_cell_of.p);
  {
    size_t tmp = 0;
    MB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, ctx->grid_cnt.p, ctx->grid_off.p, nc + 1, s));
    ctx->cub_tmp.reserve(tmp);
    ctx->n_launches += 2;  // DeviceScanInitKernel + DeviceScanKernel
    MB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, ctx->grid_cnt.p, ctx->grid_off.p, nc + 1, s));
  }
  dev_zero(ctx, ctx->grid_cnt.p, sizeof(int) * ((size_t)nc + 1));
  ctx->n_launches++;
  k_grid_scatter<<<(S.n_site + 255) / 256, 256, 0, s>>>(S.site4.p, S.n_site, ctx->grid_cell_of.p,
                                                        ctx->grid_off.p, ctx->grid_cnt.p, ctx->grid_sorted_id.p);
  ctx->n_launches++;
  k_grid_finalize<<<(nc + 127) / 128, 128, 0, s>>>(S.site4.p, ctx->grid_off.p, nc, ctx->grid_sorted_id.p,
                                                   ctx->grid_site4.p, ctx->grid_wmax0.p);
  ctx->n_launches++;
  k_grid_pyramid<<<(n1 + 127) / 128, 128, 0, s>>>(ctx->grid_wmax0.p, R, G.R1, ctx->grid_wmax1.p);
  MB_CUDA(cudaGetLastError());
  G.site4 = ctx->grid_site4.p;
  G.sorted_id = ctx->grid_sorted_id.p;
  G.cell_off = ctx->grid_off.p;
  G.wmax0 = ctx->grid_wmax0.p;
  G.wmax1 = ctx->grid_wmax1.p;
  return G;
}

// a contiguous piece of the processed tets: [first, first+count) of the mesh, or count entries of a
// device-resident id list (mb_set_tet_subset); all per-tet scratch is indexed relative to it
struct TetSpan {
  int first, count;
  const int* sel;
};

// fills cand_pad / cand_cnt (all candidates) and tet_cnt (flagged candidates = pairs)
template <int KCAP>
static void launch_grid_candidates(mb_ctx* ctx, const GridDev& G, const TetSpan& sp, int kcap_out) {
  TetMeshDev& M = ctx->mesh;
  cudaStream_t s = ctx->stream;
  unsigned long long* cnt = reinterpret_cast<unsigned long long*>(ctx->counters.p);
  constexpr int WARPS = 4;
  const size_t smem = (size_t)WARPS * (KCAP * 24 + 256);  // float4 pd + float w + int id per entry, + 64-entry queue
  const size_t smem_big = (size_t)GRID_BIG_KCAP * 24 + 256;
  static bool attr_set_dev[64] = {false};
  bool& attr_set = attr_set_dev[ctx->device & 63];  // function attributes are per device
  if (!attr_set) {
    MB_CUDA(cudaFuncSetAttribute(k_grid_candidates<GRID_BIG_KCAP, 1, 1>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_big));
    attr_set = true;
  }
  if (sp.sel == nullptr && ctx->k2_variant != 1) {
    // fast pass, cluster search: one warp per GRID_CT consecutive tets (one grid walk per cluster); clusters it cannot
    // handle come back through fb_list to the per-tet search (usually none -> the second launch returns at once)
    const size_t smem_c = (size_t)WARPS * ((size_t)GRID_KC * 36 + (size_t)KCAP * 24);
    ctx->n_launches += 2;
    auto launch_c = [&](auto kern, int ct) {
      const int n_clusters = (sp.count + sp.first % ct + ct - 1) / ct;  // aligned to global tet ids (see the kernel)
      const int blocks_c = std::max(1, std::min((n_clusters + WARPS - 1) / WARPS, ctx->sm_count * 6 * 64));
      kern<<<blocks_c, 32 * WARPS, smem_c, s>>>(M.vert4.p, M.tet_idx.p, sp.first, sp.count, G, ctx->sites.flags.p, kcap_out,
                                                ctx->cand_pad.p, ctx->cand_cnt.p, ctx->tet_cnt.p, cnt, ctx->ovf_list.p,
                                                ctx->fb_list.p);
    };
    // (measured at config 2 / config 4: clusters of 6 tets 0.42 / 3.5 ms, of 4 tets 0.47 ms, of 8 tets 0.53 / 6.1 ms --
    // 8 straddles two Kuhn cubes; 8 blocks per SM instead of 6: no change)
    launch_c(k_grid_candidates_cluster<KCAP, WARPS, GRID_CT, 6>, GRID_CT);
    k_grid_candidates<KCAP, WARPS, 2><<<ctx->sm_count * 6, 32 * WARPS, smem, s>>>(
        M.vert4.p, M.tet_idx.p, sp.first, sp.count, sp.sel, G, ctx->sites.flags.p, kcap_out, ctx->cand_pad.p,
        ctx->cand_cnt.p, ctx->tet_cnt.p, cnt, ctx->ovf_list.p, ctx->fb_list.p);
  } else {
    // fast pass: one warp per tet.  The cost per tet varies several-fold, so the hardware block scheduler
    // balances better than a static stride: plain grid up to 64 waves, strided beyond
    const int want = (sp.count + WARPS - 1) / WARPS;
    const int blocks = std::max(1, std::min(want, ctx->sm_count * 6 * 64));
    ctx->n_launches++;
    k_grid_candidates<KCAP, WARPS, 0><<<blocks, 32 * WARPS, smem, s>>>(
        M.vert4.p, M.tet_idx.p, sp.first, sp.count, sp.sel, G, ctx->sites.flags.p, kcap_out, ctx->cand_pad.p,
        ctx->cand_cnt.p, ctx->tet_cnt.p, cnt, ctx->ovf_list.p, ctx->fb_list.p);
  }
  // overflow pass: reads the number of handed-over tets on the device (usually zero -> returns)
  ctx->n_launches++;
  k_grid_candidates<GRID_BIG_KCAP, 1, 1><<<ctx->sm_count * 2, 32, smem_big, s>>>(
      M.vert4.p, M.tet_idx.p, sp.first, sp.count, sp.sel, G, ctx->sites.flags.p, kcap_out, ctx->cand_pad.p,
      ctx->cand_cnt.p, ctx->tet_cnt.p, cnt, ctx->ovf_list.p, ctx->fb_list.p);
  MB_CUDA(cudaGetLastError());
}

static void grid_candidates(mb_ctx* ctx, const GridDev& G, const TetSpan& sp, int grid_k) {
  // grid_k = expected candidates per tet: capacity of the fast pass's shared-memory survivor list
  // (32 / 96 / 256; longer lists go through the big-list pass) and row stride of the output
  const int kcap = (grid_k > 96) ? 256 : 96;
  ctx->cand_kcap = kcap;
  ctx->cand_pad.reserve((size_t)sp.count * kcap);
  ctx->cand_cnt.reserve((size_t)sp.count + 1);
  ctx->ovf_list.reserve((size_t)sp.count + 1);
  ctx->fb_list.reserve((size_t)sp.count + 8);
  if (grid_k > 0 && grid_k <= 32)
    launch_grid_candidates<32>(ctx, G, sp, kcap);
  else if (kcap == 96)
    launch_grid_candidates<96>(ctx, G, sp, kcap);
  else
    launch_grid_candidates<256>(ctx, G, sp, kcap);
}

static void grid_fill_pairs(mb_ctx* ctx, const TetSpan& sp, long long cap_pairs) {
  cudaStream_t s = ctx->stream;
  const int blocks = (sp.count + 7) / 8;
  ctx->n_launches++;
  k_grid_fill<<<blocks, 256, 0, s>>>(sp.first, sp.count, sp.sel, ctx->cand_kcap, ctx->cand_pad.p,
                                     ctx->cand_cnt.p, ctx->tet_off.p, ctx->sites.flags.p, ctx->pair_tet.p,
                                     ctx->pair_site.p, ctx->pair_local.p, cap_pairs);
  MB_CUDA(cudaGetLastError());
}

// =============================================================================================
// host orchestration
// =============================================================================================
template <typename T>
static void exclusive_scan(mb_ctx* ctx, const int* in, T* out, long long n) {
  size_t tmp = 0;
  MB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, n, ctx->stream));
  ctx->cub_tmp.reserve(tmp);
  ctx->n_launches += 2;  // DeviceScanInitKernel + DeviceScanKernel
  MB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, in, out, n, ctx->stream));
}

template <int G, bool PT, bool SMALL>
static void launch_clip_pass(mb_ctx* ctx, ClipArgs A, bool second_pass) {
  constexpr int groups = 128 / G;
  typedef CellT<SMALL ? MBK_SMALL_P : MBK_MAX_P, SMALL ? MBK_SMALL_T : MBK_MAX_T, SMALL ? MBK_SMALL_E : MBK_MAX_E> Cell;
  const size_t smem = sizeof(Cell) * groups;
  static bool attr_set_dev[64] = {false};
  bool& attr_set = attr_set_dev[ctx->device & 63];  // function attributes are per device
  if (!attr_set) {
    MB_CUDA(cudaFuncSetAttribute(k_clip<G, PT, SMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  static int per_sm_dev[64] = {0};
  int& per_sm = per_sm_dev[ctx->device & 63];
  if (per_sm < 1) {
    MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_clip<G, PT, SMALL>, 128, smem));
    if (per_sm < 1) per_sm = 1;
  }
  // persistent grid: every SM fully resident, warps pull chunks of pairs from a global cursor
  long long want = (A.n_pairs + groups - 1) / groups;
  long long grid = std::min<long long>(want, (long long)ctx->sm_count * per_sm);
  if (second_pass) grid = std::min<long long>(grid, (long long)ctx->sm_count * 2);  // usually nothing to do
  if (grid < 1) grid = 1;
  // cursor granularity: 8 rounds of pairs per grab for large runs, finer for small spans so that every
  // warp still gets >= ~16 grabs (a persistent kernel's tail is one grab long)
  {
    constexpr int NG = 32 / G;
    const long long warps = grid * 4;
    long long g = A.n_pairs / (warps * 16);
    g = (g / NG) * NG;
    A.grab = (A.n_pairs_dev || A.work_count) ? 0  // 0: derived on the device from the device-resident count
                                             : (int)std::max<long long>(NG, std::min<long long>(8 * NG, g));
  }
  ctx->n_launches++;
  k_clip<G, PT, SMALL><<<(unsigned)grid, 128, smem, ctx->stream>>>(A);
  MB_CUDA(cudaGetLastError());
}

// grid-kNN mode, first pass: the batch-synchronous compact-caps kernel (rpd_clip2.cuh)
template <int G, int NB, bool PT>
static void launch_clip_tiny(mb_ctx* ctx, ClipArgs A) {
  constexpr int groups = 128 / G;
  const size_t smem = sizeof(CellTiny) * groups;
  static bool attr_set_dev[64] = {false};
  bool& attr_set = attr_set_dev[ctx->device & 63];
  if (!attr_set) {
    MB_CUDA(cudaFuncSetAttribute(k_clip_tiny<G, NB, PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  static int per_sm_dev[64] = {0};
  int& per_sm = per_sm_dev[ctx->device & 63];
  if (per_sm < 1) {
    MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_clip_tiny<G, NB, PT>, 128, smem));
    if (per_sm < 1) per_sm = 1;
  }
  const long long want = (A.n_pairs + groups - 1) / groups;
  long long grid = std::max<long long>(1, std::min<long long>(want, (long long)ctx->sm_count * per_sm));
  {
    constexpr int NG = 32 / G;
    const long long warps = grid * 4;
    long long g = A.n_pairs / (warps * 16);
    g = (g / NG) * NG;
    A.grab = A.n_pairs_dev ? 0 : (int)std::max<long long>(NG, std::min<long long>(8 * NG, g));
  }
  ctx->n_launches++;
  k_clip_tiny<G, NB, PT><<<(unsigned)grid, 128, smem, ctx->stream>>>(A);
  MB_CUDA(cudaGetLastError());
}

// Two passes in both modes: compact-caps pass (5 blocks / SM), then the cells it could not hold at the reference's
// caps.  A cell recomputed by the second pass goes through exactly the single-pass code path, so given-neighbours
// records (array positions, overflow statuses) stay byte-identical.
template <int G, bool PT>
static void launch_clip(mb_ctx* ctx, ClipArgs A) {
  A.redo_out = nullptr;
  A.work_list = nullptr;
  A.work_count = nullptr;
  ctx->redo_list.reserve((size_t)A.n_pairs + 1);
  A.redo_out = ctx->redo_list.p;
  if constexpr (G == 8 || G == 4) {
    // first pass: the batch-synchronous compact-caps kernel (rpd_clip2.cuh) in both modes; the state-machine kernel
    // k_clip keeps the opt-in security-radius exit (a9) and the A/B switch MB_CLIP_VARIANT=1
    if (ctx->clip_variant == 1 || A.security_radius)
      launch_clip_pass<G, PT, true>(ctx, A, false);
    else if (ctx->clip_variant == 2)
      launch_clip_tiny<G, 5, PT>(ctx, A);
    else
      launch_clip_tiny<G, 6, PT>(ctx, A);
  } else {
    launch_clip_pass<G, PT, true>(ctx, A, false);
  }
  A.redo_out = nullptr;
  A.work_list = ctx->redo_list.p;
  A.work_count = A.counters + CNT_REDO;
  A.n_pairs_dev = nullptr;
  launch_clip_pass<G, PT, false>(ctx, A, true);
}

// Scalars the host needs between stages (pair count, K3 counters, record words) are PUBLISHED by a tiny
// kernel into mapped pinned host memory instead of being fetched with cudaMemcpy: a D2H copy on the
// compute stream would queue behind the streamed run's bulk record copies on the copy engine and
// serialise the two streams.
__global__ void k_publish_words(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst_mapped, int n,
                                volatile unsigned long long* seq_mapped, unsigned long long seq) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst_mapped[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) *seq_mapped = seq;  // the host spins on this word (wait_published)
}

static HostScalars* host_scalars(mb_ctx* ctx);
static inline double now_us() {
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void publish(mb_ctx* ctx, const void* src, void* dst_mapped, size_t bytes) {
  HostScalars* hs = host_scalars(ctx);
  ctx->n_launches++;
  k_publish_words<<<1, 64, 0, ctx->stream>>>(reinterpret_cast<const uint32_t*>(src),
                                             reinterpret_cast<uint32_t*>(dst_mapped), (int)(bytes / 4), &hs->seq,
                                             ++ctx->publish_seq);
  MB_CUDA(cudaGetLastError());
}

// Wait for the last publish on the stream.  The publish kernel is the last operation enqueued before the wait and
// writes the sequence word after its payload (system-scope fence), so seeing the word means everything before
// it on the stream has completed.  Spinning on mapped memory wakes the host within a microsecond or two of the
// write crossing PCIe; cudaStreamSynchronize is interrupt-driven and costs tens of microseconds per wake-up,
// paid once per tet span.  Falls back to the runtime's wait after 2 ms of spinning (long kernels) and to report
// errors.
static void wait_published(mb_ctx* ctx) {
  HostScalars* hs = host_scalars(ctx);
  const volatile unsigned long long* seq = &hs->seq;
  const unsigned long long want = ctx->publish_seq;
  const double t0 = now_us();
  while (*seq != want) {
    if (now_us() - t0 > 2000.0) {
      MB_CUDA(cudaStreamSynchronize(ctx->stream));
      break;
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);  // payload reads stay behind the sequence-word read
  MB_CUDA(cudaGetLastError());
}

static HostScalars* host_scalars(mb_ctx* ctx) {
  if (!ctx->hs) {
    // slot 0: the one-at-a-time publishes; slots 1-2: the pipelined ranges of the staged streamed run
    MB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&ctx->hs), 4 * sizeof(HostScalars), cudaHostAllocMapped));
    memset(ctx->hs, 0, 4 * sizeof(HostScalars));
  }
  return ctx->hs;
}

static cudaEvent_t take_event(mb_ctx* ctx) {
  cudaEvent_t e = nullptr;
  if (!ctx->ev_pool.empty()) {
    e = ctx->ev_pool.back();
    ctx->ev_pool.pop_back();
  } else {
    MB_CUDA(cudaEventCreate(&e));
  }
  return e;
}

// host-side stage timers (MB_TRACE=1): where the host thread spends a span -- launching or waiting
#define TRACE(slot)                                  \
  do {                                               \
    if (ctx->trace_on) {                             \
      const double t_ = now_us();                    \
      ctx->trace_us[slot] += t_ - tr_;               \
      tr_ = t_;                                      \
    }                                                \
  } while (0)

static void trace_flush(mb_ctx* ctx, const char* what) {
  if (ctx->trace_level < 2) return;
  fprintf(stderr, "[libmat_b200 trace] %s host us: launchK2 %.0f waitK2 %.0f launchK3+scans %.0f waitK3 %.0f launchGather %.0f\n",
          what, ctx->trace_us[0], ctx->trace_us[1], ctx->trace_us[2], ctx->trace_us[3], ctx->trace_us[6]);
  for (double& v : ctx->trace_us) v = 0.0;
}

struct SpanStats {
  long long n_pairs = 0, n_cells = 0, total_words = 0;
};

// One span through K2 -> K3 -> ordering.  The ordered compact records go to `blob`, the per-cell byte
// offsets (+ base_bytes) to `cell_off`; counters are added to `res`.  Four timing events are appended
// to res->evs.  Returns with the ordering kernels enqueued (not synchronised).
static SpanStats rpd_run_span(mb_ctx* ctx, const mb_rpd_opts* opts, mb_rpd_result* res, const TetSpan& sp,
                              const GridDev* grid, DevBuf<uint32_t>& blob, DevBuf<long long>& cell_off,
                              long long base_bytes, int lean, bool force_sync = false) {
  TetMeshDev& M = ctx->mesh;
  SitesDev& S = ctx->sites;
  cudaStream_t s = ctx->stream;
  const int t_count = sp.count;
  const int G = opts && opts->lanes_per_cell ? opts->lanes_per_cell : 8;
  HostScalars* hs = host_scalars(ctx);
  double tr_ = ctx->trace_on ? now_us() : 0.0;
  cudaEvent_t ev[4];
  for (int i = 0; i < 4; i++) {
    ev[i] = take_event(ctx);
    res->evs.push_back(ev[i]);
  }
  ctx->counters.reserve(1);
  dev_zero(ctx, ctx->counters.p, sizeof(RpdCounters));
  MB_CUDA(cudaEventRecord(ev[0], s));

  // ---- K2: candidate (tet, site) pairs ---------------------------------------------------------
  long long n_pairs = 0;
  ctx->tet_cnt.reserve((size_t)t_count + 1);
  ctx->tet_off.reserve((size_t)t_count + 1);
  const bool grid_cands = grid != nullptr;
  const bool spec = grid_cands && !force_sync && t_count > 0 && ctx->pairs_per_tet_hint > 0.0;
  if (t_count > 0) {
    if (!grid_cands) {
      ctx->cand_pad.reserve((size_t)t_count * CAND_PAD);
      const int blocks = (t_count + 7) / 8;
      ctx->n_launches++;
      k_cand_given<false><<<blocks, 256, 0, s>>>(M.vert4.p, M.tet_idx.p, sp.first, t_count, sp.sel, S.site4.p,
                                                 S.flags.p, S.n_site, S.nbr.p, S.site_k, ctx->tet_cnt.p,
                                                 ctx->cand_pad.p, nullptr, nullptr, nullptr);
      MB_CUDA(cudaGetLastError());
    } else {
      grid_candidates(ctx, *grid, sp, opts ? opts->grid_k : 0);  // fills tet_cnt + cand_list pad
    }
    dev_zero(ctx, ctx->tet_cnt.p + t_count, sizeof(int));
    exclusive_scan<int>(ctx, ctx->tet_cnt.p, ctx->tet_off.p, (long long)t_count + 1);
    if (spec) {
      // SPECULATIVE: the pair count stays on the device; arrays are sized by the pairs-per-tet seen so far
      // (x1.5) and K3 / the ordering scan are launched without waiting.  The true count comes back with K3's
      // counters in the span's single synchronisation; a span that exceeded the capacity is redone.
      n_pairs = (long long)((double)t_count * ctx->pairs_per_tet_hint) + 4096;
      TRACE(0);
    } else {
      publish(ctx, ctx->tet_off.p + t_count, &hs->n_pairs, sizeof(int));
      TRACE(0);  // launch K2
      wait_published(ctx);
      TRACE(1);  // wait K2
      n_pairs = hs->n_pairs;
    }
    ctx->pair_tet.reserve((size_t)n_pairs + 1);
    ctx->pair_site.reserve((size_t)n_pairs + 1);
    ctx->pair_local.reserve((size_t)n_pairs + 1);
    if (n_pairs > 0) {
      if (!grid_cands) {
        const int blocks = (t_count + 7) / 8;
        ctx->n_launches++;
        k_cand_given<true><<<blocks, 256, 0, s>>>(M.vert4.p, M.tet_idx.p, sp.first, t_count, sp.sel, S.site4.p,
                                                  S.flags.p, S.n_site, S.nbr.p, S.site_k, ctx->tet_cnt.p,
                                                  ctx->cand_pad.p, ctx->tet_off.p, ctx->pair_tet.p,
                                                  ctx->pair_site.p);
        MB_CUDA(cudaGetLastError());
      } else {
        grid_fill_pairs(ctx, sp, n_pairs);
      }
    }
  }
  MB_CUDA(cudaEventRecord(ev[1], s));

  // ---- K3: clip ---------------------------------------------------------------------------------
  ctx->pair_status.reserve((size_t)n_pairs + 1);
  ctx->pair_blob.reserve((size_t)n_pairs + 1);
  ctx->pair_words.reserve((size_t)n_pairs + 1);
  // scratch: typical record ~80 words; retried with the exact need if it overflows
  size_t scratch_words = std::max<size_t>((size_t)n_pairs * 96 + (1u << 20), ctx->scratch.cap);
  RpdCounters hc;
  long long total_words = 0;
  DevBuf<long long>& word_off = ctx->word_off;  // packed: cell index << 40 | word offset
  for (int attempt = 0; attempt < 2 && n_pairs > 0; attempt++) {
    ctx->scratch.reserve(scratch_words);
    if (spec)  // entries beyond the true pair count are never written by K3: they must read as "no record"
      dev_zero(ctx, ctx->pair_words.p, sizeof(int) * (size_t)(n_pairs + 1));
    ClipArgs A;
    A.vert4 = M.vert4.p;
    A.tet_idx = M.tet_idx.p;
    A.tet_fadj = M.tet_fadj.p;
    A.tet_fid = M.tet_fid.p;
    A.tet_e6 = M.tet_e6.p;
    A.tet_geo = M.tet_geo.p;
    A.tet_vadj = M.tet_vadj.p;
    A.site4 = S.site4.p;
    A.n_site = S.n_site;
    if (S.given) {
      A.nbr = S.nbr.p;
      A.nbr_stride = S.site_k;
      A.nbr_cnt = nullptr;
    } else {
      A.nbr = ctx->cand_pad.p;
      A.nbr_stride = ctx->cand_kcap;
      A.nbr_cnt = ctx->cand_cnt.p;
    }
    A.tet_first = sp.first;
    A.tet_id_base = M.tet_id_base;
    A.pair_tet = ctx->pair_tet.p;
    A.pair_site = ctx->pair_site.p;
    A.pair_local = ctx->pair_local.p;
    A.n_pairs = n_pairs;
    A.n_pairs_dev = spec ? ctx->tet_off.p + t_count : nullptr;
    A.pair_status = ctx->pair_status.p;
    A.pair_blob = ctx->pair_blob.p;
    A.pair_words = ctx->pair_words.p;
    A.scratch = ctx->scratch.p;
    A.scratch_words = ctx->scratch.cap;
    A.counters = reinterpret_cast<unsigned long long*>(ctx->counters.p);
    A.no_cull = ctx->no_cull ? 1 : 0;
    A.security_radius = (opts && opts->security_radius && S.given) ? 1 : 0;
    const bool pt = A.nbr_cnt != nullptr;
    if (G == 4)
      pt ? launch_clip<4, true>(ctx, A) : launch_clip<4, false>(ctx, A);
    else if (G == 8)
      pt ? launch_clip<8, true>(ctx, A) : launch_clip<8, false>(ctx, A);
    else if (G == 16)
      pt ? launch_clip<16, true>(ctx, A) : launch_clip<16, false>(ctx, A);
    else
      pt ? launch_clip<32, true>(ctx, A) : launch_clip<32, false>(ctx, A);
    MB_CUDA(cudaEventRecord(ev[2], s));
    // ordering scans are enqueued right behind K3; K3's counters and the record total reach the host
    // with ONE synchronisation
    // one exclusive scan over packed (valid count << 40 | record words) gives every pair both its cell
    // index and its word offset in (tet, site) order
    word_off.reserve((size_t)n_pairs + 1);
    dev_zero(ctx, ctx->pair_words.p + n_pairs, sizeof(int));
    {
      unsigned long long* outp = reinterpret_cast<unsigned long long*>(word_off.p);
      size_t tmp = 0;
      ctx->n_launches += 2;
      if (lean == 2) {
        cub::TransformInputIterator<unsigned long long, PackWords<2>, const int*> in(ctx->pair_words.p, PackWords<2>());
        MB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, outp, n_pairs + 1, s));
        ctx->cub_tmp.reserve(tmp);
        MB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, in, outp, n_pairs + 1, s));
      } else if (lean == 1) {
        cub::TransformInputIterator<unsigned long long, PackWords<1>, const int*> in(ctx->pair_words.p, PackWords<1>());
        MB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, outp, n_pairs + 1, s));
        ctx->cub_tmp.reserve(tmp);
        MB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, in, outp, n_pairs + 1, s));
      } else {
        cub::TransformInputIterator<unsigned long long, PackWords<0>, const int*> in(ctx->pair_words.p, PackWords<0>());
        MB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, outp, n_pairs + 1, s));
        ctx->cub_tmp.reserve(tmp);
        MB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, in, outp, n_pairs + 1, s));
      }
    }
    ctx->n_launches++;
    k_publish_k3<<<1, 64, 0, s>>>(reinterpret_cast<const uint32_t*>(ctx->counters.p),
                                  reinterpret_cast<const unsigned long long*>(word_off.p) + n_pairs,
                                  spec ? ctx->tet_off.p + t_count : nullptr, hs, ++ctx->publish_seq);
    MB_CUDA(cudaGetLastError());
    TRACE(2);  // launch fill + K3 + ordering scans
    wait_published(ctx);
    TRACE(3);  // wait K3 + scans
    hc = hs->counters;
    total_words = hs->total_words & PACK_MASK;
    if (spec && (long long)hs->n_pairs > n_pairs) {
      // the speculative capacity was too small for this span: raise the estimate and redo it the safe way
      ctx->pairs_per_tet_hint = 1.5 * (double)hs->n_pairs / (double)t_count;
      for (int i = 0; i < 4; i++) {
        ctx->ev_pool.push_back(res->evs.back());
        res->evs.pop_back();
      }
      return rpd_run_span(ctx, opts, res, sp, grid, blob, cell_off, base_bytes, lean, true);
    }
    MB_REQUIRE(hc.n_valid <= PACK_MAX_CELLS && hc.blob_words <= PACK_MASK, MB_ERR_ARG,
               "more valid cells / record words in one tet span than the ordering scan can index: use the streamed run "
               "(mb_rpd_run_to_host) or a smaller tet range");
    if (hc.blob_words <= ctx->scratch.cap) break;
    // scratch too small: rerun K3 with the measured need (rare; statuses are recomputed)
    scratch_words = (size_t)hc.blob_words + (1u << 20);
    MB_REQUIRE(attempt == 0, MB_ERR_NOMEM, "compact scratch overflow after resize");
    {  // reset K3's counters only (the candidate-stage counters [4] and [16] stay)
      unsigned long long* c = reinterpret_cast<unsigned long long*>(ctx->counters.p);
      dev_zero(ctx, c, 4 * sizeof(unsigned long long));
      dev_zero(ctx, c + 5, 11 * sizeof(unsigned long long));
      dev_zero(ctx, c + CNT_WORK_CURSOR, sizeof(unsigned long long));
      dev_zero(ctx, c + 18, 5 * sizeof(unsigned long long));  // gc, redo count, second cursor, flagged pairs / cells
    }
    MB_CUDA(cudaEventRecord(ev[1], s));
  }
  if (n_pairs == 0) {
    if (t_count > 0) {  // candidate-stage counters of a span without pairs
      publish(ctx, ctx->counters.p, &hs->counters, sizeof(RpdCounters));
      wait_published(ctx);
      hc = hs->counters;
    } else {
      memset(&hc, 0, sizeof hc);
    }
    MB_CUDA(cudaEventRecord(ev[2], s));
  }
  if (spec) n_pairs = hs->n_pairs;  // the true count (<= capacity)
  if (grid_cands && t_count > 0)
    ctx->pairs_per_tet_hint = std::max(ctx->pairs_per_tet_hint, 1.5 * (double)n_pairs / (double)t_count);
  SpanStats st;
  st.n_pairs = n_pairs;
  st.n_cells = (long long)hc.n_valid;
  res->n_pairs += (long)n_pairs;
  res->n_cells += (long)hc.n_valid;
  res->n_clips += (long)hc.n_clips;
  res->n_culled += (long)hc.n_culled;
  res->n_exact += (long)hc.pad[0];
  res->n_cand_overflow += (long)hc.n_cand_overflow;
  res->n_ovf_tets += (long)hc.n_ovf_tets;
  res->n_redo += (long)hc.n_redo;
  res->n_gc += (long)hc.n_gc;
  res->n_flag_pairs += (long)hc.reserved[0];
  res->n_flag_cells += (long)hc.reserved[1];
  if (ctx->trace_level >= 2)
    fprintf(stderr, "[libmat_b200 trace] span of %d tets: %lld pairs, %llu tets handed back by the cluster search, %llu to the big-list pass, %llu redo cells\n",
            t_count, n_pairs, hc.reserved[2], hc.n_ovf_tets, hc.n_redo);
  for (int i = 0; i < 10; i++) res->hist[i] += (long)hc.hist[i];

  // ---- gather into (tet, site) order -------------------------------------------------------------
  cell_off.reserve((size_t)st.n_cells + 1);
  if (n_pairs > 0) {
    blob.reserve((size_t)total_words + 4);
    if (total_words > 0) {
      ctx->n_launches++;
      if (lean == 2)
        k_gather<2><<<(unsigned)((n_pairs * GATHER_LANES + 255) / 256), 256, 0, s>>>(
            ctx->scratch.p, ctx->pair_blob.p, ctx->pair_words.p, reinterpret_cast<const unsigned long long*>(word_off.p),
            n_pairs, blob.p, cell_off.p, total_words, st.n_cells, base_bytes);
      else if (lean == 1)
        k_gather<1><<<(unsigned)((n_pairs * GATHER_LANES + 255) / 256), 256, 0, s>>>(
            ctx->scratch.p, ctx->pair_blob.p, ctx->pair_words.p, reinterpret_cast<const unsigned long long*>(word_off.p),
            n_pairs, blob.p, cell_off.p, total_words, st.n_cells, base_bytes);
      else
        k_gather<0><<<(unsigned)((n_pairs * GATHER_LANES + 255) / 256), 256, 0, s>>>(
            ctx->scratch.p, ctx->pair_blob.p, ctx->pair_words.p, reinterpret_cast<const unsigned long long*>(word_off.p),
            n_pairs, blob.p, cell_off.p, total_words, st.n_cells, base_bytes);
      MB_CUDA(cudaGetLastError());
    }
  }
  if (st.n_cells == 0) MB_CUDA(cudaMemcpyAsync(cell_off.p, &base_bytes, sizeof(long long), cudaMemcpyHostToDevice, s));
  st.total_words = total_words;
  MB_CUDA(cudaEventRecord(ev[3], s));
  TRACE(6);  // launch gather
  return st;
}

// ---------------------------------------------------------------------------------------------------------------
// Staged streamed run (grid candidates).  Every launch of a small kernel costs ~7 us of stream time here (launch
// throughput, measured), and a tet span that goes through K2 -> K3 -> ordering takes 16 launches: three spans cost
// 0.5 ms more than one.  So the candidate stage runs ONCE for all tets of the run (rpd_candidates_all: K2, the pair
// offsets, the pair arrays; the pair offsets at the span cuts come back in one synchronisation) and only K3 + the
// ordering are cut into spans -- pair ranges that end on tet boundaries (rpd_clip_range: 8 launches, one wait).
// ---------------------------------------------------------------------------------------------------------------
static long long rpd_candidates_all(mb_ctx* ctx, const mb_rpd_opts* opts, mb_rpd_result* res, const TetSpan& sp,
                                    const GridDev& grid, const std::vector<int>& cut_tets, std::vector<long long>& cut_pairs) {
  cudaStream_t s = ctx->stream;
  const int t_count = sp.count;
  HostScalars* hs = host_scalars(ctx);
  double tr_ = ctx->trace_on ? now_us() : 0.0;
  cudaEvent_t ev[4];
  for (int i = 0; i < 4; i++) {
    ev[i] = take_event(ctx);
    res->evs.push_back(ev[i]);
  }
  ctx->counters.reserve(1);
  dev_zero(ctx, ctx->counters.p, sizeof(RpdCounters));
  MB_CUDA(cudaEventRecord(ev[0], s));
  ctx->tet_cnt.reserve((size_t)t_count + 1);
  ctx->tet_off.reserve((size_t)t_count + 1);
  grid_candidates(ctx, grid, sp, opts ? opts->grid_k : 0);
  dev_zero(ctx, ctx->tet_cnt.p + t_count, sizeof(int));
  exclusive_scan<int>(ctx, ctx->tet_cnt.p, ctx->tet_off.p, (long long)t_count + 1);
  CutList cl;
  cl.n = (int)cut_tets.size();
  MB_REQUIRE(cl.n >= 2 && cl.n <= 36, MB_ERR_ARG, "bad span count");
  for (int i = 0; i < cl.n; i++) cl.tet[i] = cut_tets[i];
  ctx->n_launches++;
  k_publish_cuts<<<1, 64, 0, s>>>(ctx->tet_off.p, cl, reinterpret_cast<const uint32_t*>(ctx->counters.p), hs, ++ctx->publish_seq);
  MB_CUDA(cudaGetLastError());
  TRACE(0);
  wait_published(ctx);
  TRACE(1);
  cut_pairs.assign(hs->cut_off, hs->cut_off + cl.n);
  const long long n_pairs = cut_pairs.back();
  const RpdCounters hc = hs->counters;
  res->n_cand_overflow += (long)hc.n_cand_overflow;
  res->n_ovf_tets += (long)hc.n_ovf_tets;
  if (ctx->trace_level >= 2)
    fprintf(stderr, "[libmat_b200 trace] candidates of %d tets: %lld pairs, %llu tets handed back by the cluster search, %llu to the big-list pass\n",
            t_count, n_pairs, hc.reserved[2], hc.n_ovf_tets);
  ctx->pair_tet.reserve((size_t)n_pairs + 1);
  ctx->pair_site.reserve((size_t)n_pairs + 1);
  ctx->pair_local.reserve((size_t)n_pairs + 1);
  ctx->pair_status.reserve((size_t)n_pairs + 1);
  ctx->pair_blob.reserve((size_t)n_pairs + 1);
  ctx->pair_words.reserve((size_t)n_pairs + 1);
  ctx->word_off.reserve((size_t)n_pairs + cl.n + 2);
  if (n_pairs > 0) grid_fill_pairs(ctx, sp, n_pairs);
  for (int i = 1; i < 4; i++) MB_CUDA(cudaEventRecord(ev[i], s));
  if (t_count > 0)
    ctx->pairs_per_tet_hint = std::max(ctx->pairs_per_tet_hint, 1.5 * (double)n_pairs / (double)t_count);
  res->n_pairs += (long)n_pairs;
  return n_pairs;
}

// K3 + ordering of the pairs [p0, p1) of the run prepared by rpd_candidates_all (range index c of the run)
static SpanStats rpd_clip_range(mb_ctx* ctx, const mb_rpd_opts* opts, mb_rpd_result* res, const TetSpan& sp, long long p0,
                                long long p1, int c, DevBuf<uint32_t>& blob, DevBuf<long long>& cell_off,
                                long long base_bytes, int lean) {
  TetMeshDev& M = ctx->mesh;
  SitesDev& S = ctx->sites;
  cudaStream_t s = ctx->stream;
  const long long n_pairs = p1 - p0;
  const int G = opts && opts->lanes_per_cell ? opts->lanes_per_cell : 8;
  HostScalars* hs = host_scalars(ctx);
  double tr_ = ctx->trace_on ? now_us() : 0.0;
  cudaEvent_t ev[4];
  for (int i = 0; i < 4; i++) {
    ev[i] = take_event(ctx);
    res->evs.push_back(ev[i]);
  }
  MB_CUDA(cudaEventRecord(ev[0], s));
  MB_CUDA(cudaEventRecord(ev[1], s));
  SpanStats st;
  st.n_pairs = n_pairs;
  cell_off.reserve(1);
  if (n_pairs == 0) {
    MB_CUDA(cudaEventRecord(ev[2], s));
    MB_CUDA(cudaMemcpyAsync(cell_off.p, &base_bytes, sizeof(long long), cudaMemcpyHostToDevice, s));
    MB_CUDA(cudaEventRecord(ev[3], s));
    return st;
  }
  // scratch this range may use: typical record ~80 words; retried with the exact need if it overflows.  (Not the whole
  // of ctx->scratch, which a one-shot run may have grown to the size of the full result: the range's destination
  // buffer is sized by this bound.)
  // (+ the bump allocator's granularity: every cell group of the persistent grid holds a partly used 4 KB chunk)
  size_t scratch_words = (size_t)n_pairs * 96 + (size_t)ctx->sm_count * 8 * 16 * CLIP_CHUNK_WORDS + (1u << 20);
  RpdCounters hc;
  long long total_words = 0;
  unsigned long long* word_off = reinterpret_cast<unsigned long long*>(ctx->word_off.p) + p0 + c;  // own slot per range
  for (int attempt = 0; attempt < 2; attempt++) {
    ctx->scratch.reserve(scratch_words);
    dev_zero(ctx, ctx->counters.p, sizeof(RpdCounters));
    ClipArgs A;
    A.vert4 = M.vert4.p;
    A.tet_idx = M.tet_idx.p;
    A.tet_fadj = M.tet_fadj.p;
    A.tet_fid = M.tet_fid.p;
    A.tet_e6 = M.tet_e6.p;
    A.tet_geo = M.tet_geo.p;
    A.tet_vadj = M.tet_vadj.p;
    A.site4 = S.site4.p;
    A.n_site = S.n_site;
    if (S.given) {  // given-neighbours semantics with pairs from the grid search: the listed neighbours, in list order
      A.nbr = S.nbr.p;
      A.nbr_stride = S.site_k;
      A.nbr_cnt = nullptr;
    } else {
      A.nbr = ctx->cand_pad.p;
      A.nbr_stride = ctx->cand_kcap;
      A.nbr_cnt = ctx->cand_cnt.p;
    }
    A.tet_first = sp.first;
    A.tet_id_base = M.tet_id_base;
    A.pair_tet = ctx->pair_tet.p + p0;
    A.pair_site = ctx->pair_site.p + p0;
    A.pair_local = ctx->pair_local.p + p0;
    A.n_pairs = n_pairs;
    A.n_pairs_dev = nullptr;
    A.pair_status = ctx->pair_status.p + p0;
    A.pair_blob = ctx->pair_blob.p + p0;
    A.pair_words = ctx->pair_words.p + p0;
    A.scratch = ctx->scratch.p;
    A.scratch_words = scratch_words;
    A.counters = reinterpret_cast<unsigned long long*>(ctx->counters.p);
    A.no_cull = ctx->no_cull ? 1 : 0;
    A.security_radius = 0;
    const bool pt = A.nbr_cnt != nullptr;
    if (G == 4)
      pt ? launch_clip<4, true>(ctx, A) : launch_clip<4, false>(ctx, A);
    else if (G == 8)
      pt ? launch_clip<8, true>(ctx, A) : launch_clip<8, false>(ctx, A);
    else if (G == 16)
      pt ? launch_clip<16, true>(ctx, A) : launch_clip<16, false>(ctx, A);
    else
      pt ? launch_clip<32, true>(ctx, A) : launch_clip<32, false>(ctx, A);
    MB_CUDA(cudaEventRecord(ev[2], s));
    {
      // exclusive scan over packed (valid count | record words): out[n_pairs] (the total) does not depend on in[n_pairs]
      size_t tmp = 0;
      ctx->n_launches += 2;
      const int* in_words = ctx->pair_words.p + p0;
      if (lean == 2) {
        cub::TransformInputIterator<unsigned long long, PackWords<2>, const int*> in(in_words, PackWords<2>());
        MB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, word_off, n_pairs + 1, s));
        ctx->cub_tmp.reserve(tmp);
        MB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, in, word_off, n_pairs + 1, s));
      } else if (lean == 1) {
        cub::TransformInputIterator<unsigned long long, PackWords<1>, const int*> in(in_words, PackWords<1>());
        MB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, word_off, n_pairs + 1, s));
        ctx->cub_tmp.reserve(tmp);
        MB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, in, word_off, n_pairs + 1, s));
      } else {
        cub::TransformInputIterator<unsigned long long, PackWords<0>, const int*> in(in_words, PackWords<0>());
        MB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, word_off, n_pairs + 1, s));
        ctx->cub_tmp.reserve(tmp);
        MB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, in, word_off, n_pairs + 1, s));
      }
    }
    ctx->n_launches++;
    k_publish_k3<<<1, 64, 0, s>>>(reinterpret_cast<const uint32_t*>(ctx->counters.p), word_off + n_pairs, nullptr, hs,
                                  ++ctx->publish_seq);
    MB_CUDA(cudaGetLastError());
    // the gather goes out BEFORE the host has seen the totals (they stay on the device: total_dev): the stream does not
    // idle through the host's wake-up, which takes up to ~0.1 ms while a bulk D2H copy is crossing PCIe.  Its output is
    // bounded by the scratch K3 wrote, so the destination is sized by the scratch capacity.
    cell_off.reserve((size_t)n_pairs + 1);
    blob.reserve(scratch_words + 4);
    {
      ctx->n_launches++;
      const unsigned nb = (unsigned)((n_pairs * GATHER_LANES + 255) / 256);
      if (lean == 2)
        k_gather<2><<<nb, 256, 0, s>>>(ctx->scratch.p, ctx->pair_blob.p + p0, ctx->pair_words.p + p0, word_off, n_pairs, blob.p,
                                       cell_off.p, 0, 0, base_bytes, word_off + n_pairs);
      else if (lean == 1)
        k_gather<1><<<nb, 256, 0, s>>>(ctx->scratch.p, ctx->pair_blob.p + p0, ctx->pair_words.p + p0, word_off, n_pairs, blob.p,
                                       cell_off.p, 0, 0, base_bytes, word_off + n_pairs);
      else
        k_gather<0><<<nb, 256, 0, s>>>(ctx->scratch.p, ctx->pair_blob.p + p0, ctx->pair_words.p + p0, word_off, n_pairs, blob.p,
                                       cell_off.p, 0, 0, base_bytes, word_off + n_pairs);
      MB_CUDA(cudaGetLastError());
    }
    MB_CUDA(cudaEventRecord(ev[3], s));
    TRACE(2);
    wait_published(ctx);
    TRACE(3);
    hc = hs->counters;
    total_words = hs->total_words & PACK_MASK;
    MB_REQUIRE(hc.n_valid <= PACK_MAX_CELLS && hc.blob_words <= PACK_MASK, MB_ERR_ARG,
               "more valid cells / record words in one span than the ordering scan can index: use more spans");
    if (hc.blob_words <= scratch_words) break;
    scratch_words = (size_t)hc.blob_words + (1u << 20);  // scratch too small: K3 + ordering + gather again (rare)
    MB_REQUIRE(attempt == 0, MB_ERR_NOMEM, "compact scratch overflow after resize");
  }
  st.n_cells = (long long)hc.n_valid;
  st.total_words = total_words;
  res->n_cells += (long)hc.n_valid;
  res->n_clips += (long)hc.n_clips;
  res->n_culled += (long)hc.n_culled;
  res->n_exact += (long)hc.pad[0];
  res->n_redo += (long)hc.n_redo;
  res->n_gc += (long)hc.n_gc;
  res->n_flag_pairs += (long)hc.reserved[0];
  res->n_flag_cells += (long)hc.reserved[1];
  for (int i = 0; i < 10; i++) res->hist[i] += (long)hc.hist[i];
  if (st.n_cells == 0) {
    MB_CUDA(cudaMemcpyAsync(cell_off.p, &base_bytes, sizeof(long long), cudaMemcpyHostToDevice, s));
    MB_CUDA(cudaEventRecord(ev[3], s));
  }
  TRACE(6);
  return st;
}

// ---------------------------------------------------------------------------------------------------------------
// Pipelined form of rpd_clip_range: range c+1 is ENQUEUED before the host waits for range c, so the stream never idles
// through a host round trip and the launches of the small kernels are issued while the previous range's K3 runs.
// What the host used to hand over between ranges now stays on the device: the base offset of a range's cell offsets
// is the device accumulator of the earlier ranges' record words (k_publish_range adds, k_gather reads), the totals
// reach the host through one of two mapped slots.  A range that overflows its scratch bound makes range_collect
// return false: the caller drains the stream and redoes the run with the one-range-at-a-time path (rare).
// ---------------------------------------------------------------------------------------------------------------
struct RangeJob {
  long long p0 = 0, p1 = 0;
  int c = 0;
  size_t scratch_words = 0;
  unsigned long long seq = 0;
  HostScalars* slot = nullptr;
};

static RangeJob range_enqueue(mb_ctx* ctx, const mb_rpd_opts* opts, mb_rpd_result* res, const TetSpan& sp, long long p0,
                              long long p1, int c, DevBuf<uint32_t>& blob, DevBuf<long long>& cell_off, int lean) {
  TetMeshDev& M = ctx->mesh;
  SitesDev& S = ctx->sites;
  cudaStream_t s = ctx->stream;
  RangeJob J;
  J.p0 = p0;
  J.p1 = p1;
  J.c = c;
  const long long n_pairs = p1 - p0;
  const int G = opts && opts->lanes_per_cell ? opts->lanes_per_cell : 8;
  J.slot = host_scalars(ctx) + 1 + (c & 1);
  double tr_ = ctx->trace_on ? now_us() : 0.0;
  cudaEvent_t ev[4];
  for (int i = 0; i < 4; i++) {
    ev[i] = take_event(ctx);
    res->evs.push_back(ev[i]);
  }
  MB_CUDA(cudaEventRecord(ev[0], s));
  MB_CUDA(cudaEventRecord(ev[1], s));
  unsigned long long* acc = reinterpret_cast<unsigned long long*>(ctx->acc_words.p);
  dev_zero(ctx, ctx->counters.p, sizeof(RpdCounters));
  if (n_pairs == 0) {
    cell_off.reserve(1);
    MB_CUDA(cudaEventRecord(ev[2], s));
    ctx->n_launches += 2;
    k_base_offset<<<1, 1, 0, s>>>(acc, cell_off.p);
    J.seq = ++ctx->publish_seq;
    k_publish_range<<<1, 64, 0, s>>>(reinterpret_cast<const uint32_t*>(ctx->counters.p), nullptr, acc, J.slot, J.seq);
    MB_CUDA(cudaGetLastError());
    MB_CUDA(cudaEventRecord(ev[3], s));
    return J;
  }
  J.scratch_words = (size_t)n_pairs * 96 + (size_t)ctx->sm_count * 8 * 16 * CLIP_CHUNK_WORDS + (1u << 20);
  if (ctx->debug_small_scratch) J.scratch_words = (size_t)n_pairs * 8 + 4096;  // tests: force the overflow fallback
  ctx->scratch.reserve(J.scratch_words);
  unsigned long long* word_off = reinterpret_cast<unsigned long long*>(ctx->word_off.p) + p0 + c;
  ClipArgs A;
  A.vert4 = M.vert4.p;
  A.tet_idx = M.tet_idx.p;
  A.tet_fadj = M.tet_fadj.p;
  A.tet_fid = M.tet_fid.p;
  A.tet_e6 = M.tet_e6.p;
  A.tet_geo = M.tet_geo.p;
  A.tet_vadj = M.tet_vadj.p;
  A.site4 = S.site4.p;
  A.n_site = S.n_site;
  if (S.given) {
    A.nbr = S.nbr.p;
    A.nbr_stride = S.site_k;
    A.nbr_cnt = nullptr;
  } else {
    A.nbr = ctx->cand_pad.p;
    A.nbr_stride = ctx->cand_kcap;
    A.nbr_cnt = ctx->cand_cnt.p;
  }
  A.tet_first = sp.first;
  A.tet_id_base = M.tet_id_base;
  A.pair_tet = ctx->pair_tet.p + p0;
  A.pair_site = ctx->pair_site.p + p0;
  A.pair_local = ctx->pair_local.p + p0;
  A.n_pairs = n_pairs;
  A.n_pairs_dev = nullptr;
  A.pair_status = ctx->pair_status.p + p0;
  A.pair_blob = ctx->pair_blob.p + p0;
  A.pair_words = ctx->pair_words.p + p0;
  A.scratch = ctx->scratch.p;
  A.scratch_words = J.scratch_words;
  A.counters = reinterpret_cast<unsigned long long*>(ctx->counters.p);
  A.no_cull = ctx->no_cull ? 1 : 0;
  A.security_radius = 0;
  const bool pt = A.nbr_cnt != nullptr;
  if (G == 4)
    pt ? launch_clip<4, true>(ctx, A) : launch_clip<4, false>(ctx, A);
  else if (G == 8)
    pt ? launch_clip<8, true>(ctx, A) : launch_clip<8, false>(ctx, A);
  else if (G == 16)
    pt ? launch_clip<16, true>(ctx, A) : launch_clip<16, false>(ctx, A);
  else
    pt ? launch_clip<32, true>(ctx, A) : launch_clip<32, false>(ctx, A);
  MB_CUDA(cudaEventRecord(ev[2], s));
  {
    size_t tmp = 0;
    ctx->n_launches += 2;
    const int* in_words = ctx->pair_words.p + p0;
    if (lean == 2) {
      cub::TransformInputIterator<unsigned long long, PackWords<2>, const int*> in(in_words, PackWords<2>());
      MB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, word_off, n_pairs + 1, s));
      ctx->cub_tmp.reserve(tmp);
      MB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, in, word_off, n_pairs + 1, s));
    } else if (lean == 1) {
      cub::TransformInputIterator<unsigned long long, PackWords<1>, const int*> in(in_words, PackWords<1>());
      MB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, word_off, n_pairs + 1, s));
      ctx->cub_tmp.reserve(tmp);
      MB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, in, word_off, n_pairs + 1, s));
    } else {
      cub::TransformInputIterator<unsigned long long, PackWords<0>, const int*> in(in_words, PackWords<0>());
      MB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, word_off, n_pairs + 1, s));
      ctx->cub_tmp.reserve(tmp);
      MB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, in, word_off, n_pairs + 1, s));
    }
  }
  cell_off.reserve((size_t)n_pairs + 1);
  blob.reserve(J.scratch_words + 4);
  ctx->n_launches += 3;
  k_base_offset<<<1, 1, 0, s>>>(acc, cell_off.p);  // a range without valid cells still has its first offset
  const unsigned nb = (unsigned)((n_pairs * GATHER_LANES + 255) / 256);
  if (lean == 2)
    k_gather<2><<<nb, 256, 0, s>>>(ctx->scratch.p, ctx->pair_blob.p + p0, ctx->pair_words.p + p0, word_off, n_pairs, blob.p,
                                   cell_off.p, 0, 0, 0, word_off + n_pairs, acc);
  else if (lean == 1)
    k_gather<1><<<nb, 256, 0, s>>>(ctx->scratch.p, ctx->pair_blob.p + p0, ctx->pair_words.p + p0, word_off, n_pairs, blob.p,
                                   cell_off.p, 0, 0, 0, word_off + n_pairs, acc);
  else
    k_gather<0><<<nb, 256, 0, s>>>(ctx->scratch.p, ctx->pair_blob.p + p0, ctx->pair_words.p + p0, word_off, n_pairs, blob.p,
                                   cell_off.p, 0, 0, 0, word_off + n_pairs, acc);
  J.seq = ++ctx->publish_seq;
  k_publish_range<<<1, 64, 0, s>>>(reinterpret_cast<const uint32_t*>(ctx->counters.p), word_off + n_pairs, acc, J.slot, J.seq);
  MB_CUDA(cudaGetLastError());
  MB_CUDA(cudaEventRecord(ev[3], s));
  TRACE(2);
  return J;
}

// waits for the range's publish; false = the range overflowed its scratch bound (nothing was accumulated)
static bool range_collect(mb_ctx* ctx, mb_rpd_result* res, const RangeJob& J, SpanStats& st) {
  double tr_ = ctx->trace_on ? now_us() : 0.0;
  const volatile unsigned long long* seq = &J.slot->seq;
  const double t0 = now_us();
  while (*seq != J.seq) {
    if (now_us() - t0 > 20000.0) {
      MB_CUDA(cudaStreamSynchronize(ctx->stream));
      break;
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  MB_CUDA(cudaGetLastError());
  TRACE(3);
  const RpdCounters hc = J.slot->counters;
  const long long total_words = J.slot->total_words & PACK_MASK;
  st = SpanStats();
  st.n_pairs = J.p1 - J.p0;
  if (st.n_pairs == 0) return true;
  MB_REQUIRE(hc.n_valid <= PACK_MAX_CELLS && hc.blob_words <= PACK_MASK, MB_ERR_ARG,
             "more valid cells / record words in one span than the ordering scan can index: use more spans");
  if (hc.blob_words > J.scratch_words) return false;
  st.n_cells = (long long)hc.n_valid;
  st.total_words = total_words;
  res->n_cells += (long)hc.n_valid;
  res->n_clips += (long)hc.n_clips;
  res->n_culled += (long)hc.n_culled;
  res->n_exact += (long)hc.pad[0];
  res->n_redo += (long)hc.n_redo;
  res->n_gc += (long)hc.n_gc;
  res->n_flag_pairs += (long)hc.reserved[0];
  res->n_flag_cells += (long)hc.reserved[1];
  for (int i = 0; i < 10; i++) res->hist[i] += (long)hc.hist[i];
  return true;
}

static void run_prologue(mb_ctx* ctx, const mb_rpd_opts* opts, mb_rpd_result* res, int& t_first, int& t_count) {
  TetMeshDev& M = ctx->mesh;
  SitesDev& S = ctx->sites;
  MB_REQUIRE(M.n_tet > 0, MB_ERR_STATE, "mb_set_tetmesh must be called before mb_rpd3d");
  MB_REQUIRE(S.n_site > 0, MB_ERR_STATE, "no sites uploaded");
  t_first = M.n_sel > 0 ? 0 : M.range_first;
  t_count = M.n_sel > 0 ? M.n_sel : (M.range_count < 0 ? M.n_tet - t_first : M.range_count);
  MB_REQUIRE(t_first >= 0 && t_count >= 0 && (M.n_sel > 0 || t_first + t_count <= M.n_tet), MB_ERR_ARG, "bad tet range");
  const int G = opts && opts->lanes_per_cell ? opts->lanes_per_cell : 8;
  MB_REQUIRE(G == 4 || G == 8 || G == 16 || G == 32, MB_ERR_ARG, "lanes_per_cell must be 4, 8, 16 or 32");
  res->ctx = ctx;
  res->n_site = S.n_site;
  res->want_volumes = opts && opts->want_volumes;
  res->n_pairs = res->n_cells = res->n_clips = res->n_culled = res->n_exact = 0;
  res->n_cand_overflow = res->n_ovf_tets = 0;
  res->n_redo = res->n_gc = 0;
  res->n_flag_pairs = res->n_flag_cells = 0;
  for (int i = 0; i < 10; i++) res->hist[i] = 0;
}

// =============================================================================================
// f2 / config 5: incremental recompute without neighbour rings.
//   reference: RPD3D_GPU::calculate_partial (src/rpd3d_api/rpd_api.cxx:147-313) rebuilds the CGAL regular triangulation,
//   takes the changed spheres + their 1-ring (+ 2-ring as clip-only sites, triangulation.cxx:442-549), recomputes the
//   tets of those spheres' previous cells (load_partial_tet_given_spheres :482-535) and merges (:432-479).
//   here: the cells of a tet are a function of the tet and of its candidate list (site ids + their centres / weights /
//   flags) -- nothing else.  K2 costs a fraction of K3, so the candidate lists of ALL tets are recomputed with the new
//   sites and compared with the previous run's: a tet is AFFECTED iff its list differs or lists a changed site.  That
//   set is exact (every other tet's records are bit-for-bit those of the previous run) and needs no triangulation.
// =============================================================================================
__global__ void k_inc_affected(int n_tet, int kcap, const int* __restrict__ cnt, const int* __restrict__ pad,
                               const int* __restrict__ pcnt, const int* __restrict__ ppad,
                               const float4* __restrict__ site4, const unsigned* __restrict__ flags,
                               const float4* __restrict__ psite4, const unsigned* __restrict__ pflags, int n_prev_site,
                               int* __restrict__ flag_out) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= n_tet) return;
  const int n = cnt[t];
  bool diff = n != pcnt[t];
  for (int i = lane; i < n && !diff; i += 32) {
    const int s = pad[(size_t)t * kcap + i];
    if (s != ppad[(size_t)t * kcap + i] || s >= n_prev_site) {
      diff = true;
    } else {
      const float4 a = site4[s], b = psite4[s];
      diff = __float_as_uint(a.x) != __float_as_uint(b.x) || __float_as_uint(a.y) != __float_as_uint(b.y) ||
             __float_as_uint(a.z) != __float_as_uint(b.z) || __float_as_uint(a.w) != __float_as_uint(b.w) ||
             flags[s] != pflags[s];
    }
  }
  const bool any = __any_sync(0xffffffffu, diff);
  if (lane == 0) flag_out[t] = any ? 1 : 0;
}

__global__ void k_inc_compact(int n_tet, const int* __restrict__ flag, const int* __restrict__ pos, int* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_tet && flag[t]) out[pos[t]] = t;
}

int rpd_incremental_select(mb_ctx* ctx, const mb_rpd_opts* opts) {
  TetMeshDev& M = ctx->mesh;
  SitesDev& S = ctx->sites;
  cudaStream_t s = ctx->stream;
  MB_REQUIRE(M.n_tet > 0 && S.n_site > 0, MB_ERR_STATE, "mesh and sites first");
  MB_REQUIRE(!S.given, MB_ERR_STATE, "the incremental recompute runs in grid-kNN mode (upload the sites without site_knn)");
  M.n_sel = 0;
  M.range_first = 0;
  M.range_count = -1;
  const int n_tet = M.n_tet;
  // candidate lists of ALL tets with the new sites (K1 + K2)
  ctx->counters.reserve(1);
  dev_zero(ctx, ctx->counters.p, sizeof(RpdCounters));
  ctx->tet_cnt.reserve((size_t)n_tet + 1);
  const GridDev G = grid_build(ctx);
  const TetSpan all = {0, n_tet, nullptr};
  grid_candidates(ctx, G, all, opts ? opts->grid_k : 0);
  const int kcap = ctx->cand_kcap;
  const bool have_prev = ctx->inc_valid && ctx->inc_n_tet == n_tet && ctx->inc_kcap == kcap;
  int n_aff = n_tet;
  ctx->inc_affected.reserve((size_t)n_tet + 1);
  if (have_prev) {
    ctx->inc_flag.reserve((size_t)n_tet + 1);
    ctx->inc_pos.reserve((size_t)n_tet + 1);
    ctx->n_launches += 4;
    k_inc_affected<<<(unsigned)(((size_t)n_tet * 32 + 255) / 256), 256, 0, s>>>(
        n_tet, kcap, ctx->cand_cnt.p, ctx->cand_pad.p, ctx->inc_cand_cnt.p, ctx->inc_cand_pad.p, S.site4.p, S.flags.p,
        ctx->inc_site4.p, ctx->inc_flags.p, ctx->inc_n_site, ctx->inc_flag.p);
    dev_zero(ctx, ctx->inc_flag.p + n_tet, sizeof(int));
    exclusive_scan<int>(ctx, ctx->inc_flag.p, ctx->inc_pos.p, (long long)n_tet + 1);
    k_inc_compact<<<(n_tet + 255) / 256, 256, 0, s>>>(n_tet, ctx->inc_flag.p, ctx->inc_pos.p, ctx->inc_affected.p);
    MB_CUDA(cudaGetLastError());
    MB_CUDA(cudaMemcpyAsync(&n_aff, ctx->inc_pos.p + n_tet, sizeof(int), cudaMemcpyDeviceToHost, s));
  }
  // this run's lists and sites become the reference point of the next one
  ctx->inc_cand_pad.reserve((size_t)n_tet * kcap);
  ctx->inc_cand_cnt.reserve((size_t)n_tet + 1);
  ctx->inc_site4.reserve((size_t)S.n_site);
  ctx->inc_flags.reserve((size_t)S.n_site);
  // (the comparison above has been enqueued before these copies on the same stream)
  MB_CUDA(cudaMemcpyAsync(ctx->inc_cand_pad.p, ctx->cand_pad.p, sizeof(int) * (size_t)n_tet * kcap, cudaMemcpyDeviceToDevice, s));
  MB_CUDA(cudaMemcpyAsync(ctx->inc_cand_cnt.p, ctx->cand_cnt.p, sizeof(int) * (size_t)n_tet, cudaMemcpyDeviceToDevice, s));
  MB_CUDA(cudaMemcpyAsync(ctx->inc_site4.p, S.site4.p, sizeof(float4) * (size_t)S.n_site, cudaMemcpyDeviceToDevice, s));
  MB_CUDA(cudaMemcpyAsync(ctx->inc_flags.p, S.flags.p, sizeof(unsigned) * (size_t)S.n_site, cudaMemcpyDeviceToDevice, s));
  MB_CUDA(cudaStreamSynchronize(s));
  ctx->inc_n_tet = n_tet;
  ctx->inc_n_site = S.n_site;
  ctx->inc_kcap = kcap;
  ctx->inc_valid = true;
  ctx->inc_n_affected = n_aff;
  if (have_prev) {
    // the affected tets become the context's subset (device-resident list, ascending)
    M.tet_sel.reserve((size_t)std::max(n_aff, 1));
    if (n_aff > 0) MB_CUDA(cudaMemcpyAsync(M.tet_sel.p, ctx->inc_affected.p, sizeof(int) * (size_t)n_aff, cudaMemcpyDeviceToDevice, s));
    MB_CUDA(cudaStreamSynchronize(s));
    M.n_sel = n_aff;
  }
  return n_aff;
}

void rpd_run(mb_ctx* ctx, const mb_rpd_opts* opts, mb_rpd_result* res) {
  int t_first, t_count;
  run_prologue(ctx, opts, res, t_first, t_count);
  const bool grid_cands = !ctx->sites.given || (opts && opts->grid_candidates);
  GridDev G;
  // K1 is timed with the span's candidate stage: the span's first event is recorded after it, so
  // record an extra leading event here
  cudaEvent_t e0 = take_event(ctx);
  MB_CUDA(cudaEventRecord(e0, ctx->stream));
  if (grid_cands && t_count > 0) G = grid_build(ctx);
  const TetSpan sp = {t_first, t_count, ctx->mesh.sel_ptr()};
  MB_REQUIRE(!(opts && opts->lean_records), MB_ERR_ARG,
             "lean_records is a transport format of the streamed runs (mb_rpd_run_to_host / mb_rpd_run_to_sink)");
  const SpanStats st = rpd_run_span(ctx, opts, res, sp, (grid_cands && t_count > 0) ? &G : nullptr, res->blob,
                                    res->cell_off, 0, 0);
  // fold K1 into the candidate stage: replace the span's start event by e0
  ctx->ev_pool.push_back(res->evs[0]);
  res->evs[0] = e0;
  trace_flush(ctx, "run");
  res->n_spans = 1;
  res->host_only = false;
  res->compact_bytes = (long)(st.total_words * 4);
  if (res->want_volumes) rpd_volumes(ctx, res);
  res->synced = false;
}

// Streamed run: the processed tets are cut into n_chunks spans; while span c+1 is searched and clipped
// on the compute stream, the ordered records of span c travel to pinned host memory on a second
// stream (PCIe D2H overlapped with K2/K3).  Spans are contiguous in tet order, so the concatenation
// is the global (tet, site) order and the offsets / ids are those of the one-shot run.
//
// The destination is the library's own growable pinned host buffer (dst_blob == nullptr), or caller memory
// of fixed capacity reachable by cudaMemcpyDefault: pinned / registered host memory (e.g. a shared-memory
// segment every rank of a multi-GPU job writes its shard into), device memory of this GPU, or device
// memory of a PEER GPU opened through CUDA IPC -- then span c crosses NVLink by copy-engine DMA while
// span c+1 is clipped (the multi-GPU gather, fused into the run instead of a collective after it).
void rpd_run_to_host(mb_ctx* ctx, const mb_rpd_opts* opts, int n_chunks, mb_rpd_result* res, void* dst_blob,
                     size_t dst_cap_bytes, long long* dst_off, size_t dst_cap_cells) {
  const bool own = dst_blob == nullptr;
  const int lean = opts ? opts->lean_records : 0;
  MB_REQUIRE(lean >= 0 && lean <= 2, MB_ERR_ARG, "lean_records must be 0 (full), 1 (lean) or 2 (slim)");
  res->lean = lean;
  int t_first, t_count;
  run_prologue(ctx, opts, res, t_first, t_count);
  MB_REQUIRE(!res->want_volumes, MB_ERR_ARG, "want_volumes is not available in the streamed run");
  cudaStream_t s = ctx->stream;
  if (!ctx->copy_stream) {
    MB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; b++) {
      MB_CUDA(cudaEventCreateWithFlags(&ctx->ev_gathered[b], cudaEventDisableTiming));
      MB_CUDA(cudaEventCreateWithFlags(&ctx->ev_copied[b], cudaEventDisableTiming));
    }
  }
  cudaStream_t cs = ctx->copy_stream;
  bool decreasing = false;
  {
    cudaPointerAttributes pa;
    decreasing = dst_blob && cudaPointerGetAttributes(&pa, dst_blob) == cudaSuccess && pa.type == cudaMemoryTypeDevice;
    (void)cudaGetLastError();
  }
  if (n_chunks <= 0) {
    // automatic span count.  Host destinations: the PCIe copy is the long pole, so spans are short (~32k tets) and
    // the first bytes leave early.  Device destinations (this GPU or a peer over NVLink): the copy is ~10x faster
    // than the kernels, so only the last span's copy is exposed and fewer, longer spans waste less on kernel tails
    // and launch gaps (~0.1-0.2 ms per span).
    // Every span costs ~0.1-0.3 ms of kernel tails and launch gaps and only the LAST span's copy is exposed, so a
    // device destination gets two spans of sizes 2 : 1 (measured: three equal spans 2.48 ms, four decreasing ones
    // 2.81 ms per step at N = 2); callers that know how contended the destination's ingress is pass n_chunks
    // (libmat_b200.dist.ShardSink: one span up to two ranks).
    const bool dev_dst = decreasing;
    // (measured at config 2 with the staged run, slim records: 3 / 4 / 6 / 8 spans -> 2.81 / 2.72 / 2.70 / 2.79 ms)
    const int per_span = lean ? 40000 : 32768;
    n_chunks = dev_dst ? (t_count >= 65536 ? 2 : 1) : std::max(1, std::min(32, (t_count + per_span - 1) / per_span));
  }
  n_chunks = std::max(1, std::min(n_chunks, std::max(1, t_count)));
  const bool grid_cands = !ctx->sites.given || (opts && opts->grid_candidates);
  GridDev G;
  cudaEvent_t e0 = take_event(ctx);
  MB_CUDA(cudaEventRecord(e0, s));
  if (grid_cands && t_count > 0) G = grid_build(ctx);
  long long acc_bytes = 0, acc_cells = 0;
  // MB_TRACE=2: device timeline of the spans (kernels done / copy done, ms after the run's first event)
  std::vector<cudaEvent_t> tl;
  std::vector<long long> tl_bytes;
  const bool timeline = ctx->trace_level >= 2;
  const double wall0 = timeline ? now_us() : 0.0;
  // own destination: kept from the previous run; first run sizes it from the first span
  if (own) {
    if (!ctx->pin_off.p) ctx->pin_off.reserve_keep(sizeof(long long) * 1024, 0);
    reinterpret_cast<long long*>(ctx->pin_off.p)[0] = 0;
  } else {
    MB_REQUIRE(dst_off && dst_cap_cells >= 1, MB_ERR_ARG, "sink offsets buffer missing");
    HostScalars* hs = host_scalars(ctx);
    hs->zero = 0;
    MB_CUDA(cudaMemcpyAsync(dst_off, &hs->zero, sizeof(long long), cudaMemcpyDefault, cs));
  }
  // span c = tets [cut(c), cut(c+1)): equal sizes, or weights n, n-1, ..., 1 (cut(c) = t_count * (1 - tri(n-c)/tri(n)))
  // Host destinations with >= 4 spans: the first and the last span are half as long as the others -- the first bytes
  // leave early (what matters when the D2H is the long pole) and the exposed copy of the last span is short (what
  // matters when the kernels are).
  auto cut = [&](int k) -> long long {
    if (!decreasing && n_chunks >= 4) {
      const long long units = 2LL * n_chunks - 2;  // in half-spans: 1 + 2 (n - 2) + 1
      const long long at = k == 0 ? 0 : (k == n_chunks ? units : 2LL * k - 1);
      return (long long)t_count * at / units;
    }
    if (!decreasing) return (long long)t_count * k / n_chunks;
    const long long tri_n = (long long)n_chunks * (n_chunks + 1) / 2, rest = (long long)(n_chunks - k) * (n_chunks - k + 1) / 2;
    return (long long)t_count * (tri_n - rest) / tri_n;
  };
  // staged: the candidate stage once for all tets, K3 + ordering per span (see rpd_candidates_all)
  const bool staged = grid_cands && t_count > 0 && ctx->stream_variant != 1 && !(opts && opts->security_radius);
  const TetSpan sp_all = ctx->mesh.n_sel > 0 ? TetSpan{0, t_count, ctx->mesh.tet_sel.p} : TetSpan{t_first, t_count, nullptr};
  std::vector<long long> cut_pairs;
  if (staged) {
    std::vector<int> cut_tets;
    for (int c = 0; c <= n_chunks; c++) cut_tets.push_back((int)cut(c));
    rpd_candidates_all(ctx, opts, res, sp_all, G, cut_tets, cut_pairs);
    ctx->ev_pool.push_back(res->evs[0]);
    res->evs[0] = e0;
  }
  // hand the ordered records of span c (in span_blob / span_off [c & 1], complete at ev_gathered[c & 1]) to the copy stream
  std::vector<cudaEvent_t> tl_k;  // trace: "kernels done" events, recorded right behind a span's last kernel
  auto deliver = [&](int c, const SpanStats& st) {
    const int b = c & 1;
    const long long bytes = st.total_words * 4;
    // pinned destination large enough for this span (first run: extrapolate from the spans so far)
    const size_t need_blob = (size_t)(acc_bytes + bytes), need_off = sizeof(long long) * (size_t)(acc_cells + st.n_cells + 1);
    if (!own) {
      if (need_blob > dst_cap_bytes || (size_t)(acc_cells + st.n_cells + 1) > dst_cap_cells) {
        MB_CUDA(cudaStreamSynchronize(cs));
        MB_CUDA(cudaStreamSynchronize(s));
        throw MbError{MB_ERR_NOMEM, "sink too small: needs more than " + std::to_string(need_blob) + " bytes / " +
                                        std::to_string(acc_cells + st.n_cells + 1) + " offsets"};
      }
    } else if (need_blob > ctx->pin_blob.cap || need_off > ctx->pin_off.cap) {
      MB_CUDA(cudaStreamSynchronize(cs));  // earlier spans have landed: safe to move them
      const double scale = 1.1 * (double)t_count / (double)std::max<long long>(1, cut(c + 1));
      ctx->pin_blob.reserve_keep(std::max(need_blob, (size_t)(need_blob * scale)), (size_t)acc_bytes);
      ctx->pin_off.reserve_keep(std::max(need_off, (size_t)(need_off * scale)), sizeof(long long) * (size_t)(acc_cells + 1));
    }
    MB_CUDA(cudaStreamWaitEvent(cs, ctx->ev_gathered[b], 0));
    unsigned char* out_blob = own ? (unsigned char*)ctx->pin_blob.p : (unsigned char*)dst_blob;
    long long* out_off = own ? reinterpret_cast<long long*>(ctx->pin_off.p) : dst_off;
    if (bytes > 0)
      MB_CUDA(cudaMemcpyAsync(out_blob + acc_bytes, ctx->span_blob[b].p, (size_t)bytes, cudaMemcpyDefault, cs));
    if (st.n_cells > 0)
      MB_CUDA(cudaMemcpyAsync(out_off + acc_cells, ctx->span_off[b].p, sizeof(long long) * (size_t)(st.n_cells + 1),
                              cudaMemcpyDefault, cs));
    MB_CUDA(cudaEventRecord(ctx->ev_copied[b], cs));
    if (timeline) {
      cudaEvent_t d;
      MB_CUDA(cudaEventCreate(&d));
      MB_CUDA(cudaEventRecord(d, cs));
      tl.push_back(tl_k[c]);
      tl.push_back(d);
      tl_bytes.push_back(bytes);
    }
    acc_bytes += bytes;
    acc_cells += st.n_cells;
  };
  auto mark_gathered = [&](int c) {
    MB_CUDA(cudaEventRecord(ctx->ev_gathered[c & 1], s));
    if (timeline) {
      cudaEvent_t a;
      MB_CUDA(cudaEventCreate(&a));
      MB_CUDA(cudaEventRecord(a, s));
      tl_k.push_back(a);
    }
  };
  // pipelined ranges (staged runs): range c+1 is enqueued before the host waits for range c (see range_enqueue)
  // Opt-in (MB_STREAM_VARIANT=3): with the gather already launched ahead of the host wait there is no round trip left
  // to hide -- measured 2.675 vs 2.681 ms at config 2 on one GPU and 2.89 vs 2.79 ms at config 4 on eight.
  bool pipelined = staged && ctx->stream_variant == 3;
  if (pipelined) {
    const size_t evs_before = res->evs.size();
    // range-accumulated statistics, restored if the run has to be redone
    const long snap[8] = {res->n_cells, res->n_clips, res->n_culled, res->n_exact, res->n_redo, res->n_gc, res->n_flag_pairs,
                          res->n_flag_cells};
    long snap_hist[10];
    for (int i = 0; i < 10; i++) snap_hist[i] = res->hist[i];
    ctx->acc_words.reserve(2);
    dev_zero(ctx, ctx->acc_words.p, 2 * sizeof(unsigned long long));
    std::vector<RangeJob> jobs((size_t)n_chunks);
    bool ok = true;
    SpanStats st;
    for (int c = 0; c < n_chunks && ok; c++) {
      const int b = c & 1;
      if (c >= 2) MB_CUDA(cudaStreamWaitEvent(s, ctx->ev_copied[b], 0));  // span c-2 has left the buffer
      jobs[(size_t)c] = range_enqueue(ctx, opts, res, sp_all, cut_pairs[(size_t)c], cut_pairs[(size_t)c + 1], c,
                                      ctx->span_blob[b], ctx->span_off[b], lean);
      mark_gathered(c);
      if (c >= 1) {
        ok = range_collect(ctx, res, jobs[(size_t)c - 1], st);
        if (ok) deliver(c - 1, st);
      }
    }
    if (ok) {
      ok = range_collect(ctx, res, jobs[(size_t)n_chunks - 1], st);
      if (ok) deliver(n_chunks - 1, st);
    }
    if (!ok) {
      // a range overflowed its scratch bound: drain, forget the partial run, redo it one range at a time (that path
      // grows the scratch and retries by itself)
      MB_CUDA(cudaStreamSynchronize(s));
      MB_CUDA(cudaStreamSynchronize(cs));
      res->n_cells = snap[0];
      res->n_clips = snap[1];
      res->n_culled = snap[2];
      res->n_exact = snap[3];
      res->n_redo = snap[4];
      res->n_gc = snap[5];
      res->n_flag_pairs = snap[6];
      res->n_flag_cells = snap[7];
      for (int i = 0; i < 10; i++) res->hist[i] = snap_hist[i];
      while (res->evs.size() > evs_before) {
        ctx->ev_pool.push_back(res->evs.back());
        res->evs.pop_back();
      }
      for (cudaEvent_t e : tl) cudaEventDestroy(e);
      for (size_t k = tl.size() / 2; k < tl_k.size(); k++) cudaEventDestroy(tl_k[k]);
      tl.clear();
      tl_k.clear();
      tl_bytes.clear();
      acc_bytes = acc_cells = 0;
      pipelined = false;
    }
  }
  for (int c = 0; c < n_chunks && !pipelined; c++) {
    const int c_first = (int)cut(c);
    const int c_count = (int)cut(c + 1) - c_first;
    const TetSpan sp = ctx->mesh.n_sel > 0 ? TetSpan{0, c_count, ctx->mesh.tet_sel.p + c_first}
                                           : TetSpan{t_first + c_first, c_count, nullptr};
    const int b = c & 1;
    if (c >= 2) MB_CUDA(cudaStreamWaitEvent(s, ctx->ev_copied[b], 0));  // span c-2 has left the buffer
    const SpanStats st = staged ? rpd_clip_range(ctx, opts, res, sp_all, cut_pairs[c], cut_pairs[c + 1], c, ctx->span_blob[b],
                                                 ctx->span_off[b], acc_bytes, lean)
                                : rpd_run_span(ctx, opts, res, sp, (grid_cands && c_count > 0) ? &G : nullptr,
                                               ctx->span_blob[b], ctx->span_off[b], acc_bytes, lean);
    if (c == 0 && !staged) {
      ctx->ev_pool.push_back(res->evs[0]);
      res->evs[0] = e0;
    }
    mark_gathered(c);
    deliver(c, st);
  }
  MB_CUDA(cudaStreamSynchronize(cs));
  MB_CUDA(cudaStreamSynchronize(s));
  if (timeline) {
    std::string line = "[libmat_b200 trace] spans (kernels done / copy done ms, MB):";
    for (size_t i = 0; i + 1 < tl.size(); i += 2) {
      float ka = 0.f, kd = 0.f;
      cudaEventElapsedTime(&ka, e0, tl[i]);
      cudaEventElapsedTime(&kd, e0, tl[i + 1]);
      char buf[96];
      snprintf(buf, sizeof buf, " [%.3f / %.3f, %.1f]", ka, kd, tl_bytes[i / 2] / 1e6);
      line += buf;
      cudaEventDestroy(tl[i]);
      cudaEventDestroy(tl[i + 1]);
    }
    fprintf(stderr, "%s  host: start %.0f us (epoch mod 1e7), wall %.0f us\n", line.c_str(), std::fmod(wall0, 1e7), now_us() - wall0);
  }
  trace_flush(ctx, "run_to_host");
  res->n_spans = n_chunks;
  res->host_only = true;
  res->host_blob = nullptr;
  res->host_off = nullptr;
  if (own) {
    res->host_blob = reinterpret_cast<const uint32_t*>(ctx->pin_blob.p);
    res->host_off = reinterpret_cast<const long long*>(ctx->pin_off.p);
    res->generation = ++ctx->stream_generation;  // older handles on the same buffers become stale (MB_ERR_STATE)
  } else {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, dst_blob) == cudaSuccess && pa.type == cudaMemoryTypeHost) {
      res->host_blob = reinterpret_cast<const uint32_t*>(dst_blob);
      res->host_off = dst_off;
    }
    (void)cudaGetLastError();
  }
  res->sink_owned = own;
  res->compact_bytes = (long)acc_bytes;
  res->synced = false;
}

// flagged class: bit 30 of record word 2 (cells) and of pair_words (pairs)
__global__ void k_cell_flags(const uint32_t* __restrict__ blob, const long long* __restrict__ cell_off, long long n,
                             unsigned char* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (blob[cell_off[i] / 4 + 2] & MB_FLAG_BIT) ? 1 : 0;
}
__global__ void k_pair_flags(const int* __restrict__ pair_words, long long n, unsigned char* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = ((unsigned)pair_words[i] & MB_FLAG_BIT) ? 1 : 0;
}

void rpd_fetch_flags(mb_ctx* ctx, mb_rpd_result* res, unsigned char* cell_flag, unsigned char* pair_flag) {
  cudaStream_t s = ctx->stream;
  DevBuf<unsigned char> tmp;
  tmp.reserve((size_t)std::max(res->n_cells, res->n_pairs) + 1);
  if (cell_flag && res->n_cells > 0) {
    ctx->n_launches++;
    k_cell_flags<<<(unsigned)((res->n_cells + 255) / 256), 256, 0, s>>>(res->blob.p, res->cell_off.p, res->n_cells, tmp.p);
    MB_CUDA(cudaGetLastError());
    MB_CUDA(cudaMemcpyAsync(cell_flag, tmp.p, (size_t)res->n_cells, cudaMemcpyDeviceToHost, s));
    MB_CUDA(cudaStreamSynchronize(s));
  }
  if (pair_flag && res->n_pairs > 0) {
    ctx->n_launches++;
    k_pair_flags<<<(unsigned)((res->n_pairs + 255) / 256), 256, 0, s>>>(ctx->pair_words.p, res->n_pairs, tmp.p);
    MB_CUDA(cudaGetLastError());
    MB_CUDA(cudaMemcpyAsync(pair_flag, tmp.p, (size_t)res->n_pairs, cudaMemcpyDeviceToHost, s));
    MB_CUDA(cudaStreamSynchronize(s));
  }
}

void rpd_sync(mb_ctx* ctx, mb_rpd_result* res) {
  MB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (!res->synced && !res->evs.empty()) {
    float acc[3] = {0, 0, 0};
    for (size_t i = 0; i + 3 < res->evs.size(); i += 4)
      for (int k = 0; k < 3; k++) {
        float ms = 0.f;
        MB_CUDA(cudaEventElapsedTime(&ms, res->evs[i + k], res->evs[i + k + 1]));
        acc[k] += ms;
      }
    if (ctx->trace_level >= 2 && res->evs.size() > 4) {
      std::string line = "[libmat_b200 trace] span stages (candidates / clip / order ms, offset of span start):";
      for (size_t i = 0; i + 3 < res->evs.size(); i += 4) {
        float a = 0, b = 0, c = 0, o = 0;
        cudaEventElapsedTime(&a, res->evs[i], res->evs[i + 1]);
        cudaEventElapsedTime(&b, res->evs[i + 1], res->evs[i + 2]);
        cudaEventElapsedTime(&c, res->evs[i + 2], res->evs[i + 3]);
        cudaEventElapsedTime(&o, res->evs[0], res->evs[i]);
        char buf[96];
        snprintf(buf, sizeof buf, " [%.3f / %.3f / %.3f @ %.3f]", a, b, c, o);
        line += buf;
      }
      fprintf(stderr, "%s\n", line.c_str());
    }
    res->ms[0] = acc[0];
    res->ms[1] = acc[1];
    res->ms[2] = acc[2];
    MB_CUDA(cudaEventElapsedTime(&res->ms[3], res->evs.front(), res->evs.back()));
    res->synced = true;
  }
}
