// f3: the dist2mat candidate lists built ON THE DEVICE.
//
// Reference (CPU, std::map / std::set per sample, then a replicated upload):
//   gather_point_to_sites            src/matfun_fix/fix_geo_error.cxx:149-178   fid2sites from every power cell's
//                                    cell_to_surfv2fid (K4's surface facets), sample -> sites of its surface face
//   gather_point_to_slab_and_cone    :180-215   per sample, for each site in ascending id: its medial faces (slabs,
//                                    MedialSphere::faces_, ascending face id), its medial edges (cones, edges_,
//                                    ascending edge id), then the sphere itself; nothing is de-duplicated
//   load_and_compute_sample_dist2mat_gpubuffer   :300-366   one private copy of that list per SAMPLE: ~290 bytes of
//                                    int3 per sample cross PCIe (3 GB at config 3)
// All samples of one surface face share one list.  Here the per-FACE lists are built once on the device as a CSR
// (incidence by radix sort, lists by count -> scan -> fill) and a sample contributes 16 bytes (position + face id);
// its (offset, count) simply point at its face's run, which K5 (dist2mat_kernels.cu) already accepts.  List order and
// content are exactly the reference's, so closest_id -- an index into the list -- means the same primitive.
#include <cub/cub.cuh>

#include <vector>

#include "mb_internal.h"

namespace {

__global__ void k_inc_keys(const int* __restrict__ prim, int n, int arity, unsigned long long* __restrict__ keys) {
  // (site, primitive id) incidence keys of the medial faces (arity 3) or edges (arity 2)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * arity) return;
  keys[i] = ((unsigned long long)(unsigned)prim[i] << 32) | (unsigned)(i / arity);
}

__global__ void k_pair_keys(const int* __restrict__ rows2, long n, unsigned long long* __restrict__ keys) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = ((unsigned long long)(unsigned)rows2[2 * i] << 32) | (unsigned)rows2[2 * i + 1];
}

// surface facets of an emitted RPD result -> (fid, site) keys; other facets -> ~0 (sorted to the end)
__global__ void k_facet_keys(const int* __restrict__ f_cell, const int* __restrict__ f_key,
                             const unsigned char* __restrict__ f_istet, long n, int max_surf_fid,
                             const uint32_t* __restrict__ blob, const long long* __restrict__ cell_off,
                             unsigned long long* __restrict__ keys) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long k = ~0ull;
  if (f_istet[i] && f_key[i] <= max_surf_fid) {
    const unsigned site = blob[cell_off[f_cell[i]] / 4 + 1];
    k = ((unsigned long long)(unsigned)f_key[i] << 32) | site;
  }
  keys[i] = k;
}

// CSR row starts of sorted keys by their high word: first[h] = lower bound of h, for h in [0, n_rows]
__global__ void k_row_starts(const unsigned long long* __restrict__ keys, long n, int n_rows, int* __restrict__ first) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h > n_rows) return;
  long lo = 0, hi = n;
  while (lo < hi) {
    const long mid = (lo + hi) >> 1;
    if ((keys[mid] >> 32) < (unsigned long long)h) lo = mid + 1; else hi = mid;
  }
  first[h] = (int)lo;
}

// unique flags of sorted (fid, site) keys (valid keys only)
__global__ void k_unique_flags(const unsigned long long* __restrict__ keys, long n, int* __restrict__ flag) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flag[i] = (keys[i] != ~0ull && (i == 0 || keys[i] != keys[i - 1])) ? 1 : 0;
}
__global__ void k_compact_keys(const unsigned long long* __restrict__ keys, const int* __restrict__ flag,
                               const int* __restrict__ pos, long n, unsigned long long* __restrict__ out) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flag[i]) out[pos[i]] = keys[i];
}

// list length of a surface face: sum over its sites of (#slabs + #cones + 1)
__global__ void k_face_len(const unsigned long long* __restrict__ fs_keys, const int* __restrict__ fs_first, int n_fid,
                           const int* __restrict__ sf_first, const int* __restrict__ se_first, int* __restrict__ len) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f > n_fid) return;
  int L = 0;
  if (f < n_fid)
    for (int q = fs_first[f]; q < fs_first[f + 1]; q++) {
      const int s = (int)(fs_keys[q] & 0xffffffffu);
      L += (sf_first[s + 1] - sf_first[s]) + (se_first[s + 1] - se_first[s]) + 1;
    }
  len[f] = L;
}

// one warp per surface face: [slabs of s | cones of s | sphere s] for its sites s in ascending id
__global__ void k_face_fill(const unsigned long long* __restrict__ fs_keys, const int* __restrict__ fs_first, int n_fid,
                            const unsigned long long* __restrict__ sf_keys, const int* __restrict__ sf_first,
                            const unsigned long long* __restrict__ se_keys, const int* __restrict__ se_first,
                            const int* __restrict__ mm_faces, const int* __restrict__ mm_edges,
                            const long long* __restrict__ list_off, int* __restrict__ prims) {
  const int f = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (f >= n_fid) return;
  long long o = list_off[f];
  for (int q = fs_first[f]; q < fs_first[f + 1]; q++) {
    const int s = (int)(fs_keys[q] & 0xffffffffu);
    const int f0 = sf_first[s], nf = sf_first[s + 1] - f0, e0 = se_first[s], ne = se_first[s + 1] - e0;
    for (int i = lane; i < nf; i += 32) {
      const int id = (int)(sf_keys[f0 + i] & 0xffffffffu);
      int* p = prims + 3 * (o + i);
      p[0] = mm_faces[3 * id];
      p[1] = mm_faces[3 * id + 1];
      p[2] = mm_faces[3 * id + 2];
    }
    o += nf;
    for (int i = lane; i < ne; i += 32) {
      const int id = (int)(se_keys[e0 + i] & 0xffffffffu);
      int* p = prims + 3 * (o + i);
      p[0] = -1;
      p[1] = mm_edges[2 * id];
      p[2] = mm_edges[2 * id + 1];
    }
    o += ne;
    if (lane == 0) {
      int* p = prims + 3 * o;
      p[0] = -1;
      p[1] = -1;
      p[2] = s;
    }
    o += 1;
  }
}

__global__ void k_sample_lists(const int* __restrict__ fid, int n, int n_fid, const long long* __restrict__ list_off,
                               unsigned* __restrict__ offset, unsigned* __restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int f = fid[i];
  if (f < 0 || f >= n_fid) {  // a face no power cell touches: empty list (fix_geo_error.cxx:168-170)
    offset[i] = 0;
    count[i] = 0;
  } else {
    offset[i] = (unsigned)list_off[f];
    count[i] = (unsigned)(list_off[f + 1] - list_off[f]);
  }
}

__global__ void k_closest_prim(const int* __restrict__ closest, const unsigned* __restrict__ offset,
                               const int* __restrict__ prims, int n, int* __restrict__ out3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = closest[i];
  if (c < 0) {
    out3[3 * i] = out3[3 * i + 1] = out3[3 * i + 2] = -1;  // fix_geo_error.cxx:371-375
  } else {
    const int* p = prims + 3 * ((long long)offset[i] + c);
    out3[3 * i] = p[0];
    out3[3 * i + 1] = p[1];
    out3[3 * i + 2] = p[2];
  }
}

inline unsigned nb(long n, int b) { return (unsigned)((n + b - 1) / b); }

void sort_keys(mb_ctx* ctx, DevBuf<unsigned long long>& a, DevBuf<unsigned long long>& b, long n) {
  if (n <= 0) return;
  b.reserve((size_t)n);
  size_t tmp = 0;
  MB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp, a.p, b.p, (int)n, 0, 64, ctx->stream));
  ctx->cub_tmp.reserve(tmp);
  ctx->n_launches += 8;
  MB_CUDA(cub::DeviceRadixSort::SortKeys(ctx->cub_tmp.p, tmp, a.p, b.p, (int)n, 0, 64, ctx->stream));
  std::swap(a.p, b.p);
  std::swap(a.cap, b.cap);
}

// sorted keys -> unique keys (in place), returns the count; invalid (~0) keys are dropped
long unique_keys(mb_ctx* ctx, DevBuf<unsigned long long>& keys, long n) {
  if (n <= 0) return 0;
  cudaStream_t s = ctx->stream;
  DevBuf<int>&flag = ctx->d2m_lists.tmp_flag, &pos = ctx->d2m_lists.tmp_pos;
  DevBuf<unsigned long long>& out = ctx->d2m_lists.tmp_out;
  flag.reserve((size_t)n + 1);
  pos.reserve((size_t)n + 1);
  out.reserve((size_t)n);
  ctx->n_launches += 4;
  k_unique_flags<<<nb(n, 256), 256, 0, s>>>(keys.p, n, flag.p);
  MB_CUDA(cudaMemsetAsync(flag.p + n, 0, sizeof(int), s));
  size_t tmp = 0;
  MB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, flag.p, pos.p, (int)n + 1, s));
  ctx->cub_tmp.reserve(tmp);
  MB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, flag.p, pos.p, (int)n + 1, s));
  k_compact_keys<<<nb(n, 256), 256, 0, s>>>(keys.p, flag.p, pos.p, n, out.p);
  int total = 0;
  MB_CUDA(cudaMemcpyAsync(&total, pos.p + n, sizeof(int), cudaMemcpyDeviceToHost, s));
  MB_CUDA(cudaStreamSynchronize(s));
  std::swap(keys.p, out.p);
  std::swap(keys.cap, out.cap);
  return total;
}

}  // namespace

void d2m_set_medial_mesh(mb_ctx* ctx, const float* spheres, int n_sph, const int* faces, int n_faces, const int* edges,
                         int n_edges) {
  D2MLists& L = ctx->d2m_lists;
  D2MDev& D = ctx->d2m;
  cudaStream_t s = ctx->stream;
  D.spheres.reserve((size_t)n_sph);
  L.mm_faces.reserve(3 * (size_t)n_faces + 1);
  L.mm_edges.reserve(2 * (size_t)n_edges + 1);
  MB_CUDA(cudaMemcpyAsync(D.spheres.p, spheres, sizeof(float4) * (size_t)n_sph, cudaMemcpyHostToDevice, s));
  if (n_faces) MB_CUDA(cudaMemcpyAsync(L.mm_faces.p, faces, sizeof(int) * 3 * (size_t)n_faces, cudaMemcpyHostToDevice, s));
  if (n_edges) MB_CUDA(cudaMemcpyAsync(L.mm_edges.p, edges, sizeof(int) * 2 * (size_t)n_edges, cudaMemcpyHostToDevice, s));
  D.n_sph = n_sph;
  L.n_faces = n_faces;
  L.n_edges = n_edges;
  // sphere -> incident medial faces / edges in ascending id (MedialSphere::faces_ / edges_ are std::set<int>)
  L.sf_keys.reserve(3 * (size_t)n_faces + 1);
  L.se_keys.reserve(2 * (size_t)n_edges + 1);
  L.sf_first.reserve((size_t)n_sph + 2);
  L.se_first.reserve((size_t)n_sph + 2);
  DevBuf<unsigned long long>& tmp = L.tmp_keys;
  ctx->n_launches += 4;
  if (n_faces) k_inc_keys<<<nb(3L * n_faces, 256), 256, 0, s>>>(L.mm_faces.p, n_faces, 3, L.sf_keys.p);
  if (n_edges) k_inc_keys<<<nb(2L * n_edges, 256), 256, 0, s>>>(L.mm_edges.p, n_edges, 2, L.se_keys.p);
  sort_keys(ctx, L.sf_keys, tmp, 3L * n_faces);
  sort_keys(ctx, L.se_keys, tmp, 2L * n_edges);
  // a face / edge listing the same sphere twice would appear twice: std::set semantics need unique keys
  L.n_sf = unique_keys(ctx, L.sf_keys, 3L * n_faces);
  L.n_se = unique_keys(ctx, L.se_keys, 2L * n_edges);
  k_row_starts<<<nb(n_sph + 1, 256), 256, 0, s>>>(L.sf_keys.p, L.n_sf, n_sph, L.sf_first.p);
  k_row_starts<<<nb(n_sph + 1, 256), 256, 0, s>>>(L.se_keys.p, L.n_se, n_sph, L.se_first.p);
  MB_CUDA(cudaGetLastError());
  MB_CUDA(cudaStreamSynchronize(s));
  L.have_mesh = true;
  L.have_lists = false;
}

// keys (fid << 32 | site), unsorted, possibly with duplicates and invalid (~0) entries, already in L.fs_keys
static void build_face_lists(mb_ctx* ctx, long n_keys, int n_fid) {
  D2MLists& L = ctx->d2m_lists;
  D2MDev& D = ctx->d2m;
  cudaStream_t s = ctx->stream;
  MB_REQUIRE(L.have_mesh, MB_ERR_STATE, "mb_dist2mat_set_medial_mesh first");
  DevBuf<unsigned long long>& tmp = L.tmp_keys;
  sort_keys(ctx, L.fs_keys, tmp, n_keys);
  L.n_fs = unique_keys(ctx, L.fs_keys, n_keys);
  L.n_fid = n_fid;
  L.fs_first.reserve((size_t)n_fid + 2);
  DevBuf<int>& len = L.tmp_len;
  len.reserve((size_t)n_fid + 2);
  L.list_off.reserve((size_t)n_fid + 2);
  ctx->n_launches += 4;
  k_row_starts<<<nb(n_fid + 1, 256), 256, 0, s>>>(L.fs_keys.p, L.n_fs, n_fid, L.fs_first.p);
  k_face_len<<<nb(n_fid + 1, 256), 256, 0, s>>>(L.fs_keys.p, L.fs_first.p, n_fid, L.sf_first.p, L.se_first.p, len.p);
  size_t tb = 0;
  MB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, len.p, L.list_off.p, n_fid + 1, s));
  ctx->cub_tmp.reserve(tb);
  MB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tb, len.p, L.list_off.p, n_fid + 1, s));
  long long total = 0;
  MB_CUDA(cudaMemcpyAsync(&total, L.list_off.p + n_fid, sizeof(long long), cudaMemcpyDeviceToHost, s));
  MB_CUDA(cudaStreamSynchronize(s));
  MB_REQUIRE(total < (1ll << 32), MB_ERR_ARG, "per-face primitive lists exceed the 32-bit offsets of the kernel interface");
  D.prims.reserve(3 * (size_t)total + 3);
  D.n_prims = (long)total;
  if (total > 0) {
    k_face_fill<<<nb(32L * n_fid, 256), 256, 0, s>>>(L.fs_keys.p, L.fs_first.p, n_fid, L.sf_keys.p, L.sf_first.p, L.se_keys.p,
                                                    L.se_first.p, L.mm_faces.p, L.mm_edges.p, L.list_off.p, D.prims.p);
    MB_CUDA(cudaGetLastError());
  }
  MB_CUDA(cudaStreamSynchronize(s));
  L.have_lists = true;
}

void d2m_set_face_sites(mb_ctx* ctx, const int* fid_site_rows, long n_rows, int n_fid) {
  D2MLists& L = ctx->d2m_lists;
  cudaStream_t s = ctx->stream;
  L.fs_keys.reserve((size_t)n_rows + 1);
  DevBuf<int>& rows = L.tmp_rows;
  rows.reserve(2 * (size_t)n_rows + 2);
  if (n_rows) {
    MB_CUDA(cudaMemcpyAsync(rows.p, fid_site_rows, sizeof(int) * 2 * (size_t)n_rows, cudaMemcpyHostToDevice, s));
    ctx->n_launches++;
    k_pair_keys<<<nb(n_rows, 256), 256, 0, s>>>(rows.p, n_rows, L.fs_keys.p);
    MB_CUDA(cudaGetLastError());
  }
  build_face_lists(ctx, n_rows, n_fid);
}

void d2m_set_face_sites_from_rpd(mb_ctx* ctx, mb_rpd_result* res, int max_surf_fid) {
  D2MLists& L = ctx->d2m_lists;
  cudaStream_t s = ctx->stream;
  MB_REQUIRE(res->emitted, MB_ERR_STATE, "mb_rpd_emit must be called first");
  const long nf = res->emit_counts.n_facets;
  L.fs_keys.reserve((size_t)nf + 1);
  if (nf) {
    ctx->n_launches++;
    k_facet_keys<<<nb(nf, 256), 256, 0, s>>>(res->f_cell.p, res->f_key.p, res->f_istet.p, nf, max_surf_fid, res->blob.p,
                                            reinterpret_cast<const long long*>(res->cell_off.p), L.fs_keys.p);
    MB_CUDA(cudaGetLastError());
  }
  build_face_lists(ctx, nf, max_surf_fid + 1);
}

void d2m_upload_by_face(mb_ctx* ctx, const float* samples, const int* sample_fid, int n_samples) {
  D2MLists& L = ctx->d2m_lists;
  D2MDev& D = ctx->d2m;
  cudaStream_t s = ctx->stream;
  MB_REQUIRE(L.have_lists, MB_ERR_STATE, "mb_dist2mat_set_face_sites (or _from_rpd) first");
  D.samples.reserve(3 * (size_t)n_samples + 3);
  D.offset.reserve((size_t)n_samples + 1);
  D.count.reserve((size_t)n_samples + 1);
  D.result.reserve((size_t)n_samples + 1);
  D.closest.reserve((size_t)n_samples + 1);
  D.tie.reserve((size_t)n_samples + 1);
  L.sample_fid.reserve((size_t)n_samples + 1);
  if (n_samples) {
    MB_CUDA(cudaMemcpyAsync(D.samples.p, samples, sizeof(float) * 3 * (size_t)n_samples, cudaMemcpyHostToDevice, s));
    MB_CUDA(cudaMemcpyAsync(L.sample_fid.p, sample_fid, sizeof(int) * (size_t)n_samples, cudaMemcpyHostToDevice, s));
    ctx->n_launches++;
    k_sample_lists<<<nb(n_samples, 256), 256, 0, s>>>(L.sample_fid.p, n_samples, L.n_fid, L.list_off.p, D.offset.p, D.count.p);
    MB_CUDA(cudaGetLastError());
  }
  D.n_samples = n_samples;
  MB_CUDA(cudaStreamSynchronize(s));  // header contract: the caller's arrays may be freed on return
}

void d2m_fetch_closest_prims(mb_ctx* ctx, int* prim3) {
  D2MDev& D = ctx->d2m;
  cudaStream_t s = ctx->stream;
  if (D.n_samples <= 0) return;
  DevBuf<int>& out = ctx->d2m_lists.tmp_prim3;
  out.reserve(3 * (size_t)D.n_samples);
  ctx->n_launches++;
  k_closest_prim<<<nb(D.n_samples, 256), 256, 0, s>>>(D.closest.p, D.offset.p, D.prims.p, D.n_samples, out.p);
  MB_CUDA(cudaGetLastError());
  MB_CUDA(cudaMemcpyAsync(prim3, out.p, sizeof(int) * 3 * (size_t)D.n_samples, cudaMemcpyDeviceToHost, s));
  MB_CUDA(cudaStreamSynchronize(s));
}

void d2m_fetch_face_lists(mb_ctx* ctx, long long* list_off, int* prims3) {
  D2MLists& L = ctx->d2m_lists;
  D2MDev& D = ctx->d2m;
  cudaStream_t s = ctx->stream;
  MB_REQUIRE(L.have_lists, MB_ERR_STATE, "no device-built lists");
  if (list_off) MB_CUDA(cudaMemcpyAsync(list_off, L.list_off.p, sizeof(long long) * ((size_t)L.n_fid + 1), cudaMemcpyDeviceToHost, s));
  if (prims3 && D.n_prims > 0) MB_CUDA(cudaMemcpyAsync(prims3, D.prims.p, sizeof(int) * 3 * (size_t)D.n_prims, cudaMemcpyDeviceToHost, s));
  MB_CUDA(cudaStreamSynchronize(s));
}
