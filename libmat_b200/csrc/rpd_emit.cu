// K4 (second half): stream compaction of power-cell facets, vertices and edges from the ordered
// compact records.  Device counterpart of the per-cell part of the reference's CPU post-processing:
//   reload_active          src/rpd3d_base/voronoi_defs.cxx:76-106
//   cal_cell_euler         src/rpd3d_base/voronoi_defs.cxx:190-222
//   compute_vertex_coordinates  voronoi_defs.cxx:33-74
//   get_all_voro_info      src/rpd3d_base/rpd_update.cxx:112-301 (facet / vertex / edge keys)
// Deterministic two-pass count -> scan -> write; items are ordered by (cell, local index).
#include <cub/cub.cuh>

#include "mb_internal.h"
#include "rpd_device.cuh"

namespace {

struct CellView {
  int tet, site, nb_v, nb_p, nb_e;
  const uint32_t* ver;    // nb_v words (uchar4)
  const uint32_t* plane;  // 4*nb_p words
  const uint32_t* meta;   // 3*nb_p words: id2.x, id2.y, h
  const unsigned char* edge;
};

__device__ __forceinline__ CellView view(const uint32_t* w) {
  CellView c;
  c.tet = (int)w[0];
  c.site = (int)w[1];
  c.nb_v = w[2] & 0xff;
  c.nb_p = (w[2] >> 8) & 0xff;
  c.nb_e = (w[2] >> 16) & 0xff;
  c.ver = w + 4;
  c.plane = c.ver + c.nb_v;
  c.meta = c.plane + 4 * c.nb_p;
  c.edge = reinterpret_cast<const unsigned char*>(c.meta + 3 * c.nb_p);
  return c;
}

__device__ __forceinline__ int neigh_of(const CellView& c, int p) {
  // neigh(plane) = (id2.x == voro_id) ? id2.y : id2.x  (rpd_update.cxx:126)
  const int a = (int)c.meta[3 * p], b = (int)c.meta[3 * p + 1];
  return a == c.site ? b : a;
}
__device__ __forceinline__ bool is_bisector(const CellView& c, int p) { return (int)c.meta[3 * p + 1] != -1; }

// active planes as a 64-bit mask (a plane is active iff some vertex references it)
__device__ __forceinline__ unsigned long long active_planes(const CellView& c) {
  unsigned long long m = 0;
  for (int t = 0; t < c.nb_v; t++) {
    const uint32_t v = c.ver[t];
    m |= 1ull << (v & 0xff);
    m |= 1ull << ((v >> 8) & 0xff);
    m |= 1ull << ((v >> 16) & 0xff);
  }
  return m;
}

// an edge is active iff both planes are active and they share >= 2 vertices; returns the two
// smallest shared vertex ids (ascending) in v0, v1
__device__ __forceinline__ bool edge_active(const CellView& c, unsigned long long ap, int e, int& v0, int& v1) {
  const int a = c.edge[3 * e], b = c.edge[3 * e + 1];
  if (!((ap >> a) & 1ull) || !((ap >> b) & 1ull)) return false;
  int n = 0;
  v0 = v1 = -1;
  for (int t = 0; t < c.nb_v; t++) {
    const uint32_t v = c.ver[t];
    const int x = v & 0xff, y = (v >> 8) & 0xff, z = (v >> 16) & 0xff;
    const bool ha = (x == a) | (y == a) | (z == a), hb = (x == b) | (y == b) | (z == b);
    if (ha && hb) {
      if (n == 0) v0 = t;
      if (n == 1) v1 = t;
      n++;
    }
  }
  return n >= 2;
}

// vertex key: sorted neighbour ids of its bisector planes (set semantics), -1 appended when only 2
__device__ __forceinline__ bool vertex_key(const CellView& c, int t, int max_surf_fid, int key[3], int& surf) {
  const uint32_t v = c.ver[t];
  const int pl[3] = {(int)(v & 0xff), (int)((v >> 8) & 0xff), (int)((v >> 16) & 0xff)};
  int n = 0;
  surf = -1;
  int ks[3];
  for (int i = 0; i < 3; i++) {
    if (is_bisector(c, pl[i])) {
      const int nb = neigh_of(c, pl[i]);
      bool dup = false;
      for (int j = 0; j < n; j++) dup |= (ks[j] == nb);
      if (!dup) ks[n++] = nb;
    } else {
      const int fid = (int)c.meta[3 * pl[i]];
      if (fid <= max_surf_fid) surf = fid;  // last one wins (rpd_update.cxx:161-163)
    }
  }
  if (n < 2) return false;
  if (n == 2) ks[n++] = -1;
  // sort 3
  if (ks[0] > ks[1]) { int s = ks[0]; ks[0] = ks[1]; ks[1] = s; }
  if (ks[1] > ks[2]) { int s = ks[1]; ks[1] = ks[2]; ks[2] = s; }
  if (ks[0] > ks[1]) { int s = ks[0]; ks[0] = ks[1]; ks[1] = s; }
  key[0] = ks[0];
  key[1] = ks[1];
  key[2] = ks[2];
  return true;
}

// vertex coordinates, compute_vertex_coordinates voronoi_defs.cxx:33-49 (exact float operation order, perspective divide)
__device__ __forceinline__ void vertex_pos(const CellView& c, int t, float out[3]) {
  const uint32_t v = c.ver[t];
  const float* P = reinterpret_cast<const float*>(c.plane);
  const float* p1 = P + 4 * (v & 0xff);
  const float* p2 = P + 4 * ((v >> 8) & 0xff);
  const float* p3 = P + 4 * ((v >> 16) & 0xff);
  const float rx = -det3_exact(p1[3], p1[1], p1[2], p2[3], p2[1], p2[2], p3[3], p3[1], p3[2]);
  const float ry = -det3_exact(p1[0], p1[3], p1[2], p2[0], p2[3], p2[2], p3[0], p3[3], p3[2]);
  const float rz = -det3_exact(p1[0], p1[1], p1[3], p2[0], p2[1], p2[3], p3[0], p3[1], p3[3]);
  const float rw = det3_exact(p1[0], p1[1], p1[2], p2[0], p2[1], p2[2], p3[0], p3[1], p3[2]);
  out[0] = __fdiv_rn(rx, rw);
  out[1] = __fdiv_rn(ry, rw);
  out[2] = __fdiv_rn(rz, rw);
}

// Centroid of the face loop of an active plane = the pc_face centroid of get_cell_v2surffid (rpd_update.cxx:20-42):
// the loop's vertices in the cyclic order of reload_pc_explicit's walk (voronoi_defs.cxx:141-182 -- start at the
// first vertex referencing the plane, go on with the vertex whose previous plane is this one's next plane), summed
// sequentially in float and divided by the count (cplus3 / cdivide3, common_cxx.h:119-127).
__device__ void face_loop_centroid(const CellView& c, int plane, float out[3]) {
  // entry of vertex t that is `plane` (0..2), or -1
  auto slot = [&](int t) {
    const uint32_t v = c.ver[t];
    return (int)(v & 0xff) == plane ? 0 : ((int)((v >> 8) & 0xff) == plane ? 1 : ((int)((v >> 16) & 0xff) == plane ? 2 : -1));
  };
  int m = 0, first = -1;
  for (int t = 0; t < c.nb_v; t++)
    if (slot(t) >= 0) {
      if (first < 0) first = t;
      m++;
    }
  float sx = 0.f, sy = 0.f, sz = 0.f;
  int i = first, n = 0;
  while (n < m && i >= 0) {
    float p[3];
    vertex_pos(c, i, p);
    sx = __fadd_rn(sx, p[0]);
    sy = __fadd_rn(sy, p[1]);
    sz = __fadd_rn(sz, p[2]);
    n++;
    const int want = (int)((c.ver[i] >> (8 * ((slot(i) + 1) % 3))) & 0xff);
    int nxt = -1;
    for (int j = 0; j < c.nb_v && nxt < 0; j++) {
      const int sj = slot(j);
      if (sj >= 0 && (int)((c.ver[j] >> (8 * ((sj + 2) % 3))) & 0xff) == want) nxt = j;
    }
    i = nxt;
  }
  out[0] = __fdiv_rn(sx, (float)m);
  out[1] = __fdiv_rn(sy, (float)m);
  out[2] = __fdiv_rn(sz, (float)m);
}

// slot of the tet edge between local faces a < b in the per-tet feature-edge table: (0,1)(0,2)(0,3)(1,2)(1,3)(2,3)
__device__ __forceinline__ int fe_slot(int a, int b) { return a == 0 ? b - 1 : (a == 1 ? b + 1 : 5); }

template <bool WRITE>
__global__ void k_emit(const uint32_t* __restrict__ blob, const long long* __restrict__ cell_off, long n_cells,
                       int max_surf_fid, int* __restrict__ cnt_f, int* __restrict__ cnt_v, int* __restrict__ cnt_e,
                       const long long* __restrict__ off_f, const long long* __restrict__ off_v,
                       const long long* __restrict__ off_e, int* __restrict__ f_cell, int* __restrict__ f_key,
                       unsigned char* __restrict__ f_istet, int* __restrict__ v_cell, int* __restrict__ v_lvid,
                       int* __restrict__ v_key3, float* __restrict__ v_pos3, int* __restrict__ v_surf,
                       int* __restrict__ e_cell, int* __restrict__ e_key2, int* __restrict__ e_lvid2,
                       float* __restrict__ c_euler, float* __restrict__ f_centroid3,
                       // feature edges (rpd_update.cxx:209-259): per-tet table of 6 row indices into fe_rows (or -1)
                       const int* __restrict__ fe_table, const int* __restrict__ fe_rows6, int tet_id_base,
                       int* __restrict__ cnt_fe, const long long* __restrict__ off_fe, int* __restrict__ fe_hit6,
                       int* __restrict__ fe_end4, float* __restrict__ fe_end_pos3) {
  const long cell = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= n_cells) return;
  const CellView c = view(blob + cell_off[cell] / 4);
  const unsigned long long ap = active_planes(c);
  int nf = 0, nv = 0, ne = 0, nfe = 0;
  long long of = 0, ov = 0, oe = 0, ofe = 0;
  if (WRITE) {
    of = off_f[cell];
    ov = off_v[cell];
    oe = off_e[cell];
    if (fe_table) ofe = off_fe[cell];
  }
  // facets (rpd_update.cxx:121-140)
  for (int p = 0; p < c.nb_p; p++) {
    if (!((ap >> p) & 1ull)) continue;
    if (WRITE) {
      const bool bis = is_bisector(c, p);
      f_cell[of + nf] = (int)cell;
      f_key[of + nf] = bis ? neigh_of(c, p) : (int)c.meta[3 * p];
      f_istet[of + nf] = bis ? 0 : 1;
      // pc_face centroid of a surface facet (cell_to_surfv2fid, rpd_update.cxx:129-133); zero for the others
      float ctr[3] = {0.f, 0.f, 0.f};
      if (!bis && (int)c.meta[3 * p] <= max_surf_fid) face_loop_centroid(c, p, ctr);
      f_centroid3[3 * (of + nf) + 0] = ctr[0];
      f_centroid3[3 * (of + nf) + 1] = ctr[1];
      f_centroid3[3 * (of + nf) + 2] = ctr[2];
    }
    nf++;
  }
  // vertices (rpd_update.cxx:147-191)
  for (int t = 0; t < c.nb_v; t++) {
    int key[3], surf;
    if (!vertex_key(c, t, max_surf_fid, key, surf)) continue;
    if (WRITE) {
      float pos[3];
      vertex_pos(c, t, pos);
      v_cell[ov + nv] = (int)cell;
      v_lvid[ov + nv] = t;
      v_key3[3 * (ov + nv) + 0] = key[0];
      v_key3[3 * (ov + nv) + 1] = key[1];
      v_key3[3 * (ov + nv) + 2] = key[2];
      v_pos3[3 * (ov + nv) + 0] = pos[0];
      v_pos3[3 * (ov + nv) + 1] = pos[1];
      v_pos3[3 * (ov + nv) + 2] = pos[2];
      v_surf[ov + nv] = surf;
    }
    nv++;
  }
  // edges between two bisectors (rpd_update.cxx:195-200, 261-295) + Euler edge sum
  double sum_e = 0.0;
  for (int e = 0; e < c.nb_e; e++) {
    int v0, v1;
    if (!edge_active(c, ap, e, v0, v1)) continue;
    if (WRITE) sum_e += 1. / (double)c.edge[3 * e + 2];
    const int a = c.edge[3 * e], b = c.edge[3 * e + 1];
    // covered feature edge: both planes are tet faces (local faces 0..3) and the tet edge is in tet_es2fe_map
    if (fe_table && !is_bisector(c, a) && !is_bisector(c, b)) {
      const int lo = min(a, b), hi = max(a, b);
      const int row = fe_table[(size_t)(c.tet - tet_id_base) * 6 + fe_slot(lo, hi)];
      if (row >= 0) {
        if (WRITE) {
          // (cell, kind, lv1 < lv2, fe_line_id, fe_id) -- se_covered_lvids / ce_covered_lvids -- and per end vertex
          // its FIRST half-plane's neighbour + position (se_line_endpos; neigh = -1: the vertex lies on no half-plane)
          int* h = fe_hit6 + 6 * (ofe + nfe);
          h[0] = (int)cell;
          h[1] = fe_rows6[6 * row + 3];
          h[2] = v0;
          h[3] = v1;
          h[4] = fe_rows6[6 * row + 5];
          h[5] = fe_rows6[6 * row + 4];
          for (int k = 0; k < 2; k++) {
            const int lv = k == 0 ? v0 : v1;
            const uint32_t v = c.ver[lv];
            int neigh = -1;
            for (int i = 0; i < 3 && neigh == -1; i++) {
              const int pl = (int)((v >> (8 * i)) & 0xff);
              if (is_bisector(c, pl)) neigh = neigh_of(c, pl);
            }
            int* en = fe_end4 + 4 * (2 * (ofe + nfe) + k);
            en[0] = (int)cell;
            en[1] = lv;
            en[2] = neigh;
            en[3] = fe_rows6[6 * row + 5];
            float pos[3] = {0.f, 0.f, 0.f};
            if (neigh != -1) vertex_pos(c, lv, pos);
            fe_end_pos3[3 * (2 * (ofe + nfe) + k) + 0] = pos[0];
            fe_end_pos3[3 * (2 * (ofe + nfe) + k) + 1] = pos[1];
            fe_end_pos3[3 * (2 * (ofe + nfe) + k) + 2] = pos[2];
          }
        }
        nfe++;
      }
    }
    if (!is_bisector(c, a) || !is_bisector(c, b)) continue;
    if (WRITE) {
      int k0 = neigh_of(c, a), k1 = neigh_of(c, b);
      if (k0 > k1) { int s = k0; k0 = k1; k1 = s; }
      e_cell[oe + ne] = (int)cell;
      e_key2[2 * (oe + ne) + 0] = k0;
      e_key2[2 * (oe + ne) + 1] = k1;
      e_lvid2[2 * (oe + ne) + 0] = v0;
      e_lvid2[2 * (oe + ne) + 1] = v1;
    }
    ne++;
  }
  if (WRITE) {
    // cal_cell_euler (voronoi_defs.cxx:190-222): Scalar = double accumulators, float result
    double sum_v = 0.0, sum_f = 0.0;
    for (int t = 0; t < c.nb_v; t++) sum_v += (double)__fdiv_rn(1.f, (float)(int)(c.ver[t] >> 24));
    for (int p = 0; p < c.nb_p; p++)
      if ((ap >> p) & 1ull) sum_f += 1. / (double)__uint_as_float(c.meta[3 * p + 2]);
    c_euler[cell] = (float)(sum_v - sum_e + sum_f);
  } else {
    cnt_f[cell] = nf;
    cnt_v[cell] = nv;
    cnt_e[cell] = ne;
    if (fe_table) cnt_fe[cell] = nfe;
  }
}

// ---- a12: per-cell volume and barycentre sums (atomic_add_bary_and_volume, convex_cell.cu:1008-1069,
// with get_tet_decomposition_of_vertex :986-1006): every vertex contributes 6 tets built from the
// projections of the seed on the vertex's three planes; the literal float expressions (this TU is
// compiled with -fmad=false).  Per-cell values are deterministic; the per-site sums are accumulated
// with float atomics like the reference (:1045-1054), so their last bits depend on the order.
struct V4 {
  float x, y, z, w;
};
__device__ __forceinline__ V4 v_minus(V4 A, V4 B) { return {A.x - B.x, A.y - B.y, A.z - B.z, A.w - B.w}; }
__device__ __forceinline__ V4 v_plus(V4 A, V4 B) { return {A.x + B.x, A.y + B.y, A.z + B.z, A.w + B.w}; }
__device__ __forceinline__ V4 v_mul3(float s, V4 A) { return {s * A.x, s * A.y, s * A.z, 1.f}; }
__device__ __forceinline__ float v_dot4(V4 A, V4 B) { return A.x * B.x + A.y * B.y + A.z * B.z + A.w * B.w; }
__device__ __forceinline__ float v_dot3(V4 A, V4 B) { return A.x * B.x + A.y * B.y + A.z * B.z; }
__device__ __forceinline__ V4 v_cross3(V4 A, V4 B) {
  return {A.y * B.z - A.z * B.y, A.z * B.x - A.x * B.z, A.x * B.y - A.y * B.x, 0.f};
}
// common_cuda.h:235-240
__device__ __forceinline__ V4 project_on_plane(V4 P, V4 plane) {
  const V4 n = {plane.x, plane.y, plane.z, 0.f};
  const float n_2 = v_dot4(n, n);
  const float lambda = (double)n_2 > 1e-2 ? (v_dot4(n, P) + plane.w) / n_2 : 0.0f;
  return v_plus(P, v_mul3(-lambda, n));
}

__global__ void k_cell_volumes(const uint32_t* __restrict__ blob, const long long* __restrict__ cell_off,
                               long n_cells, const float4* __restrict__ site4, int n_site,
                               float* __restrict__ cell_vol, float* __restrict__ site_vol,
                               float* __restrict__ site_bary) {
  const long cell = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= n_cells) return;
  const CellView c = view(blob + cell_off[cell] / 4);
  const float4 sd = site4[c.site];
  const V4 C = {sd.x, sd.y, sd.z, sd.w};
  const float* PL = reinterpret_cast<const float*>(c.plane);
  V4 bary_sum = {0.f, 0.f, 0.f, 0.f};
  float cur = 0.f;
  for (int t = 0; t < c.nb_v; t++) {
    const uint32_t v = c.ver[t];
    const int pi[3] = {(int)(v & 0xff), (int)((v >> 8) & 0xff), (int)((v >> 16) & 0xff)};
    const float* p1 = PL + 4 * pi[0];
    const float* p2 = PL + 4 * pi[1];
    const float* p3 = PL + 4 * pi[2];
    // compute_vertex_coordinates with perspective divide (convex_cell.cu:319-351)
    const float rx = -det3_exact(p1[3], p1[1], p1[2], p2[3], p2[1], p2[2], p3[3], p3[1], p3[2]);
    const float ry = -det3_exact(p1[0], p1[3], p1[2], p2[0], p2[3], p2[2], p3[0], p3[3], p3[2]);
    const float rz = -det3_exact(p1[0], p1[1], p1[3], p2[0], p2[1], p2[3], p3[0], p3[1], p3[3]);
    const float rw = det3_exact(p1[0], p1[1], p1[2], p2[0], p2[1], p2[2], p3[0], p3[1], p3[2]);
    const V4 A = {rx / rw, ry / rw, rz / rw, 1.f};
    V4 P[6];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const float* q = PL + 4 * pi[i];
      P[2 * i] = project_on_plane(C, V4{q[0], q[1], q[2], q[3]});
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const V4 n = v_cross3(v_minus(P[2 * i], C), v_minus(P[(2 * (i + 1)) % 6], C));
      const V4 pl = {n.x, n.y, n.z, -v_dot3(C, n)};
      P[2 * i + 1] = project_on_plane(A, pl);
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
      const V4 a = v_minus(P[i], A), b = v_minus(P[(i + 1) % 6], A), cc = v_minus(C, A);
      const float tv = (float)((double)(-det3_exact(a.x, a.y, a.z, b.x, b.y, b.z, cc.x, cc.y, cc.z)) / 6.);
      const V4 q0 = P[i], q1 = P[(i + 1) % 6];
      const V4 tb = {.25f * (q0.x + q1.x + C.x + A.x), .25f * (q0.y + q1.y + C.y + A.y),
                     .25f * (q0.z + q1.z + C.z + A.z), 1.0f};
      bary_sum = v_plus(bary_sum, v_mul3(tv, tb));
      cur += tv;
    }
  }
  cell_vol[cell] = cur;
  if ((double)fabsf(cur) < 0.1) return;  // the reference flips the status and adds nothing (:1040)
  atomicAdd(&site_bary[c.site], bary_sum.x);
  atomicAdd(&site_bary[c.site + n_site], bary_sum.y);
  atomicAdd(&site_bary[c.site + 2 * (size_t)n_site], bary_sum.z);
  atomicAdd(&site_vol[c.site], cur);
}

void scan_ll(mb_ctx* ctx, const int* in, long long* out, long n) {
  size_t tmp = 0;
  MB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, n, ctx->stream));
  ctx->cub_tmp.reserve(tmp);
  ctx->n_launches += 2;  // DeviceScanInitKernel + DeviceScanKernel
  MB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, in, out, n, ctx->stream));
}

}  // namespace

void rpd_volumes(mb_ctx* ctx, mb_rpd_result* res) {
  cudaStream_t s = ctx->stream;
  const long n = res->n_cells;
  res->site_vol.reserve((size_t)res->n_site + 1);
  res->site_bary.reserve(3 * (size_t)res->n_site + 1);
  res->cell_vol.reserve((size_t)n + 1);
  MB_CUDA(cudaMemsetAsync(res->site_vol.p, 0, sizeof(float) * (size_t)res->n_site, s));
  MB_CUDA(cudaMemsetAsync(res->site_bary.p, 0, sizeof(float) * 3 * (size_t)res->n_site, s));
  if (n == 0) return;
  ctx->n_launches++;
  k_cell_volumes<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(res->blob.p, res->cell_off.p, n, ctx->sites.site4.p,
                                                           res->n_site, res->cell_vol.p, res->site_vol.p,
                                                           res->site_bary.p);
  MB_CUDA(cudaGetLastError());
}

void rpd_emit(mb_ctx* ctx, mb_rpd_result* res, int max_surf_fid) {
  cudaStream_t s = ctx->stream;
  const long n = res->n_cells;
  res->emit_counts = {0, 0, 0};
  res->emitted = n == 0;  // set only once the work below has succeeded
  if (n == 0) return;
  DevBuf<int> cf, cv, ce, cfe;
  DevBuf<long long> of, ov, oe, ofe;
  cf.reserve(n + 1); cv.reserve(n + 1); ce.reserve(n + 1);
  of.reserve(n + 1); ov.reserve(n + 1); oe.reserve(n + 1);
  MB_CUDA(cudaMemsetAsync(cf.p + n, 0, sizeof(int), s));
  MB_CUDA(cudaMemsetAsync(cv.p + n, 0, sizeof(int), s));
  MB_CUDA(cudaMemsetAsync(ce.p + n, 0, sizeof(int), s));
  const TetMeshDev& M = ctx->mesh;
  const int* fe_table = M.n_fe > 0 ? M.fe_table.p : nullptr;
  if (fe_table) {
    cfe.reserve(n + 1);
    ofe.reserve(n + 1);
    MB_CUDA(cudaMemsetAsync(cfe.p + n, 0, sizeof(int), s));
  }
  const unsigned blocks = (unsigned)((n + 127) / 128);
  ctx->n_launches++;
  k_emit<false><<<blocks, 128, 0, s>>>(res->blob.p, res->cell_off.p, n, max_surf_fid, cf.p, cv.p, ce.p, nullptr,
                                       nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                       nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, fe_table,
                                       M.fe_rows.p, M.tet_id_base, cfe.p, nullptr, nullptr, nullptr, nullptr);
  MB_CUDA(cudaGetLastError());
  scan_ll(ctx, cf.p, of.p, n + 1);
  scan_ll(ctx, cv.p, ov.p, n + 1);
  scan_ll(ctx, ce.p, oe.p, n + 1);
  if (fe_table) scan_ll(ctx, cfe.p, ofe.p, n + 1);
  long long tot[4] = {0, 0, 0, 0};
  MB_CUDA(cudaMemcpyAsync(&tot[0], of.p + n, sizeof(long long), cudaMemcpyDeviceToHost, s));
  MB_CUDA(cudaMemcpyAsync(&tot[1], ov.p + n, sizeof(long long), cudaMemcpyDeviceToHost, s));
  MB_CUDA(cudaMemcpyAsync(&tot[2], oe.p + n, sizeof(long long), cudaMemcpyDeviceToHost, s));
  if (fe_table) MB_CUDA(cudaMemcpyAsync(&tot[3], ofe.p + n, sizeof(long long), cudaMemcpyDeviceToHost, s));
  MB_CUDA(cudaStreamSynchronize(s));
  res->n_fe_hits = (long)tot[3];
  res->f_centroid3.reserve(3 * tot[0] + 1);
  res->fe_hit6.reserve(6 * tot[3] + 1);
  res->fe_end4.reserve(8 * tot[3] + 1);
  res->fe_end_pos3.reserve(6 * tot[3] + 1);
  res->emit_counts.n_facets = (long)tot[0];
  res->emit_counts.n_vertices = (long)tot[1];
  res->emit_counts.n_edges = (long)tot[2];
  res->f_cell.reserve(tot[0] + 1); res->f_key.reserve(tot[0] + 1); res->f_istet.reserve(tot[0] + 1);
  res->v_cell.reserve(tot[1] + 1); res->v_lvid.reserve(tot[1] + 1); res->v_key3.reserve(3 * tot[1] + 1);
  res->v_pos3.reserve(3 * tot[1] + 1); res->v_surf.reserve(tot[1] + 1);
  res->e_cell.reserve(tot[2] + 1); res->e_key2.reserve(2 * tot[2] + 1); res->e_lvid2.reserve(2 * tot[2] + 1);
  res->c_euler.reserve(n + 1);
  ctx->n_launches++;
  k_emit<true><<<blocks, 128, 0, s>>>(res->blob.p, res->cell_off.p, n, max_surf_fid, nullptr, nullptr, nullptr,
                                      of.p, ov.p, oe.p, res->f_cell.p, res->f_key.p, res->f_istet.p,
                                      res->v_cell.p, res->v_lvid.p, res->v_key3.p, res->v_pos3.p, res->v_surf.p,
                                      res->e_cell.p, res->e_key2.p, res->e_lvid2.p, res->c_euler.p, res->f_centroid3.p,
                                      fe_table, M.fe_rows.p, M.tet_id_base, nullptr, ofe.p, res->fe_hit6.p,
                                      res->fe_end4.p, res->fe_end_pos3.p);
  MB_CUDA(cudaGetLastError());
  MB_CUDA(cudaStreamSynchronize(s));
  res->emitted = true;
}
