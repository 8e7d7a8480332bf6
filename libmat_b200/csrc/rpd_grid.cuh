// K1 + K2 (grid-kNN mode): uniform grid over the sites with a max-weight pyramid, and the per-tet
// candidate search.  Replaces, functionally, the reference's CGAL regular triangulation neighbour
// source (src/rpd3d_base/triangulation.cxx:62-144,442-549) + dense power-distance matrix
// (kNN-CUDA/knncuda.cu:334-374) + dense tet-sphere relation (voronoi.cu:154-322).
//
// For a tet T with vertices p_i, centroid g and radius R_t:
//   U(T)  = min over sites m of max_i pd_m(p_i)      (pd_m is convex => its max over T is at a vertex)
//   L_s   = (max(0,|g-c_s|-R_t))^2 - w_s             (<= min over T of pd_s)
// Every site whose power cell meets T satisfies L_s <= U(T): the candidate set C(T) is a superset
// of M(T) = {s : cell(s) meets T}.  A site strictly dominated at all 4 vertices by another
// candidate cannot meet T (pd_m - pd_s is affine) and is removed.  Clipping the cell (T,s)
// against the bisectors of C(T)\{s} is exact: any point of T not in cell(s) belongs to the cell
// of some m in M(T).
#pragma once

#include "mb_internal.h"
#include "rpd_device.cuh"

#define GRID_MAX_K 512

struct GridDev {
  const float4* site4;      // cell-sorted sites (x,y,z,w)
  const int* sorted_id;     // original site id of sorted slot
  const int* cell_off;      // R^3 + 1
  const float* wmax0;       // R^3   max weight per fine cell (-inf if empty)
  const float* wmax1;       // R1^3  max weight per coarse node (4^3 fine cells)
  int R, R1;
  float minx, miny, minz, h, inv_h;
};

__global__ void k_grid_count(const float4* __restrict__ site4, int n_site, GridDev G,
                             int* __restrict__ cnt, int* __restrict__ cell_of) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_site) return;
  const float4 c = site4[s];
  int i = min(G.R - 1, max(0, (int)floorf((c.x - G.minx) * G.inv_h)));
  int j = min(G.R - 1, max(0, (int)floorf((c.y - G.miny) * G.inv_h)));
  int k = min(G.R - 1, max(0, (int)floorf((c.z - G.minz) * G.inv_h)));
  int cell = (i * G.R + j) * G.R + k;
  cell_of[s] = cell;
  atomicAdd(&cnt[cell], 1);
}

// deterministic scatter: rank of a site inside its cell = number of sites with smaller id in the
// same cell (cells hold a handful of sites; one thread per cell does an insertion by id)
__global__ void k_grid_scatter(const float4* __restrict__ site4, int n_site,
                               const int* __restrict__ cell_of, const int* __restrict__ cell_off,
                               int* __restrict__ cursor, int* __restrict__ sorted_id) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_site) return;
  int cell = cell_of[s];
  int pos = atomicAdd(&cursor[cell], 1);
  sorted_id[cell_off[cell] + pos] = s;
}

__global__ void k_grid_finalize(const float4* __restrict__ site4, const int* __restrict__ cell_off,
                                int n_cells, int* __restrict__ sorted_id, float4* __restrict__ out4,
                                float* __restrict__ wmax0) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  const int b = cell_off[c], e = cell_off[c + 1];
  // sort ids inside the cell (ascending) so that the layout is run-to-run deterministic
  for (int i = b + 1; i < e; i++) {
    int x = sorted_id[i], j = i - 1;
    while (j >= b && sorted_id[j] > x) {
      sorted_id[j + 1] = sorted_id[j];
      j--;
    }
    sorted_id[j + 1] = x;
  }
  float wm = -INFINITY;
  for (int i = b; i < e; i++) {
    const float4 s = site4[sorted_id[i]];
    out4[i] = s;
    wm = fmaxf(wm, s.w);
  }
  wmax0[c] = wm;
}

__global__ void k_grid_pyramid(const float* __restrict__ wmax0, int R, int R1, float* __restrict__ wmax1) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= R1 * R1 * R1) return;
  int k1 = n % R1, j1 = (n / R1) % R1, i1 = n / (R1 * R1);
  float wm = -INFINITY;
  for (int a = 0; a < 4; a++)
    for (int b = 0; b < 4; b++)
      for (int c = 0; c < 4; c++) {
        int i = 4 * i1 + a, j = 4 * j1 + b, k = 4 * k1 + c;
        if (i < R && j < R && k < R) wm = fmaxf(wm, wmax0[(i * R + j) * R + k]);
      }
  wmax1[n] = wm;
}

__device__ __forceinline__ float box_dist2(float gx, float gy, float gz, const GridDev& G, int i, int j,
                                           int k, float H) {
  const float lx = G.minx + i * H, ly = G.miny + j * H, lz = G.minz + k * H;
  const float dx = fmaxf(0.f, fmaxf(lx - gx, gx - (lx + H)));
  const float dy = fmaxf(0.f, fmaxf(ly - gy, gy - (ly + H)));
  const float dz = fmaxf(0.f, fmaxf(lz - gz, gz - (lz + H)));
  return dx * dx + dy * dy + dz * dz;
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float pd_plain(float4 s, float4 p) {
  const float dx = s.x - p.x, dy = s.y - p.y, dz = s.z - p.z;
  return dx * dx + dy * dy + dz * dz - s.w;
}

// One warp per tet.  PHASE A finds U(T); PHASE B collects {s : L_s <= U + eps}; then the
// domination filter and an ascending-id rank sort.  Output: padded list [t_local][kcap] of ALL
// candidates (the neighbour list of every cell of this tet), cand_cnt[t_local], and
// pair_cnt[t_local] = number of flagged candidates (cells to clip).
template <int KCAP>
__global__ void __launch_bounds__(128) k_grid_candidates(
    const float4* __restrict__ vert4, const int4* __restrict__ tet_idx, int tet_first, int tet_count,
    GridDev G, const unsigned* __restrict__ flags, int* __restrict__ cand_pad, int* __restrict__ cand_cnt,
    int* __restrict__ pair_cnt, unsigned long long* __restrict__ counters) {
  __shared__ int s_id[4][KCAP];
  __shared__ float s_pd[4][KCAP][4];
  __shared__ float s_w[4][KCAP];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = blockIdx.x * 4 + wib;
  if (warp >= tet_count) return;
  const int t = tet_first + warp;
  const int4 vi = tet_idx[t];
  const float4 p0 = vert4[vi.x], p1 = vert4[vi.y], p2 = vert4[vi.z], p3 = vert4[vi.w];
  const float gx = 0.25f * (p0.x + p1.x + p2.x + p3.x), gy = 0.25f * (p0.y + p1.y + p2.y + p3.y),
              gz = 0.25f * (p0.z + p1.z + p2.z + p3.z);
  const float4 g4 = make_float4(gx, gy, gz, 0.f);
  float Rt2 = fmaxf(fmaxf(pd_plain(g4, p0), pd_plain(g4, p1)), fmaxf(pd_plain(g4, p2), pd_plain(g4, p3)));
  const float Rt = sqrtf(Rt2) * 1.0001f + 1e-3f;
  const int R = G.R, R1 = G.R1;
  const float H1 = 4.f * G.h;

  // ---- phase A: U = min_s max_i pd_s(p_i) ------------------------------------------------------
  float u_lane = INFINITY;
  float U = INFINITY;
  {
    // seed with the fine cell that contains the centroid (and its 26 neighbours)
    const int ci = min(R - 1, max(0, (int)floorf((gx - G.minx) * G.inv_h)));
    const int cj = min(R - 1, max(0, (int)floorf((gy - G.miny) * G.inv_h)));
    const int ck = min(R - 1, max(0, (int)floorf((gz - G.minz) * G.inv_h)));
    if (lane < 27) {
      const int i = ci + lane / 9 - 1, j = cj + (lane / 3) % 3 - 1, k = ck + lane % 3 - 1;
      if (i >= 0 && j >= 0 && k >= 0 && i < R && j < R && k < R) {
        const int c = (i * R + j) * R + k;
        for (int q = G.cell_off[c]; q < G.cell_off[c + 1]; q++) {
          const float4 s = G.site4[q];
          const float m = fmaxf(fmaxf(pd_plain(s, p0), pd_plain(s, p1)), fmaxf(pd_plain(s, p2), pd_plain(s, p3)));
          u_lane = fminf(u_lane, m);
        }
      }
    }
    U = warp_min(u_lane);
  }
  const int n1 = R1 * R1 * R1;
  for (int b1 = 0; b1 < n1; b1 += 32) {
    const int n = b1 + lane;
    bool keep = false;
    if (n < n1) {
      const int k1 = n % R1, j1 = (n / R1) % R1, i1 = n / (R1 * R1);
      keep = box_dist2(gx, gy, gz, G, i1, j1, k1, H1) - G.wmax1[n] < U;
    }
    unsigned m1 = __ballot_sync(0xffffffffu, keep);
    while (m1) {
      const int n_ = b1 + __ffs(m1) - 1;
      m1 &= m1 - 1;
      const int k1 = n_ % R1, j1 = (n_ / R1) % R1, i1 = n_ / (R1 * R1);
#pragma unroll
      for (int half = 0; half < 2; half++) {
        const int ch = half * 32 + lane;
        const int i = 4 * i1 + (ch >> 4), j = 4 * j1 + ((ch >> 2) & 3), k = 4 * k1 + (ch & 3);
        int c = -1;
        if (i < R && j < R && k < R) {
          c = (i * R + j) * R + k;
          if (!(box_dist2(gx, gy, gz, G, i, j, k, G.h) - G.wmax0[c] < U)) c = -1;
        }
        if (c >= 0) {
          for (int q = G.cell_off[c]; q < G.cell_off[c + 1]; q++) {
            const float4 s = G.site4[q];
            const float m = fmaxf(fmaxf(pd_plain(s, p0), pd_plain(s, p1)), fmaxf(pd_plain(s, p2), pd_plain(s, p3)));
            u_lane = fminf(u_lane, m);
          }
        }
      }
      U = warp_min(u_lane);
    }
  }
  // float rounding of pd: |err| <= ~4e-7 (d^2 + w); make U an over-estimate
  const float Ue = U + 1e-5f * (fabsf(U) + 1.f) + 1e-2f;

  // ---- phase B: collect candidates with L_s <= Ue ----------------------------------------------
  int cnt = 0;
  bool overflow = false;
  for (int b1 = 0; b1 < n1; b1 += 32) {
    const int n = b1 + lane;
    bool keep = false;
    if (n < n1) {
      const int k1 = n % R1, j1 = (n / R1) % R1, i1 = n / (R1 * R1);
      const float d = fmaxf(0.f, sqrtf(box_dist2(gx, gy, gz, G, i1, j1, k1, H1)) - Rt);
      keep = d * d - G.wmax1[n] <= Ue;
    }
    unsigned m1 = __ballot_sync(0xffffffffu, keep);
    while (m1) {
      const int n_ = b1 + __ffs(m1) - 1;
      m1 &= m1 - 1;
      const int k1 = n_ % R1, j1 = (n_ / R1) % R1, i1 = n_ / (R1 * R1);
      for (int half = 0; half < 2; half++) {
        const int ch = half * 32 + lane;
        const int i = 4 * i1 + (ch >> 4), j = 4 * j1 + ((ch >> 2) & 3), k = 4 * k1 + (ch & 3);
        int c = -1;
        if (i < R && j < R && k < R) {
          c = (i * R + j) * R + k;
          const float d = fmaxf(0.f, sqrtf(box_dist2(gx, gy, gz, G, i, j, k, G.h)) - Rt);
          if (!(d * d - G.wmax0[c] <= Ue)) c = -1;
        }
        unsigned m0 = __ballot_sync(0xffffffffu, c >= 0);
        while (m0) {
          const int src = __ffs(m0) - 1;
          m0 &= m0 - 1;
          const int cc = __shfl_sync(0xffffffffu, c, src);
          const int qb = G.cell_off[cc], qe = G.cell_off[cc + 1];
          for (int q0 = qb; q0 < qe; q0 += 32) {
            const int q = q0 + lane;
            bool ok = false;
            float4 s = make_float4(0, 0, 0, 0);
            if (q < qe) {
              s = G.site4[q];
              const float dg = sqrtf(pd_plain(make_float4(s.x, s.y, s.z, 0.f), g4));
              const float d = fmaxf(0.f, dg - Rt);
              ok = d * d - s.w <= Ue;
            }
            const unsigned mk = __ballot_sync(0xffffffffu, ok);
            if (ok) {
              const int pos = cnt + __popc(mk & ((1u << lane) - 1u));
              if (pos < KCAP) {
                s_id[wib][pos] = G.sorted_id[q];
                s_pd[wib][pos][0] = pd_plain(s, p0);
                s_pd[wib][pos][1] = pd_plain(s, p1);
                s_pd[wib][pos][2] = pd_plain(s, p2);
                s_pd[wib][pos][3] = pd_plain(s, p3);
                s_w[wib][pos] = s.w;
              }
            }
            cnt += __popc(mk);
          }
        }
      }
    }
  }
  if (cnt > KCAP) {
    overflow = true;
    cnt = KCAP;
  }
  __syncwarp();
  // ---- domination filter + ascending-id rank sort ------------------------------------------------
  int n_keep = 0, n_flag = 0;
  for (int b = 0; b < cnt; b += 32) {
    const int a = b + lane;
    bool keep = false;
    int my_id = 0x7fffffff;
    if (a < cnt) {
      keep = true;
      my_id = s_id[wib][a];
      const float e0 = s_pd[wib][a][0], e1 = s_pd[wib][a][1], e2 = s_pd[wib][a][2], e3 = s_pd[wib][a][3];
      const float wa = s_w[wib][a];
      if (!overflow) {  // with a truncated list the dominating site may be missing: keep all
        for (int m = 0; m < cnt && keep; m++) {
          if (m == a) continue;
          const float wm = s_w[wib][m];
          const float f0 = s_pd[wib][m][0], f1 = s_pd[wib][m][1], f2 = s_pd[wib][m][2], f3 = s_pd[wib][m][3];
          const float base = 2.f * (wa + wm) + 1.f;
          const bool dom = f0 < e0 - 2e-6f * (fabsf(e0) + fabsf(f0) + base) &&
                           f1 < e1 - 2e-6f * (fabsf(e1) + fabsf(f1) + base) &&
                           f2 < e2 - 2e-6f * (fabsf(e2) + fabsf(f2) + base) &&
                           f3 < e3 - 2e-6f * (fabsf(e3) + fabsf(f3) + base);
          if (dom) keep = false;
        }
      }
    }
    // mark removed entries with id = INT_MAX so that the rank sort pushes them to the tail
    if (a < cnt && !keep) s_id[wib][a] = 0x7fffffff;
    const unsigned mk = __ballot_sync(0xffffffffu, keep);
    const unsigned mf = __ballot_sync(0xffffffffu, keep && flags[my_id == 0x7fffffff ? 0 : my_id] == 1u);
    n_keep += __popc(mk);
    n_flag += __popc(mf);
  }
  __syncwarp();
  for (int b = 0; b < cnt; b += 32) {
    const int a = b + lane;
    if (a < cnt) {
      const int id = s_id[wib][a];
      if (id != 0x7fffffff) {
        int rank = 0;
        for (int m = 0; m < cnt; m++) rank += (s_id[wib][m] < id);
        cand_pad[(size_t)warp * KCAP + rank] = id;
      }
    }
  }
  if (lane == 0) {
    cand_cnt[warp] = n_keep;
    pair_cnt[warp] = n_flag;
    if (overflow) atomicAdd(&counters[4], 1ull);
  }
}

// pairs (flagged candidates only) in (tet, site) order + compact CSR copy of the neighbour lists
__global__ void k_grid_fill(int tet_first, int tet_count, int kcap, const int* __restrict__ cand_pad,
                            const int* __restrict__ cand_cnt, const int* __restrict__ pair_off,
                            const unsigned* __restrict__ flags, int* __restrict__ pair_tet,
                            int* __restrict__ pair_site) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= tet_count) return;
  const int n = cand_cnt[warp];
  int o = pair_off[warp];
  for (int b = 0; b < n; b += 32) {
    const int a = b + lane;
    int id = -1;
    bool f = false;
    if (a < n) {
      id = cand_pad[(size_t)warp * kcap + a];
      f = flags[id] == 1u;
    }
    const unsigned m = __ballot_sync(0xffffffffu, f);
    if (f) {
      const int pos = o + __popc(m & ((1u << lane) - 1u));
      pair_tet[pos] = tet_first + warp;
      pair_site[pos] = id;
    }
    o += __popc(m);
  }
}
