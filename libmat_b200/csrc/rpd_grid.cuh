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

#define GRID_BIG_KCAP 2048  // survivor list of the overflow pass

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

// order-preserving float <-> uint key: warp minima go through one REDUX instruction instead of a 5-step
// shuffle butterfly (inf maps above every finite value; NaN never occurs in these reductions' inputs)
__device__ __forceinline__ unsigned grid_fkey(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float grid_funkey(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ float warp_min(float v) {
  return grid_funkey(__reduce_min_sync(0xffffffffu, grid_fkey(v)));
}

__device__ __forceinline__ float pd_plain(float4 s, float4 p) {
  const float dx = s.x - p.x, dy = s.y - p.y, dz = s.z - p.z;
  return dx * dx + dy * dy + dz * dz - s.w;
}

// domination with rounding slack: site m (vertex power distances f, weight wm) strictly
// dominates site a (e, wa) at all 4 tet vertices => pd_m < pd_a on the whole tet (affine
// difference) => cell(a) does not meet the tet.  pd is evaluated in binary32: |err| <= ~2.4e-7 *
// (d^2 + w) <= 2.4e-7 * (|pd| + 2w); the slack below is ~8x that for both operands.
__device__ __forceinline__ bool dominates(float4 f, float wm, float4 e, float wa) {
  const float base = 2.f * (wa + wm) + 1.f;
  return f.x < e.x - 2e-6f * (fabsf(e.x) + fabsf(f.x) + base) &&
         f.y < e.y - 2e-6f * (fabsf(e.y) + fabsf(f.y) + base) &&
         f.z < e.z - 2e-6f * (fabsf(e.z) + fabsf(f.z) + base) &&
         f.w < e.w - 2e-6f * (fabsf(e.w) + fabsf(f.w) + base);
}

__device__ __forceinline__ float4 pd4(float4 s, float4 p0, float4 p1, float4 p2, float4 p3) {
  return make_float4(pd_plain(s, p0), pd_plain(s, p1), pd_plain(s, p2), pd_plain(s, p3));
}

// argmin over the warp of (val, slot): the slot of the smallest val (ties: the lowest lane holding it)
__device__ __forceinline__ int warp_argmin(float v, int q) {
  const unsigned k = grid_fkey(v);
  const unsigned m = __reduce_min_sync(0xffffffffu, k);
  const unsigned who = __ballot_sync(0xffffffffu, k == m);
  return __shfl_sync(0xffffffffu, q, __ffs(who) - 1);
}

// all-pairs domination on the survivor list + ascending-id rank sort + output (shared by the per-tet and the
// cluster kernel); s_pd / s_w / s_id hold cnt entries of the calling warp
__device__ __forceinline__ void grid_finish_candidates(float4* s_pd, float* s_w, int* s_id, int cnt, int lane, int warp,
                                                       int kcap_out, const unsigned* __restrict__ flags,
                                                       int* __restrict__ cand_pad, int* __restrict__ cand_cnt,
                                                       int* __restrict__ pair_cnt, unsigned long long* __restrict__ counters) {
  __syncwarp();
  // ---- all-pairs domination on the survivors, one lane per ORDERED PAIR (a, m): a list of 8 survivors keeps all 32
  // lanes busy for 2 steps instead of 8 lanes for 8.  A removed entry gets id = INT_MAX (several lanes may write the
  // same value) so that the rank sort pushes it to the tail; its s_pd stays -- a dominated site may still dominate
  // others (domination is transitive), so the outcome does not depend on the order of the tests.
  const int npair = cnt * cnt;
  const unsigned magic = (65536u + (unsigned)cnt - 1u) / (unsigned)max(cnt, 1);  // a = p / cnt by multiply + fix-up (cnt <= 2048)
  for (int p0 = 0; p0 < npair; p0 += 32) {
    const int p = p0 + lane;
    if (p < npair) {
      int a = cnt <= 96 ? (int)(((unsigned)p * magic) >> 16) : p / cnt;
      if (a * cnt > p) a--;
      else if ((a + 1) * cnt <= p) a++;
      const int m = p - a * cnt;
      if (m != a && dominates(s_pd[m], s_w[m], s_pd[a], s_w[a])) s_id[a] = 0x7fffffff;
    }
  }
  __syncwarp();
  int n_keep = 0;
  for (int b = 0; b < cnt; b += 32) {
    const int a2 = b + lane;
    n_keep += __popc(__ballot_sync(0xffffffffu, a2 < cnt && s_id[a2] != 0x7fffffff));
  }
  if (n_keep > kcap_out) {
    if (lane == 0) atomicAdd(&counters[4], 1ull);  // truncated: reported, never silent
  }
  // ---- rank of every kept id among the kept ids (ascending site id), pair-parallel as well; the weights are no
  // longer needed: their slots hold the rank counters
  int* s_rank = reinterpret_cast<int*>(s_w);
  for (int b = lane; b < cnt; b += 32) s_rank[b] = 0;
  __syncwarp();
  for (int p0 = 0; p0 < npair; p0 += 32) {
    const int p = p0 + lane;
    if (p < npair) {
      int a = cnt <= 96 ? (int)(((unsigned)p * magic) >> 16) : p / cnt;
      if (a * cnt > p) a--;
      else if ((a + 1) * cnt <= p) a++;
      const int m = p - a * cnt;
      const int ida = s_id[a];
      if (ida != 0x7fffffff && s_id[m] < ida) atomicAdd(&s_rank[a], 1);
    }
  }
  __syncwarp();
  int n_flag = 0;
  for (int b = 0; b < cnt; b += 32) {
    const int a2 = b + lane;
    bool fl = false;
    if (a2 < cnt) {
      const int id = s_id[a2];
      if (id != 0x7fffffff) {
        const int rank = s_rank[a2];
        if (rank < kcap_out) {
          cand_pad[(size_t)warp * kcap_out + rank] = id;
          fl = flags[id] == 1u;
        }
      }
    }
    n_flag += __popc(__ballot_sync(0xffffffffu, fl));
  }
  if (lane == 0) {
    cand_cnt[warp] = min(n_keep, kcap_out);
    pair_cnt[warp] = n_flag;
  }
  __syncwarp();
}

// One warp per tet, ONE walk over the grid:
//   seed     the 27 fine cells around the centroid give a first U(T) = min_s max_i pd_s(p_i) and the
//            KERNEL SET K: the best site seen for each of the 4 vertices and for U (<= 5 sites).
//   walk     every site with L_s <= U (U keeps shrinking while better sites are met) that is not
//            dominated by a member of K goes to a shared-memory survivor list (a few tens at most).
//            Cells are enumerated directly over the box of radius rho = R_t + sqrt(U + w_max) when
//            that box is small (bounded radii: the common case), through the max-weight pyramid
//            otherwise (heavy-tailed radii).
//   then     all-pairs domination on the survivors + ascending-id rank sort.
// Any weaker streaming filter only lengthens the survivor list: the all-pairs pass decides.  The owner
// of every tet vertex passes both tests, so the final list does not depend on the walk order.
// Output: padded list [t_local][kcap_out] of ALL candidates (the neighbour list of every cell of
// the tet), cand_cnt[t_local], pair_cnt[t_local] = number of flagged candidates (cells to clip).
// A tet whose survivor list exceeds KCAP goes to ovf_list and is redone by the SRC = 1
// instantiation with a 2048-entry list; only a list that still exceeds kcap_out AFTER the
// all-pairs filter is truncated (counted in counters[CNT_CANDOVF], reported by mb_rpd_stats).
// SRC: 0 = the tets of the span, 1 = ovf_list (big-list pass, truncating), 2 = fb_list (tets the cluster kernel
// below handed back; survivor overflow still goes on to ovf_list).
template <int KCAP, int WARPS, int SRC>
__global__ void __launch_bounds__(32 * WARPS, WARPS == 4 ? 6 : 1) k_grid_candidates(
    const float4* __restrict__ vert4, const int4* __restrict__ tet_idx, int tet_first, int tet_count,
    const int* __restrict__ tet_sel, GridDev G, const unsigned* __restrict__ flags, int kcap_out, int* __restrict__ cand_pad,
    int* __restrict__ cand_cnt, int* __restrict__ pair_cnt, unsigned long long* __restrict__ counters,
    int* __restrict__ ovf_list, const int* __restrict__ fb_list) {
  constexpr bool FROM_LIST = SRC == 1;
  extern __shared__ __align__(16) unsigned char grid_smem[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4* s_pd = reinterpret_cast<float4*>(grid_smem) + (size_t)wib * KCAP;
  float* s_w = reinterpret_cast<float*>(reinterpret_cast<float4*>(grid_smem) + (size_t)WARPS * KCAP) + (size_t)wib * KCAP;
  int* s_id = reinterpret_cast<int*>(reinterpret_cast<float*>(reinterpret_cast<float4*>(grid_smem) + (size_t)WARPS * KCAP) +
                                     (size_t)WARPS * KCAP) + (size_t)wib * KCAP;
  int* s_pend = reinterpret_cast<int*>(reinterpret_cast<float*>(reinterpret_cast<float4*>(grid_smem) + (size_t)WARPS * KCAP) +
                                       2 * (size_t)WARPS * KCAP) + (size_t)wib * 64;  // queue of sites awaiting evaluation
  const int R = G.R, R1 = G.R1;
  const float H1 = 4.f * G.h;
  const int n1 = R1 * R1 * R1;
  const int n_work = SRC == 1 ? (int)counters[CNT_OVF_TETS] : SRC == 2 ? (int)counters[CNT_FB_TETS] : tet_count;
  for (int work = blockIdx.x * WARPS + wib; work < n_work; work += gridDim.x * WARPS) {
    const int warp = SRC == 1 ? ovf_list[work] : SRC == 2 ? fb_list[work] : work;  // local tet index
    const int t = tet_sel ? tet_sel[warp] : tet_first + warp;  // global tet id
    const int4 vi = tet_idx[t];
    const float4 p0 = vert4[vi.x], p1 = vert4[vi.y], p2 = vert4[vi.z], p3 = vert4[vi.w];
    const float gx = 0.25f * (p0.x + p1.x + p2.x + p3.x), gy = 0.25f * (p0.y + p1.y + p2.y + p3.y),
                gz = 0.25f * (p0.z + p1.z + p2.z + p3.z);
    const float4 g4 = make_float4(gx, gy, gz, 0.f);
    const float Rt2 = fmaxf(fmaxf(pd_plain(g4, p0), pd_plain(g4, p1)), fmaxf(pd_plain(g4, p2), pd_plain(g4, p3)));
    const float Rt = sqrtf(Rt2) * 1.0001f + 1e-3f;
    const int ci = min(R - 1, max(0, (int)floorf((gx - G.minx) * G.inv_h)));
    const int cj = min(R - 1, max(0, (int)floorf((gy - G.miny) * G.inv_h)));
    const int ck = min(R - 1, max(0, (int)floorf((gz - G.minz) * G.inv_h)));

    // ---- seed: U and the kernel set from the 27 cells around the centroid ------------------------
    float bv0 = INFINITY, bv1 = INFINITY, bv2 = INFINITY, bv3 = INFINITY, bvu = INFINITY;
    int bq0 = -1, bq1 = -1, bq2 = -1, bq3 = -1, bqu = -1;
    auto evalq = [&](int q) {
      const float4 e = pd4(G.site4[q], p0, p1, p2, p3);
      const float m = fmaxf(fmaxf(e.x, e.y), fmaxf(e.z, e.w));
      if (e.x < bv0) { bv0 = e.x; bq0 = q; }
      if (e.y < bv1) { bv1 = e.y; bq1 = q; }
      if (e.z < bv2) { bv2 = e.z; bq2 = q; }
      if (e.w < bv3) { bv3 = e.w; bq3 = q; }
      if (m < bvu) { bvu = m; bqu = q; }
    };
    // rings of growing radius around the centroid's cell until a site is met (tets of the boundary
    // layer can lie a few cells away from the nearest medial sphere)
    float U = INFINITY;
    for (int ar = 1; ar <= 4 && !isfinite(U); ar++) {
      const int sd = 2 * ar + 1, nb = sd * sd * sd;
      for (int b = 0; b < nb; b += 32) {
        const int n = b + lane;
        if (n < nb) {
          const int di = n / (sd * sd) - ar, dj = (n / sd) % sd - ar, dk = n % sd - ar;
          const int i = ci + di, j = cj + dj, k = ck + dk;
          // the inner box was scanned (and found empty) by the previous ring
          if (max(abs(di), max(abs(dj), abs(dk))) == ar || ar == 1)
            if (i >= 0 && j >= 0 && k >= 0 && i < R && j < R && k < R) {
              const int c = (i * R + j) * R + k;
              for (int q = G.cell_off[c]; q < G.cell_off[c + 1]; q++) evalq(q);
            }
        }
      }
      U = warp_min(bvu);
    }
    if (!isfinite(U)) {
      // no site near the centroid (sparse or far-away sites): find U with a pyramid walk first
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
          for (int half = 0; half < 2; half++) {
            const int ch = half * 32 + lane;
            const int i = 4 * i1 + (ch >> 4), j = 4 * j1 + ((ch >> 2) & 3), k = 4 * k1 + (ch & 3);
            if (i < R && j < R && k < R) {
              const int c = (i * R + j) * R + k;
              if (box_dist2(gx, gy, gz, G, i, j, k, G.h) - G.wmax0[c] < U)
                for (int q = G.cell_off[c]; q < G.cell_off[c + 1]; q++) evalq(q);
            }
          }
          U = warp_min(bvu);
        }
      }
    }
    int kq[5];
    kq[0] = warp_argmin(bv0, bq0);
    kq[1] = warp_argmin(bv1, bq1);
    kq[2] = warp_argmin(bv2, bq2);
    kq[3] = warp_argmin(bv3, bq3);
    kq[4] = warp_argmin(bvu, bqu);
    float4 kE[5];
    float kW[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
      if (kq[k] >= 0) {
        const float4 s = G.site4[kq[k]];
        kE[k] = pd4(s, p0, p1, p2, p3);
        kW[k] = s.w;
      } else {
        kE[k] = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
        kW[k] = 0.f;
      }
    }
    // U is a binary32 value: Ue is a certain over-estimate (|err| <= ~2.4e-7 (|U| + 2 w) of any site)
    const float wall = fmaxf(G.wmax_all, 0.f);
    float Ue = U + 4e-6f * (fabsf(U) + 2.f * wall) + 1e-3f;

    int cnt = 0, npend = 0;
    // Sites that pass the cheap L_s <= U test (a few lanes per step) are only QUEUED; the expensive part --
    // the 4 vertex power distances and the domination tests against the kernel set -- runs on the queue with
    // converged lanes, 32 sites at a time.
    auto flush = [&]() {
      __syncwarp();
      while (npend > 0) {
        const int take = min(npend, 32);
        bool ok = false;
        float4 e = make_float4(0, 0, 0, 0);
        float w = 0.f;
        int q = 0;
        if (lane < take) {
          q = s_pend[npend - take + lane];
          const float4 s = G.site4[q];
          e = pd4(s, p0, p1, p2, p3);
          w = s.w;
          bvu = fminf(bvu, fmaxf(fmaxf(e.x, e.y), fmaxf(e.z, e.w)));
          ok = true;
#pragma unroll
          for (int kk = 0; kk < 5; kk++) ok = ok && !dominates(kE[kk], kW[kk], e, w);
        }
        const unsigned mk = __ballot_sync(0xffffffffu, ok);
        if (ok) {
          const int pos = cnt + __popc(mk & ((1u << lane) - 1u));
          if (pos < KCAP) {
            s_id[pos] = G.sorted_id[q];
            s_pd[pos] = e;
            s_w[pos] = w;
          }
        }
        cnt += __popc(mk);
        npend -= take;
      }
      __syncwarp();
    };
    // every lane brings one run [qb, qe) of the cell-sorted site array; the runs are walked FLATTENED, 32 sites
    // per step whatever their lengths (owner of flat index f by binary search over the inclusive scan)
    auto walk = [&](int qb, int qe) {
      const int len = max(qe - qb, 0);
      int incl = len;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      for (int f0 = 0; f0 < total; f0 += 32) {
        const int f = f0 + lane;
        int lo = 0;  // smallest lane whose inclusive count exceeds f
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1) {
          const int v = __shfl_sync(0xffffffffu, incl, lo + step - 1);
          if (v <= f) lo += step;
        }
        lo = min(lo, 31);
        const int o_incl = __shfl_sync(0xffffffffu, incl, lo), o_len = __shfl_sync(0xffffffffu, len, lo);
        const int o_qb = __shfl_sync(0xffffffffu, qb, lo);
        bool pass = false;
        int q = 0;
        if (f < total) {
          q = o_qb + (f - (o_incl - o_len));
          const float4 s = G.site4[q];
          const float dg2 = pd_plain(make_float4(s.x, s.y, s.z, 0.f), g4);
          const float d = fmaxf(0.f, sqrtf(dg2) * 0.9999f - Rt);
          pass = d * d - s.w - 4e-6f * (dg2 + s.w) <= Ue;
        }
        const unsigned mk = __ballot_sync(0xffffffffu, pass);
        if (pass) s_pend[npend + __popc(mk & ((1u << lane) - 1u))] = q;
        npend += __popc(mk);
        if (npend >= 32) flush();
      }
      flush();
    };
    auto cell_range = [&](int i, int j, int k, int& qb, int& qe) {
      qb = qe = 0;
      if (i >= 0 && j >= 0 && k >= 0 && i < R && j < R && k < R) {
        const int c = (i * R + j) * R + k;
        const float d = fmaxf(0.f, sqrtf(box_dist2(gx, gy, gz, G, i, j, k, G.h)) * 0.9999f - Rt);
        if (d * d - G.wmax0[c] <= Ue) {
          qb = G.cell_off[c];
          qe = G.cell_off[c + 1];
        }
      }
    };

    // no candidate lies farther than rho from the centroid (w_s <= w_max); rho in cells:
    const float rho = Rt + sqrtf(fmaxf(0.f, Ue + wall)) * 1.0001f;
    const int a = (int)ceilf(rho * G.inv_h * 1.0001f);
    const int side = 2 * a + 1;
    if (isfinite(U) && side <= 8) {
      // ---- direct walk over the (2a+1)^3 box of fine cells, one lane per ROW of cells along k: the
      // sites are sorted by cell and k is the fastest cell index, so a row is one contiguous run of the
      // sorted site array (a handful of sites per lane instead of a lane per mostly-empty cell).  Rows
      // wholly farther than rho are skipped; every site still takes its own L_s <= U test in walk().
      const int nrows = side * side;
      const int k0 = max(0, ck - a), k1 = min(R - 1, ck + a);
      const float lz = G.minz + k0 * G.h, hz = G.minz + (k1 + 1) * G.h;
      const float dz = fmaxf(0.f, fmaxf(lz - gz, gz - hz));
      for (int b = 0; b < nrows; b += 32) {
        const int n = b + lane;
        int qb = 0, qe = 0;
        if (n < nrows) {
          const int i = ci - a + n / side, j = cj - a + n % side;
          if (i >= 0 && j >= 0 && i < R && j < R) {
            const float lx = G.minx + i * G.h, ly = G.miny + j * G.h;
            const float dx = fmaxf(0.f, fmaxf(lx - gx, gx - (lx + G.h)));
            const float dy = fmaxf(0.f, fmaxf(ly - gy, gy - (ly + G.h)));
            const float d = fmaxf(0.f, sqrtf(dx * dx + dy * dy + dz * dz) * 0.9999f - Rt);
            if (d * d - wall <= Ue) {
              const int c0 = (i * R + j) * R + k0;
              qb = G.cell_off[c0];
              qe = G.cell_off[c0 + (k1 - k0) + 1];
            }
          }
        }
        walk(qb, qe);
        U = fminf(U, warp_min(bvu));
        Ue = U + 4e-6f * (fabsf(U) + 2.f * wall) + 1e-3f;
      }
    } else {
      // ---- walk through the max-weight pyramid, restricted to the coarse nodes that meet the box of
      // radius rho (all of them when rho is not finite) ---------------------------------------------
      int lo1i = 0, lo1j = 0, lo1k = 0, hi1i = R1 - 1, hi1j = R1 - 1, hi1k = R1 - 1;
      if (isfinite(rho) && a < 2 * R) {
        lo1i = max(0, (ci - a) >> 2); hi1i = min(R1 - 1, (ci + a) >> 2);
        lo1j = max(0, (cj - a) >> 2); hi1j = min(R1 - 1, (cj + a) >> 2);
        lo1k = max(0, (ck - a) >> 2); hi1k = min(R1 - 1, (ck + a) >> 2);
      }
      const int s1j = hi1j - lo1j + 1, s1k = hi1k - lo1k + 1;
      const int nb1 = (hi1i - lo1i + 1) * s1j * s1k;
      for (int b1 = 0; b1 < nb1; b1 += 32) {
        const int n = b1 + lane;
        bool keep = false;
        int node = 0;
        if (n < nb1) {
          const int i1 = lo1i + n / (s1j * s1k), j1 = lo1j + (n / s1k) % s1j, k1 = lo1k + n % s1k;
          node = (i1 * R1 + j1) * R1 + k1;
          const float d = fmaxf(0.f, sqrtf(box_dist2(gx, gy, gz, G, i1, j1, k1, H1)) * 0.9999f - Rt);
          keep = d * d - G.wmax1[node] <= Ue;
        }
        unsigned m1 = __ballot_sync(0xffffffffu, keep);
        while (m1) {
          const int n_ = __shfl_sync(0xffffffffu, node, __ffs(m1) - 1);
          m1 &= m1 - 1;
          const int k1 = n_ % R1, j1 = (n_ / R1) % R1, i1 = n_ / (R1 * R1);
          for (int half = 0; half < 2; half++) {
            const int ch = half * 32 + lane;
            int qb, qe;
            cell_range(4 * i1 + (ch >> 4), 4 * j1 + ((ch >> 2) & 3), 4 * k1 + (ch & 3), qb, qe);
            walk(qb, qe);
          }
          U = fminf(U, warp_min(bvu));
          Ue = U + 4e-6f * (fabsf(U) + 2.f * wall) + 1e-3f;
        }
      }
    }
    if (cnt > KCAP) {
      if (!FROM_LIST) {
        // survivors do not fit: hand the tet to the big-list pass
        if (lane == 0) {
          const unsigned long long at = atomicAdd(&counters[CNT_OVF_TETS], 1ull);
          ovf_list[at] = warp;
          cand_cnt[warp] = 0;
          pair_cnt[warp] = 0;
        }
        continue;
      }
      cnt = KCAP;  // 2048 survivors of the streaming filter: give up on the rest (counted below)
      if (lane == 0) atomicAdd(&counters[4], 1ull);
    }
    grid_finish_candidates(s_pd, s_w, s_id, cnt, lane, warp, kcap_out, flags, cand_pad, cand_cnt, pair_cnt, counters);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Cluster variant of the search: neighbouring tets share almost all of their candidates (the 6 tets of a Kuhn cube
// see the same ~40 spheres), so ONE warp walks the grid once for a cluster of GRID_CT consecutive tets and then
// filters the cluster's list per tet -- the walk, the cheap L_s test and the site loads are paid once per cluster.
//
// Cluster bound: g_c, R_c = centre and radius of a ball holding every vertex of the cluster.  For a site s
//   ub_s = (|s - g_c| + R_c)^2 - w_s  >=  max over the ball of pd_s  >=  max_i pd_s(p_i) of every tet T of the cluster
// so U_c = min_s ub_s >= U(T), and every candidate of T (L_s(T) <= U(T), header of this file) satisfies
//   (max(0, |s - g_c| - R_c))^2 - w_s <= U_c.
// Those sites form the cluster list (shared memory).  Per tet: exact 4-vertex power distances of the whole list ->
// exact U(T) and kernel set (the owner of every vertex of T is in the list), the per-tet L_s <= U(T) test, kernel-set
// domination, then the same all-pairs pass as the per-tet kernel.
// A cluster that cannot be handled here (no site within 4 rings, box wider than 10 cells, list longer than
// GRID_KC: incoherent tet order or heavy-tailed radii) hands its tets to fb_list -> k_grid_candidates<.., 2>.
#define GRID_CT 6
#define GRID_KC 128

__device__ __forceinline__ float warp_max(float v) {
  return grid_funkey(__reduce_max_sync(0xffffffffu, grid_fkey(v)));
}

template <int KCAP, int WARPS, int CT, int NB>
__global__ void __launch_bounds__(32 * WARPS, NB) k_grid_candidates_cluster(
    const float4* __restrict__ vert4, const int4* __restrict__ tet_idx, int tet_first, int tet_count, GridDev G,
    const unsigned* __restrict__ flags, int kcap_out, int* __restrict__ cand_pad, int* __restrict__ cand_cnt,
    int* __restrict__ pair_cnt, unsigned long long* __restrict__ counters, int* __restrict__ ovf_list,
    int* __restrict__ fb_list) {
  extern __shared__ __align__(16) unsigned char grid_smem[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr size_t PER_WARP = (size_t)GRID_KC * 36 + (size_t)KCAP * 24;
  unsigned char* base = grid_smem + (size_t)wib * PER_WARP;
  float4* c_site = reinterpret_cast<float4*>(base);       // cluster list: site (x, y, z, w)
  float4* c_pd = c_site + GRID_KC;                         // ... its 4 vertex power distances for the current tet
  float4* s_pd = c_pd + GRID_KC;                           // survivor list of the current tet
  float* s_w = reinterpret_cast<float*>(s_pd + KCAP);
  int* s_id = reinterpret_cast<int*>(s_w + KCAP);
  int* c_id = s_id + KCAP;                                 // cluster list: original site id
  const int R = G.R;
  const float wall = fmaxf(G.wmax_all, 0.f);
  // clusters are aligned to GLOBAL tet ids divisible by CT, whatever the span's first tet: the 6 tets of a Kuhn cube
  // stay together in every span / shard (a cluster straddling two cubes has a ball 1.5x as wide and overflows the list)
  const int shift = tet_first % CT;
  const int n_clusters = (tet_count + shift + CT - 1) / CT;
  for (int cl = blockIdx.x * WARPS + wib; cl < n_clusters; cl += gridDim.x * WARPS) {
    const int t0 = max(0, cl * CT - shift), nt = min((cl + 1) * CT - shift, tet_count) - t0;
    // ---- ball around the cluster's vertices --------------------------------------------------------------------
    const bool has = lane < 4 * nt;
    float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has) {
      const int4 vi = tet_idx[tet_first + t0 + (lane >> 2)];
      const int c = lane & 3;
      pv = vert4[c == 0 ? vi.x : c == 1 ? vi.y : c == 2 ? vi.z : vi.w];
    }
    float sx = has ? pv.x : 0.f, sy = has ? pv.y : 0.f, sz = has ? pv.z : 0.f;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
      sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    const float inv_n = 1.f / (float)(4 * nt);
    const float gx = sx * inv_n, gy = sy * inv_n, gz = sz * inv_n;
    const float4 g4 = make_float4(gx, gy, gz, 0.f);
    const float Rc = sqrtf(warp_max(has ? pd_plain(g4, pv) : 0.f)) * 1.0001f + 1e-3f;
    const int ci = min(R - 1, max(0, (int)floorf((gx - G.minx) * G.inv_h)));
    const int cj = min(R - 1, max(0, (int)floorf((gy - G.miny) * G.inv_h)));
    const int ck = min(R - 1, max(0, (int)floorf((gz - G.minz) * G.inv_h)));
    // ---- seed: U_c from rings of cells around the centre ---------------------------------------------------------
    float bvu = INFINITY;
    int bqu = -1;
    float U = INFINITY;
    for (int ar = 1; ar <= 4 && !isfinite(U); ar++) {
      const int sd = 2 * ar + 1, nb = sd * sd * sd;
      for (int b = 0; b < nb; b += 32) {
        const int n = b + lane;
        if (n < nb) {
          const int di = n / (sd * sd) - ar, dj = (n / sd) % sd - ar, dk = n % sd - ar;
          const int i = ci + di, j = cj + dj, k = ck + dk;
          if (max(abs(di), max(abs(dj), abs(dk))) == ar || ar == 1)
            if (i >= 0 && j >= 0 && k >= 0 && i < R && j < R && k < R) {
              const int c = (i * R + j) * R + k;
              for (int q = G.cell_off[c]; q < G.cell_off[c + 1]; q++) {
                const float4 s = G.site4[q];
                const float r = sqrtf(pd_plain(make_float4(s.x, s.y, s.z, 0.f), g4)) * 1.0001f + Rc;
                const float ub = r * r - s.w;
                if (ub < bvu) {
                  bvu = ub;
                  bqu = q;
                }
              }
            }
        }
      }
      U = warp_min(bvu);
    }
    if (isfinite(U)) {
      // the ball bound of the best seed is loose when the seed sits off-centre: its EXACT maximum over the cluster's
      // vertices (the lanes still hold them) is an upper bound of U(T) of every tet as well, and usually 10-20 % lower
      const float4 sb = G.site4[warp_argmin(bvu, bqu)];
      const float ex = warp_max(has ? pd_plain(sb, pv) : -INFINITY);
      U = fminf(U, ex);
      bvu = fminf(bvu, U);
    }
    float Ue = U + 4e-6f * (fabsf(U) + 2.f * wall) + 1e-3f;
    const float rho = Rc + sqrtf(fmaxf(0.f, Ue + wall)) * 1.0001f;
    const int a = (int)ceilf(rho * G.inv_h * 1.0001f);
    const int side = 2 * a + 1;
    int L = 0;
    bool fallback = !isfinite(U) || side > 10;
    if (!fallback) {
      // ---- one walk over the (2a+1)^2 rows of cells; passing sites go to the cluster list -------------------------
      const int nrows = side * side;
      const int k0 = max(0, ck - a), k1 = min(R - 1, ck + a);
      const float lz = G.minz + k0 * G.h, hz = G.minz + (k1 + 1) * G.h;
      const float dz = fmaxf(0.f, fmaxf(lz - gz, gz - hz));
      for (int b = 0; b < nrows; b += 32) {
        const int n = b + lane;
        int qb = 0, qe = 0;
        if (n < nrows) {
          const int i = ci - a + n / side, j = cj - a + n % side;
          if (i >= 0 && j >= 0 && i < R && j < R) {
            const float lx = G.minx + i * G.h, ly = G.miny + j * G.h;
            const float dx = fmaxf(0.f, fmaxf(lx - gx, gx - (lx + G.h)));
            const float dy = fmaxf(0.f, fmaxf(ly - gy, gy - (ly + G.h)));
            const float d = fmaxf(0.f, sqrtf(dx * dx + dy * dy + dz * dz) * 0.9999f - Rc);
            if (d * d - wall <= Ue) {
              const int c0 = (i * R + j) * R + k0;
              qb = G.cell_off[c0];
              qe = G.cell_off[c0 + (k1 - k0) + 1];
            }
          }
        }
        // the runs of the 32 lanes walked flattened, 32 sites per step
        const int len = max(qe - qb, 0);
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        for (int f0 = 0; f0 < total; f0 += 32) {
          const int f = f0 + lane;
          int lo = 0;
#pragma unroll
          for (int step = 16; step >= 1; step >>= 1) {
            const int v = __shfl_sync(0xffffffffu, incl, lo + step - 1);
            if (v <= f) lo += step;
          }
          lo = min(lo, 31);
          const int o_incl = __shfl_sync(0xffffffffu, incl, lo), o_len = __shfl_sync(0xffffffffu, len, lo);
          const int o_qb = __shfl_sync(0xffffffffu, qb, lo);
          bool pass = false;
          int q = 0;
          float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
          if (f < total) {
            q = o_qb + (f - (o_incl - o_len));
            s = G.site4[q];
            const float dg2 = pd_plain(make_float4(s.x, s.y, s.z, 0.f), g4);
            const float sq = sqrtf(dg2);
            const float r = sq * 1.0001f + Rc;
            bvu = fminf(bvu, r * r - s.w);
            const float d = fmaxf(0.f, sq * 0.9999f - Rc);
            pass = d * d - s.w - 4e-6f * (dg2 + s.w) <= Ue;
          }
          const unsigned mk = __ballot_sync(0xffffffffu, pass);
          if (pass) {
            const int pos = L + __popc(mk & ((1u << lane) - 1u));
            if (pos < GRID_KC) {
              c_site[pos] = s;
              c_id[pos] = G.sorted_id[q];
            }
          }
          L += __popc(mk);
        }
        U = fminf(U, warp_min(bvu));
        Ue = U + 4e-6f * (fabsf(U) + 2.f * wall) + 1e-3f;
      }
      fallback = L > GRID_KC;
    }
    if (fallback) {
      unsigned long long at = 0;
      if (lane == 0) at = atomicAdd(&counters[CNT_FB_TETS], (unsigned long long)nt);
      at = __shfl_sync(0xffffffffu, at, 0);
      if (lane < nt) fb_list[at + lane] = t0 + lane;
      continue;
    }
    __syncwarp();
    // ---- per tet: exact distances of the list, U(T), kernel set, filters, all-pairs ---------------------------------
    for (int ti = 0; ti < nt; ti++) {
      const int warp = t0 + ti;  // local tet index
      const int4 vi = tet_idx[tet_first + warp];
      const float4 p0 = vert4[vi.x], p1 = vert4[vi.y], p2 = vert4[vi.z], p3 = vert4[vi.w];
      const float4 t4 = make_float4(0.25f * (p0.x + p1.x + p2.x + p3.x), 0.25f * (p0.y + p1.y + p2.y + p3.y),
                                    0.25f * (p0.z + p1.z + p2.z + p3.z), 0.f);
      const float Rt2 = fmaxf(fmaxf(pd_plain(t4, p0), pd_plain(t4, p1)), fmaxf(pd_plain(t4, p2), pd_plain(t4, p3)));
      const float Rt = sqrtf(Rt2) * 1.0001f + 1e-3f;
      float bv0 = INFINITY, bv1 = INFINITY, bv2 = INFINITY, bv3 = INFINITY, bvt = INFINITY;
      int bq0 = -1, bq1 = -1, bq2 = -1, bq3 = -1, bqt = -1;
      for (int b = 0; b < L; b += 32) {
        const int i = b + lane;
        if (i < L) {
          const float4 e = pd4(c_site[i], p0, p1, p2, p3);
          c_pd[i] = e;
          const float m = fmaxf(fmaxf(e.x, e.y), fmaxf(e.z, e.w));
          if (e.x < bv0) { bv0 = e.x; bq0 = i; }
          if (e.y < bv1) { bv1 = e.y; bq1 = i; }
          if (e.z < bv2) { bv2 = e.z; bq2 = i; }
          if (e.w < bv3) { bv3 = e.w; bq3 = i; }
          if (m < bvt) { bvt = m; bqt = i; }
        }
      }
      __syncwarp();
      int kq[5];
      kq[0] = warp_argmin(bv0, bq0);
      kq[1] = warp_argmin(bv1, bq1);
      kq[2] = warp_argmin(bv2, bq2);
      kq[3] = warp_argmin(bv3, bq3);
      kq[4] = warp_argmin(bvt, bqt);
      const float Ut = warp_min(bvt);
      const float Uet = Ut + 4e-6f * (fabsf(Ut) + 2.f * wall) + 1e-3f;
      float4 kE[5];
      float kW[5];
#pragma unroll
      for (int k = 0; k < 5; k++) {
        if (kq[k] >= 0) {
          kE[k] = c_pd[kq[k]];
          kW[k] = c_site[kq[k]].w;
        } else {
          kE[k] = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
          kW[k] = 0.f;
        }
      }
      int cnt = 0;
      for (int b = 0; b < L; b += 32) {
        const int i = b + lane;
        bool ok = false;
        float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
        float w = 0.f;
        if (i < L) {
          const float4 s = c_site[i];
          e = c_pd[i];
          w = s.w;
          const float dg2 = pd_plain(make_float4(s.x, s.y, s.z, 0.f), t4);
          const float d = fmaxf(0.f, sqrtf(dg2) * 0.9999f - Rt);
          ok = d * d - w - 4e-6f * (dg2 + w) <= Uet;
#pragma unroll
          for (int kk = 0; kk < 5; kk++) ok = ok && !dominates(kE[kk], kW[kk], e, w);
        }
        const unsigned mk = __ballot_sync(0xffffffffu, ok);
        if (ok) {
          const int pos = cnt + __popc(mk & ((1u << lane) - 1u));
          if (pos < KCAP) {
            s_id[pos] = c_id[i];
            s_pd[pos] = e;
            s_w[pos] = w;
          }
        }
        cnt += __popc(mk);
      }
      if (cnt > KCAP) {  // survivors do not fit: the big-list pass redoes the tet
        if (lane == 0) {
          const unsigned long long at = atomicAdd(&counters[CNT_OVF_TETS], 1ull);
          ovf_list[at] = warp;
          cand_cnt[warp] = 0;
          pair_cnt[warp] = 0;
        }
        __syncwarp();
        continue;
      }
      grid_finish_candidates(s_pd, s_w, s_id, cnt, lane, warp, kcap_out, flags, cand_pad, cand_cnt, pair_cnt, counters);
    }
  }
}

// pairs (flagged candidates only) in (tet, site) order + compact CSR copy of the neighbour lists
__global__ void k_grid_fill(int tet_first, int tet_count, const int* __restrict__ tet_sel, int kcap, const int* __restrict__ cand_pad,
                            const int* __restrict__ cand_cnt, const int* __restrict__ pair_off,
                            const unsigned* __restrict__ flags, int* __restrict__ pair_tet,
                            int* __restrict__ pair_site, int* __restrict__ pair_local, long long cap_pairs) {
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
    const int pos = o + __popc(m & ((1u << lane) - 1u));
    if (f && pos < cap_pairs) {  // cap_pairs: speculative capacity (the host checks the true count afterwards)
      pair_tet[pos] = tet_sel ? tet_sel[warp] : tet_first + warp;
      pair_local[pos] = warp;
      pair_site[pos] = id;
    }
    o += __popc(m);
  }
}
