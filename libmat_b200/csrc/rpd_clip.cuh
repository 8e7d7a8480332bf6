// K3: sub-warp-cooperative convex-cell clipping (one group of G lanes per (tet, site) cell).
//
// Replaces clipped_voro_cell_test_GPU_param_tet (reference src/rpd3d/convex_cell.cu:1166-1337),
// which runs ONE THREAD per cell with 3 KB of shared memory per thread (64 threads / SM) and
// evaluates a 4x4 FP64 determinant for every (vertex, plane) pair.
//
// Here G in {4,8,16,32} lanes share one polytope in shared memory and the 32/G groups of a warp
// advance in LOCKSTEP through a small state machine (fetch pair -> scan G neighbours -> clip by one
// plane -> ... -> write record), so that the order-defining serial bookkeeping of the groups
// co-issues instead of serialising the warp:
//   * scan: lanes gather G neighbour spheres (float4), build the exact bisector, and cull planes
//     that provably remove nothing with a conservative FP32 test against the 4 tet vertices (the
//     reference would clip and pop them again, convex_cell.cu:741-744);
//   * clip: lanes evaluate the conflict predicate of the cell vertices in parallel.  FILTERED
//     PREDICATE: every vertex caches the cofactor vector of its three planes (the FP64 minors of
//     det4x4, rounded to FP32) when it is created; sign(cofactors . plane) is accepted when it
//     exceeds a rigorous error bound, otherwise the reference's FP64 det4x4 is evaluated in its
//     literal operation order (common_cuda.h:195-213) -- decisions are bit-identical to the
//     reference's in both cases;
//   * one lane per group replays the reference's order-defining bookkeeping (swap partition
//     :706-721, cavity boundary walk :618-678) on the shared-memory state, then lanes create the
//     new edges / vertices (+ their cofactors) in parallel (:756-773).
// Array positions therefore equal the reference's, so records are byte-identical on the defined
// entries in given-neighbours mode.  Work is distributed dynamically (warp-level chunks from a
// global cursor) because the cost per cell varies by an order of magnitude.
#pragma once

#include "rpd_device.cuh"

struct ClipArgs {
  // mesh
  const float4* vert4;
  const int4* tet_idx;
  const int4* tet_fadj;
  const int4* tet_fid;
  const uint2* tet_e6;
  const float4* tet_geo;  // per tet: 4 face planes, 4 initial-vertex cofactor vectors (k_tet_geometry)
  const unsigned* tet_vadj;  // per tet: (uchar) v_adjs of its 4 vertices, packed
  // sites
  const float4* site4;
  int n_site;
  // neighbour lists
  const int* nbr;        // per-site lists: [n_site][nbr_stride]; per-tet lists: [local tet][nbr_stride]
  int nbr_stride;        // per-site: site_k; per-tet: kcap
  const int* nbr_cnt;    // per-tet lists: #candidates per local tet (nullptr for per-site lists)
  int tet_first;         // first tet of the processed range (per-tet lists are relative to it)
  int tet_id_base;       // added to the tet id stored in the records (mb_set_tet_id_base: sharded uploads)
  // pairs
  const int* pair_tet;
  const int* pair_site;
  const int* pair_local;  // local index of the pair's tet in the processed range / subset (per-tet lists)
  long long n_pairs;     // number of pairs; an upper bound (the arrays' capacity) when n_pairs_dev is set
  const int* n_pairs_dev;  // device-resident pair count (speculative launch: the host has not read it yet)
  int grab;              // pairs a warp takes from the global cursor at a time (multiple of 32/G); 0 = derive
  // outputs
  signed char* pair_status;
  long long* pair_blob;
  int* pair_words;
  uint32_t* scratch;
  unsigned long long scratch_words;
  unsigned long long* counters;  // RpdCounters as u64 array
  // compact-caps first pass -> full-caps second pass (grid-kNN mode)
  int* redo_out;                          // first pass: pairs whose cell outgrew the compact caps
  const int* work_list;                   // second pass: work item -> pair (nullptr: identity)
  const unsigned long long* work_count;   // second pass: number of work items (device-resident)
  int no_cull;                            // debug (MB_NO_CULL=1): clip by every listed neighbour like the reference does
  int security_radius;                    // given-neighbours mode: the reference's security-radius exit (a9), opt-in
};

// indices into RpdCounters viewed as u64[]
#define CNT_BLOB 0
#define CNT_CLIPS 1
#define CNT_CULLED 2
#define CNT_VALID 3
#define CNT_CANDOVF 4
#define CNT_HIST 5
#define CNT_EXACT 15  // conflict tests that fell through the FP32 filter to the FP64 determinant
#define CNT_WORK_CURSOR_IDX 17
#define CNT_REDO 19               // cells handed from the compact-caps pass to the full-caps pass
#define CNT_WORK_CURSOR2_IDX 20   // work cursor of the second pass
#define CNT_FLAG_PAIRS 21         // pairs with a conflict test under the static-filter bound (flagged class)
#define CNT_FLAG_CELLS 22         // ... of which valid cells
#define MB_FLAG_BIT 0x40000000u   // record word 2 / pair_words bit 30: flagged cell

template <int G>
__device__ __forceinline__ unsigned group_ballot(unsigned gmask, int gshift, bool pred) {
  unsigned m = __ballot_sync(gmask, pred);
  return (m >> gshift) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
}

// z of the edge between planes a < b (closed form of what new_edge stored, convex_cell.cu:604-616,
// and of the 6 initial tet edges :194-210)
__device__ __forceinline__ unsigned char edge_z(int a, int b, unsigned hf4, unsigned long long e6) {
  if (b < 4) {
    // face pair -> index in e_adj6 order: (2,3)->0 (1,3)->1 (1,2)->2 (0,3)->3 (0,2)->4 (0,1)->5
    int idx = (a == 2) ? 0 : (a == 1 ? (b == 3 ? 1 : 2) : (b == 3 ? 3 : (b == 2 ? 4 : 5)));
    return (unsigned char)((e6 >> (8 * idx)) & 0xff);
  }
  // max(h_a, h_b) with h = 1 for bisectors; h_a = (uchar) f_adj of tet face a
  const unsigned ha = a < 4 ? ((hf4 >> (8 * a)) & 0xffu) : 1u;
  return (unsigned char)max(ha, 1u);
}

#define CLIP_CHUNK_WORDS 1024u  // scratch is bump-allocated per group in 4 KB chunks

enum : int { GS_IDLE = 0, GS_NEW = 1, GS_RUN = 2, GS_FINISH = 3, GS_EXIT = 4 };

// cofactor vector of a vertex (p1,p2,p3): det4x4(p1,p2,p3,e) = c . e with
// c = (m234, -m134, m124, -m123)
__device__ __forceinline__ float4 cofactors_f32(const Minors& m) {
  return make_float4((float)m.m234, (float)(-m.m134), (float)m.m124, (float)(-m.m123));
}

// Garbage collection of dead planes / inactive edges (per-tet candidate lists only; see B2 in k_clip).  Rare path, kept
// out of line: all lanes of ONE group call it together.
template <int G, class CellS>
__device__ __noinline__ void cell_gc(CellS& S, int lane, unsigned gmask, int gshift, int nb_v, int& nb_p, int& nb_e) {
  unsigned long long am = 0xFull;
  for (int v = lane; v < nb_v; v += G) {
    const uchar4 tv = S.ver[v];
    am |= (1ull << tv.x) | (1ull << tv.y) | (1ull << tv.z);
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) am |= __shfl_xor_sync(gmask, am, o);
  int ne_new = 0;
  for (int eb = 0; eb < nb_e; eb += G) {
    const int ei = eb + lane;
    bool keep = false;
    unsigned char a = 0, b = 0, z = 0;
    if (ei < nb_e) {
      a = S.edge[3 * ei];
      b = S.edge[3 * ei + 1];
      z = S.edge[3 * ei + 2];
      if (((am >> a) & 1ull) && ((am >> b) & 1ull)) {
        int shared = 0;
        for (int v = 0; v < nb_v; v++) {
          const uchar4 tv = S.ver[v];
          const bool ha = tv.x == a || tv.y == a || tv.z == a, hb = tv.x == b || tv.y == b || tv.z == b;
          shared += (ha && hb);
        }
        keep = shared >= 2;
      }
    }
    const unsigned m = group_ballot<G>(gmask, gshift, keep);
    __syncwarp(gmask);  // every lane has read its entry; writes below go to indices <= eb
    if (keep) {
      const int pos = ne_new + __popc(m & ((1u << lane) - 1u));
      S.edge[3 * pos] = (unsigned char)__popcll(am & ((1ull << a) - 1ull));
      S.edge[3 * pos + 1] = (unsigned char)__popcll(am & ((1ull << b) - 1ull));
      S.edge[3 * pos + 2] = z;
    }
    ne_new += __popc(m);
    __syncwarp(gmask);
  }
  for (int v = lane; v < nb_v; v += G) {
    uchar4 tv = S.ver[v];
    tv.x = (unsigned char)__popcll(am & ((1ull << tv.x) - 1ull));
    tv.y = (unsigned char)__popcll(am & ((1ull << tv.y) - 1ull));
    tv.z = (unsigned char)__popcll(am & ((1ull << tv.z) - 1ull));
    S.ver[v] = tv;
  }
  int np_new = 0;
  for (int pb = 0; pb < nb_p; pb += G) {
    const int pi = pb + lane;
    const bool keep = pi < nb_p && ((am >> pi) & 1ull);
    float4 pl = make_float4(0, 0, 0, 0);
    int nbp = 0;
    if (keep) {
      pl = S.plane[pi];
      nbp = S.pnb[pi];
    }
    const unsigned m = group_ballot<G>(gmask, gshift, keep);
    __syncwarp(gmask);
    if (keep) {
      const int pos = np_new + __popc(m & ((1u << lane) - 1u));
      S.plane[pos] = pl;
      S.pnb[pos] = nbp;
    }
    np_new += __popc(m);
    __syncwarp(gmask);
  }
  nb_p = np_new;
  nb_e = ne_new;
}

// a9: is_security_radius_reached (convex_cell.cu:240-268, the weighted variant), group-cooperative: lanes over the cell's
// vertices for v_dist = max |vertex - seed|^2 (compute_vertex_coordinates :319-351 in its literal float order), then
// d2 = |foot of the bisector on the segment seed-neighbour - seed|^2; reached iff d2 > 4 v_dist.  Opt-in path
// (mb_rpd_opts.security_radius), out of line.
template <int G, class CellS>
__device__ __noinline__ bool security_radius_reached(const CellS& S, int lane, unsigned gmask, int nb_v, float4 seed, float4 B) {
  float vd = 0.f;
  for (int v = lane; v < nb_v; v += G) {
    const uchar4 tv = S.ver[v];
    const float4 p1 = S.plane[tv.x], p2 = S.plane[tv.y], p3 = S.plane[tv.z];
    const float rx = -det3_exact(p1.w, p1.y, p1.z, p2.w, p2.y, p2.z, p3.w, p3.y, p3.z);
    const float ry = -det3_exact(p1.x, p1.w, p1.z, p2.x, p2.w, p2.z, p3.x, p3.w, p3.z);
    const float rz = -det3_exact(p1.x, p1.y, p1.w, p2.x, p2.y, p2.w, p3.x, p3.y, p3.w);
    const float rw = det3_exact(p1.x, p1.y, p1.z, p2.x, p2.y, p2.z, p3.x, p3.y, p3.z);
    const float dx = xfsub(__fdiv_rn(rx, rw), seed.x), dy = xfsub(__fdiv_rn(ry, rw), seed.y), dz = xfsub(__fdiv_rn(rz, rw), seed.z);
    const float d2 = dot3_exact(dx, dy, dz, dx, dy, dz);
    vd = d2 > vd ? d2 : vd;  // max(d2, v_dist): a NaN d2 is ignored, like the reference's max
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) vd = fmaxf(vd, __shfl_xor_sync(gmask, vd, o));
  const float fx = xfsub(seed.x, B.x), fy = xfsub(seed.y, B.y), fz = xfsub(seed.z, B.z);
  const float r2_diff = xfsub(seed.w, B.w);
  const float dd = dot3_exact(fx, fy, fz, fx, fy, fz);
  const float w = __fdiv_rn(xfsub(dd, r2_diff), xfmul(2.f, dd));
  const float px = xfsub(xfadd(xfmul(w, fx), B.x), seed.x), py = xfsub(xfadd(xfmul(w, fy), B.y), seed.y),
              pz = xfsub(xfadd(xfmul(w, fz), B.z), seed.z);
  return dot3_exact(px, py, pz, px, py, pz) > xfmul(4.f, vd);
}

// The FP64 determinant + static-filter test is the RARE path of the conflict predicate (~1 test in 3 000): kept out of
// line so that the hot loop stays small -- the kernel is sensitive to its instruction footprint (stall_no_instruction).
// bit 0: in conflict, bit 1: flagged
__device__ __noinline__ int conflict_exact_flag_ool(float4 p1, float4 p2, float4 p3, float4 e) {
  bool fl = false;
  const bool c = conflict_exact_flag(p1, p2, p3, e, fl);
  return (c ? 1 : 0) | (fl ? 2 : 0);
}

// PT = per-tet candidate lists (grid-kNN mode).  There array positions are not part of the contract
// (SURVEY 8a canonical form), so the order-defining serial replay of C2 is replaced by a
// group-cooperative partition + adjacency-bit-matrix cavity boundary; every predicate decision and
// every stored triple (cur_p, cir, next) is unchanged.
template <int G, bool PT, bool SMALL>
__global__ void __launch_bounds__(128, SMALL ? 5 : 4) k_clip(ClipArgs A) {
  constexpr int KP = SMALL ? MBK_SMALL_P : MBK_MAX_P, KT = SMALL ? MBK_SMALL_T : MBK_MAX_T,
                KE = SMALL ? MBK_SMALL_E : MBK_MAX_E;
  constexpr int GC_P0 = KP - 8, GC_E0 = SMALL ? 80 : 96;  // first garbage-collection thresholds (per-tet mode)
  typedef CellT<KP, KT, KE> CellS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CellS* cells = reinterpret_cast<CellS*>(smem_raw);
  constexpr int NG = 32 / G;  // groups per warp
  const int lane = threadIdx.x % G;
  const int wl = threadIdx.x & 31;
  const int gshift = (wl / G) * G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gshift);
  const int src = gshift;  // warp lane of the group's rank 0
  CellS& S = cells[threadIdx.x / G];
  __shared__ unsigned long long blk_cnt[16];  // block-aggregated counters (RpdCounters layout)
  if (threadIdx.x < 16) blk_cnt[threadIdx.x] = 0;
  if (PT) {
    // the cavity adjacency matrix is all-zero between clips (rows are cleared by their writers)
    for (int i = lane; i < KP; i += G) S.adj[i] = 0ull;
  }
  __syncthreads();

  // pair count and cursor granularity (device-side when the launch was speculative)
  const long long NP = A.work_count ? (long long)*A.work_count
                                    : (A.n_pairs_dev ? min((long long)*A.n_pairs_dev, A.n_pairs) : A.n_pairs);
  const int cursor_idx = A.work_list ? CNT_WORK_CURSOR2_IDX : CNT_WORK_CURSOR_IDX;
  int GRAB = A.grab;
  if (GRAB == 0) {
    long long g = NP / ((long long)gridDim.x * 4 * 16);
    g = (g / NG) * NG;
    GRAB = (int)max((long long)NG, min((long long)(8 * NG), g));
  }
  // warp-level work queue (chunks of pairs from the global cursor) and per-group scratch chunk
  long long wq_next = 0, wq_end = 0;
  bool wq_dry = false;
  unsigned long long chunk_at = 0;
  unsigned chunk_left = 0;
  unsigned n_clips = 0, n_culled = 0, n_valid = 0, n_exact = 0;

  // group state (identical in all lanes of a group unless noted)
  int state = GS_IDLE;
  long long pair = 0;
  int t = 0, seed_id = 0;
  float4 seed = make_float4(0, 0, 0, 0);
  unsigned hf4 = 0;  // (uchar) f_adjs of the 4 tet faces
  unsigned long long e6 = 0;
  const int* list = nullptr;
  int list_len = 0, base = 0;
  constexpr bool per_tet = PT;
  bool list_done = false, cull_ok = false;
  unsigned todo = 0, valid = 0, cvalid = 0;
  int nb = -1;                                // per lane: the neighbour of this lane's slot
  float4 eqn = make_float4(0, 0, 0, 0);       // per lane: its bisector
  int nb_v = 0, nb_p = 0, nb_e = 0, status = ST_success;
  int gc_next_e = GC_E0, gc_next_p = GC_P0;         // per-tet mode: garbage-collect dead planes / edges beyond these
  unsigned n_gc = 0;
  // flagged class (conflict_exact_flag): per lane "one of my tests fell under the static-filter bound"; pm_tet /
  // pm_bis = largest |normal component| over the tet planes / the bisectors seen so far (bound the filter's eps)
  bool flagged = false;
  float pm_tet = 0.f, pm_bis = 0.f;
  // a9 (given-neighbours mode, opt-in): radius reached / last listed neighbour seen so far
  bool sr_reached = false;
  int last_nb = -1;

  for (;;) {
    __syncwarp();
    // ================= A: hand pairs to idle groups =========================================
    {
      const unsigned idle = __ballot_sync(0xffffffffu, state == GS_IDLE && lane == 0);
      if (idle) {
        if (wq_next >= wq_end && !wq_dry) {
          unsigned long long b = 0;
          if (wl == 0) b = atomicAdd(&A.counters[cursor_idx], (unsigned long long)GRAB);
          b = __shfl_sync(0xffffffffu, b, 0);
          wq_next = (long long)b;
          wq_end = min((long long)b + GRAB, NP);
          if (wq_next >= NP) wq_dry = true;
        }
        if (state == GS_IDLE) {
          const int r = __popc(idle & ((1u << src) - 1u));
          if (wq_next + r < wq_end) {
            pair = A.work_list ? (long long)A.work_list[wq_next + r] : wq_next + r;
            state = GS_NEW;
          } else if (wq_dry) {
            state = GS_EXIT;
          }
        }
        const long long avail = wq_end > wq_next ? wq_end - wq_next : 0;
        wq_next += min((long long)__popc(idle), avail);
      }
      if (__ballot_sync(0xffffffffu, state != GS_EXIT) == 0) break;
    }
    // ================= A2: initialise new cells (ConvexCell ctor, convex_cell.cu:116-214) =======
    if (state == GS_NEW) {
      t = A.pair_tet[pair];
      seed_id = A.pair_site[pair];
      const unsigned vadj4 = A.tet_vadj[t];
      const int4 fadj = A.tet_fadj[t];
      const int4 fid = A.tet_fid[t];
      const uint2 e6u = A.tet_e6[t];
      e6 = ((unsigned long long)e6u.y << 32) | e6u.x;
      seed = A.site4[seed_id];
      hf4 = (unsigned)(unsigned char)fadj.x | ((unsigned)(unsigned char)fadj.y << 8) |
            ((unsigned)(unsigned char)fadj.z << 16) | ((unsigned)(unsigned char)fadj.w << 24);
      bool ok0 = true;
      if (lane < 4) {
        // face planes and initial-vertex cofactors were computed once per tet (k_tet_geometry)
        S.plane[lane] = A.tet_geo[(size_t)t * 8 + lane];
        const float4 c = A.tet_geo[(size_t)t * 8 + 4 + lane];
        S.c0[lane] = c;   // tet-level cull filter
        S.cof[lane] = c;  // first four entries of the per-vertex filter cache
        S.pnb[lane] = lane == 0 ? fid.x : (lane == 1 ? fid.y : (lane == 2 ? fid.z : fid.w));
        // dual triangles (1,3,2) (0,2,3) (0,3,1) (0,1,2) with w = (uchar)v_adjs (:186-189)
        const unsigned char w = (unsigned char)((vadj4 >> (8 * lane)) & 0xffu);
        S.ver[lane] = lane == 0 ? make_uchar4(1, 3, 2, w)
                                : (lane == 1 ? make_uchar4(0, 2, 3, w)
                                             : (lane == 2 ? make_uchar4(0, 3, 1, w) : make_uchar4(0, 1, 2, w)));
        // a proper vertex has c.w < 0 (conflict <=> det > 0 <=> vertex on the negative side)
        ok0 = (c.w < 0.f) && isfinite(c.x) && isfinite(c.y) && isfinite(c.z) && isfinite(c.w);
        const float4 pl = S.plane[lane];
        pm_tet = fmaxf(fabsf(pl.x), fmaxf(fabsf(pl.y), fabsf(pl.z)));
      } else {
        pm_tet = 0.f;
      }
      pm_tet = fmaxf(pm_tet, __shfl_xor_sync(gmask, pm_tet, 1));
      pm_tet = fmaxf(pm_tet, __shfl_xor_sync(gmask, pm_tet, 2));
      pm_tet = __shfl_sync(gmask, pm_tet, src);
      pm_bis = 0.f;
      flagged = false;
      if (lane >= G - 6 || G < 8) {
        // edges (2,3)(1,3)(1,2)(0,3)(0,2)(0,1) with the e_adj of vertex pairs (0,1)(0,2)(0,3)(1,2)(1,3)(2,3)
        for (int q = (G < 8 ? lane : lane - (G - 6)); q < 6; q += (G < 8 ? G : 6)) {
          const unsigned char ea = q < 3 ? (q == 0 ? 2 : 1) : 0;
          const unsigned char eb = q == 0 ? 3 : (q == 1 ? 3 : (q == 2 ? 2 : (q == 3 ? 3 : (q == 4 ? 2 : 1))));
          S.edge[3 * q + 0] = ea;
          S.edge[3 * q + 1] = eb;
          S.edge[3 * q + 2] = (unsigned char)((e6 >> (8 * q)) & 0xff);
        }
      }
      cull_ok = group_ballot<G>(gmask, gshift, !ok0) == 0 && !A.no_cull && !(!PT && A.security_radius);
      sr_reached = false;
      last_nb = -1;
      cvalid = 0xfu;
      nb_v = 4;
      nb_p = 4;
      nb_e = 6;
      status = ST_success;
      if (per_tet) {
        const int tl = A.pair_local[pair];
        list = A.nbr + (size_t)tl * A.nbr_stride;
        list_len = A.nbr_cnt[tl];
      } else {
        list = A.nbr + (size_t)seed_id * A.nbr_stride;
        list_len = A.nbr_stride;
      }
      base = 0;
      todo = 0;
      valid = 0;
      list_done = false;
      gc_next_e = GC_E0;
      gc_next_p = GC_P0;
      state = GS_RUN;
      __syncwarp(gmask);
    }
    // ================= B: scan the next G listed neighbours =====================================
    if (state == GS_RUN && todo == 0) {
      if (list_done || base >= list_len) {
        state = GS_FINISH;
      } else {
        const int j = base + lane;
        nb = (j < list_len) ? list[j] : -1;
        bool is_end = false, cand = false;
        if (per_tet) {
          cand = (nb >= 0 && nb != seed_id);  // the list is the tet's candidate set; skip the seed
        } else {
          is_end = (j < list_len) && (nb == -1);  // the first -1 terminates the list (:1259)
          cand = (nb >= 0);
        }
        const bool valid_nb = cand;
        if (cand) {
          eqn = bisector_exact(seed, A.site4[nb]);
          if (cull_ok) {
            const float n1 = fabsf(eqn.x) + fabsf(eqn.y) + fabsf(eqn.z);
            const float nmax = fmaxf(fabsf(eqn.x), fmaxf(fabsf(eqn.y), fabsf(eqn.z)));
            // a culled plane must also be clear of the flagged class: |det| >= |s| - 4e-7 T > eps
            const float eps_up = filter_eps_upper(fmaxf(pm_tet, nmax));
            bool all_out = true;
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const float4 c = S.c0[i];
              const float cw = c.w * eqn.w;
              const float s = fmaf(c.x, eqn.x, fmaf(c.y, eqn.y, fmaf(c.z, eqn.z, cw)));
              const float T = fmaf(fabsf(c.x) + fabsf(c.y) + fabsf(c.z), nmax, fabsf(cw));
              const float margin = fmaxf(fmaxf(4e-6f * T, 1e-3f * n1 * fabsf(c.w)), fmaf(4e-7f, T, eps_up));
              all_out = all_out && (s < -margin);
            }
            if (all_out) {
              cand = false;  // certainly removes no vertex: the reference would pop this plane
              n_culled++;
            }
          }
        }
        const unsigned end_mask = group_ballot<G>(gmask, gshift, is_end);
        todo = group_ballot<G>(gmask, gshift, cand);
        // every listed neighbour (culled or not) makes the reference call new_plane, which refuses
        // a 65th plane (vertex_overflow, convex_cell.cu:562-565)
        valid = group_ballot<G>(gmask, gshift, valid_nb);
        if (end_mask) {
          const unsigned before = (1u << (__ffs(end_mask) - 1)) - 1u;
          todo &= before;
          valid &= before;
          list_done = true;
        }
        base += G;
        if (nb_p >= KP && valid) {
          status = ST_vertex_overflow;
          todo = 0;
          state = GS_FINISH;
        } else if (todo == 0 && (list_done || base >= list_len)) {
          state = GS_FINISH;  // nothing left to clip: the record is written in this same iteration
        }
      }
    }
    // ================= B2: garbage collection (per-tet candidate lists only) =====================
    // The reference's caps (64 planes, 152 edges) count DEAD entries.  With the reference's short
    // per-site lists that is harmless; a per-tet candidate list can be several times longer, so
    // before a clip that could hit a cap the dead planes (referenced by no vertex) and inactive
    // edges (fewer than 2 live vertices on both planes, the reload_active criterion of
    // voronoi_defs.cxx:92-103) are dropped and the survivors renumbered.  Tet faces keep indices
    // 0..3; the canonical form (active planes / edges, vertices) is unchanged.  Rare path.
    if (per_tet && state == GS_RUN && todo != 0 && (nb_e >= gc_next_e || nb_p >= gc_next_p)) {
      cell_gc<G, CellS>(S, lane, gmask, gshift, nb_v, nb_p, nb_e);
      gc_next_e = max(GC_E0, nb_e + 24);
      gc_next_p = max(GC_P0, nb_p + 4);
      if (lane == 0) n_gc++;
    }
    // ================= C: clip by the next surviving plane (clip_by_plane, :680-774) ============
    // Every sub-phase below sits at the top level of the loop with warp-uniform trip counts and a
    // full-warp barrier in front, so that the groups of the warp execute it TOGETHER.
    const bool c_act = (state == GS_RUN && todo != 0);
    int k = 0, nbk = -1;
    float4 e = make_float4(0, 0, 0, 0);
    {
      // all lanes shuffle (the source index is per lane; idle groups read garbage they never use)
      k = c_act ? (__ffs(todo) - 1) : 0;
      nbk = __shfl_sync(0xffffffffu, nb, src + k);
      e.x = __shfl_sync(0xffffffffu, eqn.x, src + k);
      e.y = __shfl_sync(0xffffffffu, eqn.y, src + k);
      e.z = __shfl_sync(0xffffffffu, eqn.z, src + k);
      e.w = __shfl_sync(0xffffffffu, eqn.w, src + k);
      if (c_act) {
        todo &= todo - 1;
        if (lane == 0) n_clips++;
      }
    }
    // ---- C1: conflict flags of all vertices (lanes strided over vertices) ----------------------
    unsigned long long f0 = 0;
    unsigned f1 = 0;
    int nb_r = 0;
    {
      const int vmax = __reduce_max_sync(0xffffffffu, c_act ? nb_v : 0);
      const float nmax = fmaxf(fabsf(e.x), fmaxf(fabsf(e.y), fabsf(e.z)));
      // flagged-class guard of the FP32 filter: eps <= K M^5 with M over the planes a vertex can involve -- the
      // bisectors only (eps_b) unless the vertex lies on a tet face (eps_t)
      const float m_b = fmaxf(pm_bis, nmax);
      const float eps_b = filter_eps_upper(m_b), eps_t = filter_eps_upper(fmaxf(pm_tet, m_b));
      if (c_act) pm_bis = m_b;
      for (int vb = 0; vb < vmax; vb += G) {
        const int v = vb + lane;
        bool cf = false;
        if (c_act && v < nb_v) {
          bool decided = false;
          const uchar4 tv = S.ver[v];
          if (v < MBK_CV && ((cvalid >> v) & 1u)) {
            // filtered predicate: |s - det_fp64| <= ~3e-7 * T, accepted beyond 4e-6 * T -- and only when that
            // also proves |det_fp64| above the static-filter bound (the test cannot be a flagged one)
            const float4 c = S.cof[v];
            const float cw = c.w * e.w;
            const float s = fmaf(c.x, e.x, fmaf(c.y, e.y, fmaf(c.z, e.z, cw)));
            const float T = fmaf(fabsf(c.x) + fabsf(c.y) + fabsf(c.z), nmax, fabsf(cw));
            const float eps_up = min(tv.x, min(tv.y, tv.z)) < 4 ? eps_t : eps_b;
            if (fabsf(s) > fmaxf(4e-6f * T, fmaf(4e-7f, T, eps_up))) {
              cf = s > 0.f;
              decided = true;
            }
          }
          if (!decided) {
            const int r2 = conflict_exact_flag_ool(S.plane[tv.x], S.plane[tv.y], S.plane[tv.z], e);
            cf = (r2 & 1) != 0;
            flagged = flagged || (r2 & 2) != 0;
            n_exact++;
          }
        }
        const unsigned m = (__ballot_sync(0xffffffffu, cf) >> gshift) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
        nb_r += __popc(m);
        if (KT <= 64 || vb < 64)  // G divides 64: a round never straddles the two words
          f0 |= (unsigned long long)m << vb;
        else
          f1 |= m << (vb - 64);
      }
    }
    // 0: nothing to do (idle, or the plane removes nothing and is dropped, :741-744)
    // 1: clip   2: the plane removes everything (:746-749)
    const int act2 = !c_act ? 0 : (nb_r == nb_v ? 2 : (nb_r != 0 ? 1 : 0));
    if (act2 == 2) {
      status = ST_no_intersection;
      todo = 0;
      state = GS_FINISH;
    }
    __syncwarp();
    // ---- C2: one lane per group replays the order-defining serial bookkeeping -----------------
    int L = 0;
    int st2 = ST_success;
    unsigned cv = cvalid;
    if (PT) {
      // ---- C2 (grid-kNN mode): group-cooperative, no order replay ------------------------------
      const bool a1 = (act2 == 1);
      const int nv_new = nb_v - nb_r;
      if (a1 && lane == 0) {
        S.plane[nb_p] = e;
        S.pnb[nb_p] = nbk;
      }
      // (i) partition: the i-th removed vertex of the head [0, nv_new) trades places with the i-th
      // kept vertex of the tail [nv_new, nb_v).  Pairs are disjoint: no barrier between read and write.
      {
        const int hmax = __reduce_max_sync(0xffffffffu, a1 ? nv_new : 0);
        // kept vertices of the tail
        unsigned long long k0 = ~f0;
        unsigned k1 = ~f1;
        if (KT <= 64) {
          k0 &= ~((1ull << nv_new) - 1ull);  // nv_new < nb_v <= 64
          if (nb_v < 64) k0 &= (1ull << nb_v) - 1ull;
          k1 = 0;
        } else {
          if (nv_new < 64) k0 &= ~((1ull << nv_new) - 1ull); else { k0 = 0; k1 &= ~((1u << (nv_new - 64)) - 1u); }
          if (nb_v < 64) { k0 &= (1ull << nb_v) - 1ull; k1 = 0; } else if (nb_v < 96) k1 &= (1u << (nb_v - 64)) - 1u;
        }
        for (int vb = 0; vb < hmax; vb += G) {
          const int v = vb + lane;
          const bool fv = (KT <= 64 || v < 64) ? ((f0 >> (v & 63)) & 1ull) : ((f1 >> (v - 64)) & 1u);
          const bool mv = a1 && v < nv_new && fv;
          bool pvalid = false;
          if (mv) {
            // rank among the removed vertices of the head
            int i = (KT <= 64 || v < 64) ? __popcll(f0 & ((1ull << (v & 63)) - 1ull))
                                         : (__popcll(f0) + __popc(f1 & ((1u << (v - 64)) - 1u)));
            int partner;
            const int pc0 = __popcll(k0);
            if (KT <= 64 || i < pc0) {
              unsigned long long kk = k0;
              for (; i > 0; i--) kk &= kk - 1ull;
              partner = __ffsll((long long)kk) - 1;
            } else {
              unsigned kk = k1;
              for (i -= pc0; i > 0; i--) kk &= kk - 1u;
              partner = 64 + __ffs((int)kk) - 1;
            }
            const uchar4 tv = S.ver[v], tp = S.ver[partner];
            pvalid = partner < MBK_CV && ((cv >> partner) & 1u);
            S.ver[v] = tp;
            S.ver[partner] = tv;
            if (pvalid && v < MBK_CV) S.cof[v] = S.cof[partner];
          }
          if (vb < 32) {
            const unsigned mm = (__ballot_sync(0xffffffffu, mv) >> gshift) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
            const unsigned ms = (__ballot_sync(0xffffffffu, mv && pvalid) >> gshift) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
            cv = (cv & ~(mm << vb)) | (ms << vb);
          }
        }
      }
      __syncwarp();
      // (ii) cavity boundary: directed dual edges a->b of the removed triangles go into a 64x64 bit
      // matrix; an edge is on the boundary iff its twin b->a is absent (compute_boundary :618-678
      // builds the same circular list one triangle at a time)
      const int rmax = __reduce_max_sync(0xffffffffu, a1 ? nb_r : 0);
      for (int rb = 0; rb < rmax; rb += G) {
        const int r = rb + lane;
        if (a1 && r < nb_r) {
          const uchar4 tv = S.ver[nv_new + r];
          unsigned* m32 = reinterpret_cast<unsigned*>(S.adj);
          atomicOr(&m32[2 * tv.x + (tv.y >> 5)], 1u << (tv.y & 31));
          atomicOr(&m32[2 * tv.y + (tv.z >> 5)], 1u << (tv.z & 31));
          atomicOr(&m32[2 * tv.z + (tv.x >> 5)], 1u << (tv.x & 31));
        }
      }
      if (a1) {  // clear the circular list (64 bytes per cell)
        if (G >= 8) { if (lane < KP / 8) reinterpret_cast<uint2*>(S.bnext)[lane] = make_uint2(0xffffffffu, 0xffffffffu); }
        else { if (lane < KP / 16) reinterpret_cast<uint4*>(S.bnext)[lane] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu); }
      }
      __syncwarp();
      int nbnd = 0, first = MBK_END;
      for (int rb = 0; rb < rmax; rb += G) {
        const int r = rb + lane;
        if (a1 && r < nb_r) {
          const uchar4 tv = S.ver[nv_new + r];
          const unsigned char pl[3] = {tv.x, tv.y, tv.z};
#pragma unroll
          for (int q = 0; q < 3; q++) {
            const int a = pl[q], b = pl[(q + 1) % 3];
            if (!((S.adj[b] >> a) & 1ull)) {
              S.bnext[a] = (unsigned char)b;
              nbnd++;
              first = min(first, a);
            }
          }
        }
      }
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) {
        nbnd += __shfl_xor_sync(0xffffffffu, nbnd, o);
        first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
      }
      __syncwarp();
      for (int rb = 0; rb < rmax; rb += G) {
        const int r = rb + lane;
        if (a1 && r < nb_r) {
          const uchar4 tv = S.ver[nv_new + r];
          S.adj[tv.x] = 0ull;
          S.adj[tv.y] = 0ull;
          S.adj[tv.z] = 0ull;
        }
      }
      // (iii) the cycle, from its smallest plane id (deterministic); a cavity whose boundary is not
      // one simple cycle is the reference's inconsistent_boundary (:626-629)
      if (a1 && lane == 0 && first != MBK_END) {
        int cir = first;
        do {
          S.cyc[L++] = (unsigned char)cir;
          cir = S.bnext[cir];
        } while (cir != first && cir != MBK_END && L < nbnd && L < KP);
        if (cir != first || L != nbnd) st2 = ST_inconsistent_boundary;
      }
    } else
    if (act2 == 1 && lane == 0) {
      const int cur_p = nb_p;
      S.plane[cur_p] = e;
      S.pnb[cur_p] = nbk;
      // swap partition, convex_cell.cu:706-721 (the filter cache follows the kept vertices)
      int nv = nb_v, i = 0;
      if (nb_v <= 32) {
        // common case: all flags in one 32-bit word
        unsigned f = (unsigned)f0;
        while (i < nv) {
          if ((f >> i) & 1u) {
            nv--;
            const unsigned fn = (f >> nv) & 1u;
            const uchar4 tmp = S.ver[i];
            S.ver[i] = S.ver[nv];
            S.ver[nv] = tmp;
            const unsigned vn = (cv >> nv) & 1u;
            if (vn) S.cof[i] = S.cof[nv];
            cv = (cv & ~(1u << i)) | (vn << i);
            f = (f & ~(1u << i)) | (fn << i);
          } else
            i++;
        }
      } else
      while (i < nv) {
        const bool fi = i < 64 ? ((f0 >> i) & 1ull) : ((f1 >> (i - 64)) & 1u);
        if (fi) {
          nv--;
          const bool fn = nv < 64 ? ((f0 >> nv) & 1ull) : ((f1 >> (nv - 64)) & 1u);
          const uchar4 tmp = S.ver[i];
          S.ver[i] = S.ver[nv];
          S.ver[nv] = tmp;
          if (i < MBK_CV) {
            const bool vn = nv < MBK_CV && ((cv >> nv) & 1u);
            if (vn) S.cof[i] = S.cof[nv];
            cv = (cv & ~(1u << i)) | ((unsigned)vn << i);
          }
          if (i < 64)
            f0 = (f0 & ~(1ull << i)) | ((unsigned long long)fn << i);
          else
            f1 = (f1 & ~(1u << (i - 64))) | ((unsigned)fn << (i - 64));
        } else
          i++;
      }
      // cavity boundary, compute_boundary convex_cell.cu:618-678
      {
        uint4* bn = reinterpret_cast<uint4*>(S.bnext);
        const uint4 ff = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        bn[0] = ff;
        if (cur_p >= 16) bn[1] = ff;
        if (cur_p >= 32) {
          bn[2] = ff;
          if (KP > 48) bn[3] = ff;
        }
      }
      int first = MBK_END;
      int r = nb_r, tt = nv, fails = 0;
      while (r > 0) {
        const uchar4 tv = S.ver[tt];
        const unsigned char pl[3] = {tv.x, tv.y, tv.z};
        bool in_border[3], opp[3];
#pragma unroll
        for (int q = 0; q < 3; q++) in_border[q] = S.bnext[pl[q]] != MBK_END;
#pragma unroll
        for (int q = 0; q < 3; q++) opp[q] = S.bnext[pl[(q + 1) % 3]] == pl[q];
        bool simple = true;
#pragma unroll
        for (int q = 0; q < 3; q++)
          if (!opp[q] && !opp[(q + 1) % 3] && in_border[(q + 1) % 3]) simple = false;
        if (!opp[0] && !opp[1] && !opp[2]) {
          if (first == MBK_END) {
#pragma unroll
            for (int q = 0; q < 3; q++) S.bnext[pl[q]] = pl[(q + 1) % 3];
            first = pl[0];
          } else
            simple = false;
        }
        if (!simple) {
          tt++;
          if (tt == nv + r) tt = nv;
          if (++fails >= r) {  // a full round without progress: the reference spins until
            st2 = ST_inconsistent_boundary;  // nb_iter > 65535 (:626-629)
            break;
          }
          continue;
        }
        fails = 0;
#pragma unroll
        for (int q = 0; q < 3; q++)
          if (!opp[q]) S.bnext[pl[q]] = pl[(q + 1) % 3];
#pragma unroll
        for (int q = 0; q < 3; q++)
          if (opp[q] && opp[(q + 1) % 3]) {
            const unsigned char pm = pl[(q + 1) % 3];
            if (first == pm) first = S.bnext[pm];
            S.bnext[pm] = MBK_END;
          }
        const uchar4 tmp = S.ver[tt];
        S.ver[tt] = S.ver[nv + r - 1];
        S.ver[nv + r - 1] = tmp;
        tt = nv;
        r--;
      }
      if (st2 == ST_success && first != MBK_END) {
        int cir = first;
        do {
          S.cyc[L++] = (unsigned char)cir;
          cir = S.bnext[cir];
        } while (cir != first && cir != MBK_END && L < KP);
      }
    }
    __syncwarp();
    L = __shfl_sync(0xffffffffu, L, src);
    st2 = __shfl_sync(0xffffffffu, st2, src);
    cv = __shfl_sync(0xffffffffu, cv, src);
    const int cur_p = nb_p;
    bool do_new = false;
    if (act2 == 1) {
      nb_p++;
      nb_v -= nb_r;
      // slots >= nb_v no longer hold live vertices
      cvalid = nb_v >= 32 ? cv : (cv & ((1u << nb_v) - 1u));
      if (st2 != ST_success)
        status = st2;
      else if (L != 0) {  // first_boundary_ != END_OF_LIST (:754)
        if (nb_e + L > KE)
          status = ST_edge_overflow;
        else
          do_new = true;
      }
    }
    // ---- C3: new edges (:756-762) and new vertices (:764-773), lanes over the cycle -----------
    {
      const int Lmax = __reduce_max_sync(0xffffffffu, do_new ? L : 0);
      bool perturb = false;
      for (int jb = 0; jb < Lmax; jb += G) {
        const int jj = jb + lane;
        bool pj = false;
        if (do_new && jj < L) {
          const int cir = S.cyc[jj];
          const int nxt = S.cyc[jj + 1 == L ? 0 : jj + 1];
          const unsigned char z1 = edge_z(cir, cur_p, hf4, e6);
          S.edge[3 * (nb_e + jj) + 0] = (unsigned char)cir;
          S.edge[3 * (nb_e + jj) + 1] = (unsigned char)cur_p;
          S.edge[3 * (nb_e + jj) + 2] = z1;
          const unsigned char z2 = edge_z(nxt, cur_p, hf4, e6);
          const unsigned char z3 = edge_z(min(cir, nxt), max(cir, nxt), hf4, e6);
          const unsigned char w = max(max(z1, z2), z3);
          const float4 p2 = S.plane[cir], p3 = S.plane[nxt];
          const int slot = nb_v + jj;
          if (slot + 1 < KT) {
            S.ver[slot] = make_uchar4(cur_p, cir, nxt, w);
            if (slot < MBK_CV) S.cof[slot] = cofactors_f32(minors_exact(e, p2, p3));
          }
          // is_vertex_perturb (:274-316): w-component of the vertex == 0
          const float wdet = det3_exact(e.x, e.y, e.z, p2.x, p2.y, p2.z, p3.x, p3.y, p3.z);
          pj = (wdet == 0.f);
        }
        const unsigned pm = (__ballot_sync(0xffffffffu, pj) >> gshift) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
        if (pm && !perturb) {
          perturb = true;
          const int jp = jb + __ffs(pm) - 1;          // first perturbed vertex
          const int jo = KT - 1 - nb_v;        // first overflowing vertex (if < L)
          status = (jo < L && jo < jp) ? ST_triangle_overflow : ST_needs_perturb;
        }
      }
      if (do_new) {
        if (!perturb && nb_v + L + 1 > KT) status = ST_triangle_overflow;  // nb_v+1 >= 96 (:525)
        // the new vertices occupy slots [nb_v, nb_v+L): their cache entries are fresh
        const int lo = min(nb_v, 32), hi = min(nb_v + L, 32);
        const unsigned add = (hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u)) & ~((lo >= 32) ? 0xffffffffu : ((1u << lo) - 1u));
        cvalid |= add;
        nb_e += L;
        nb_v += L;
      }
    }
    if (act2 == 1) {
      // a later listed neighbour of this batch would be refused by new_plane (:562-565)
      if (status == ST_success && nb_p >= KP && (todo || (valid & ~((2u << k) - 1u)))) status = ST_vertex_overflow;
      if (status != ST_success) {
        todo = 0;
        state = GS_FINISH;
      }
    }
    // a9: after every listed neighbour, has the security radius been reached?  (convex_cell.cu:1285-1296)
    if (!PT && A.security_radius && c_act && state == GS_RUN && status == ST_success) {
      last_nb = nbk;
      if (security_radius_reached<G, CellS>(S, lane, gmask, nb_v, seed, A.site4[nbk])) {
        sr_reached = true;
        todo = 0;
        list_done = true;
      }
    }
    // the last plane of the list has been dealt with (clipped or dropped): finish in this same iteration
    if (c_act && state == GS_RUN && todo == 0 && (list_done || base >= list_len)) state = GS_FINISH;
    __syncwarp();
    // ================= D: write the record (copy(), convex_cell.cu:933-949) =======================
    if (state == GS_FINISH) {
      // a9: a cell whose last listed neighbour does not reach the security radius (convex_cell.cu:1304-1316)
      if (!PT && A.security_radius && status == ST_success && !sr_reached) {
        if (last_nb < 0 || !security_radius_reached<G, CellS>(S, lane, gmask, nb_v, seed, A.site4[last_nb]))
          status = ST_security_radius_not_reached;
      }
      long long blob_at = -1;
      int words = 0;
      const unsigned fbit = group_ballot<G>(gmask, gshift, flagged) ? MB_FLAG_BIT : 0u;
      if (status == ST_success) {
        words = compact_words(nb_v, nb_p, nb_e);
        unsigned long long at = 0;
        if (lane == 0) {
          if ((unsigned)words > chunk_left) {  // one global atomic per ~13 cells instead of per cell
            const unsigned grab = max(CLIP_CHUNK_WORDS, (unsigned)words);
            chunk_at = atomicAdd(&A.counters[CNT_BLOB], (unsigned long long)grab);
            chunk_left = grab;
          }
          at = chunk_at;
          chunk_at += words;
          chunk_left -= words;
        }
        at = __shfl_sync(gmask, at, src);
        if (at + (unsigned long long)words <= A.scratch_words) {
          blob_at = (long long)at;
          uint32_t* o = A.scratch + at;
          if (lane == 0) {
            o[0] = (uint32_t)(t + A.tet_id_base);
            o[1] = (uint32_t)seed_id;
            o[2] = (uint32_t)nb_v | ((uint32_t)nb_p << 8) | ((uint32_t)nb_e << 16) | ((uint32_t)status << 24) | fbit;
            o[3] = __float_as_uint(seed.w);
          }
          o += 4;
          const uint32_t* sv = reinterpret_cast<const uint32_t*>(S.ver);
          for (int i = lane; i < nb_v; i += G) o[i] = sv[i];
          o += nb_v;
          // planes are written word-wise: the record is only 4-byte aligned
          for (int i = lane; i < 4 * nb_p; i += G) o[i] = reinterpret_cast<const uint32_t*>(S.plane)[i];
          o += 4 * nb_p;
          for (int i = lane; i < nb_p; i += G) {
            int ida, idb;
            float h;
            if (i < 4) {
              ida = S.pnb[i];
              idb = -1;
              h = (float)((hf4 >> (8 * i)) & 0xffu);
            } else {
              const int nbid = S.pnb[i];
              ida = min(seed_id, nbid);
              idb = max(seed_id, nbid);
              h = 1.f;
            }
            o[3 * i + 0] = (uint32_t)ida;
            o[3 * i + 1] = (uint32_t)idb;
            o[3 * i + 2] = __float_as_uint(h);
          }
          o += 3 * nb_p;
          const int ew = (3 * nb_e + 3) / 4;
          const uint32_t* se = reinterpret_cast<const uint32_t*>(S.edge);
          // the pad bytes of the last word are zeroed: the blob is deterministic byte for byte
          const uint32_t tail_mask = (3 * nb_e) & 3 ? (0xffffffffu >> (8 * (4 - ((3 * nb_e) & 3)))) : 0xffffffffu;
          for (int i = lane; i < ew; i += G) o[i] = (i == ew - 1) ? (se[i] & tail_mask) : se[i];
        }
      }
      // a cell that outgrew the compact caps is not final: the second pass recomputes it at the reference's caps
      const bool redo = SMALL && (status == ST_triangle_overflow || status == ST_vertex_overflow || status == ST_edge_overflow);
      if (redo && lane == 0) {
        const unsigned long long at = atomicAdd(&A.counters[CNT_REDO], 1ull);
        A.redo_out[at] = (int)pair;
        A.pair_status[pair] = (signed char)ST_early_return;
        A.pair_blob[pair] = -1;
        A.pair_words[pair] = 0;
      }
      if (!redo && lane == 0) {
        A.pair_status[pair] = (signed char)status;
        A.pair_blob[pair] = blob_at;
        // low 16 bits: record words; high 16 bits: nb_p (the lean transport format drops 4 * nb_p words)
        // bit 30: flagged class (kept for pairs without a record too: mb_rpd_fetch_flags)
        A.pair_words[pair] = (int)(((blob_at >= 0) ? (unsigned)(words | (nb_p << 16)) : 0u) | fbit);
        if (status == ST_success && blob_at >= 0) n_valid++;
        if (fbit) {  // rare: straight to the global counters
          atomicAdd(&A.counters[CNT_FLAG_PAIRS], 1ull);
          if (status == ST_success && blob_at >= 0) atomicAdd(&A.counters[CNT_FLAG_CELLS], 1ull);
        }
        if (status != ST_success) atomicAdd(&blk_cnt[CNT_HIST + status + 1], 1ull);
      }
      state = GS_IDLE;
      __syncwarp(gmask);
    }
  }
  // ---- statistics: per-lane registers -> block -> global -------------------------------------
  atomicAdd(&blk_cnt[CNT_CULLED], (unsigned long long)n_culled);
  atomicAdd(&blk_cnt[CNT_EXACT], (unsigned long long)n_exact);
  if (lane == 0) {
    atomicAdd(&blk_cnt[CNT_CLIPS], (unsigned long long)n_clips);
    atomicAdd(&blk_cnt[CNT_VALID], (unsigned long long)n_valid);
    if (n_gc) atomicAdd(&A.counters[18], (unsigned long long)n_gc);
    atomicAdd(&blk_cnt[CNT_HIST + ST_success + 1], (unsigned long long)n_valid);
  }
  __syncthreads();
  if (threadIdx.x >= 1 && threadIdx.x < 16 && blk_cnt[threadIdx.x])
    atomicAdd(&A.counters[threadIdx.x], blk_cnt[threadIdx.x]);
}
