// K3, grid-kNN mode, first pass: batch-synchronous sub-warp clipping with compact caps.
//
// k_clip (rpd_clip.cuh) walks a per-group state machine whose whole body -- assign, init, scan, garbage collection,
// clip, write -- is executed by the warp on EVERY iteration as long as one of its groups needs any phase; ncu shows the
// consequence: ~1 600 warp instructions per iteration, 16 of 32 lanes active, and an instruction footprint (68 KB)
// that misses the instruction cache (stall_no_instruction).  Array positions are not part of the grid-kNN contract
// (canonical form), so this kernel is organised around the statistics of the workload instead -- cells
// are small (8 vertices, 8 planes, 2-3 cutting bisectors):
//   * the NG = 32/G groups of a warp take NG consecutive pairs (almost always cells of the same tet: same candidate
//     list, same number of clips) and run init -> scan -> [clip]* -> write TOGETHER; only the clip loop iterates, so
//     init / scan / write are executed once per cell instead of once per iteration;
//   * no swap partition: removed vertices leave holes that the new vertices of the same clip fill (slot list `rm`
//     built by ballot rank while the conflict flags are computed); the rare clip that removes more vertices than it
//     creates compacts the tail afterwards;
//   * the directed dual edges of the removed triangles are marked in the adjacency bit matrix inside the conflict
//     loop itself (one pass less);
//   * compact caps (24 planes / 32 vertices / 96 edges, 1.7 KB per cell): every vertex has a cofactor-filter entry,
//     the conflict flags fit 32 bits, no garbage collection code.  A cell that outgrows them is handed to k_clip at the
//     reference's caps (redo list), exactly like k_clip's own compact pass does.
// Decisions (filtered predicate -> FP64 det4x4, flagged class), plane equations and the stored triples
// (cur_p, cir, next) are those of k_clip; only array positions differ.
#pragma once

#include "rpd_clip.cuh"

#define MBK_TINY_P 24
#define MBK_TINY_T 32
#define MBK_TINY_E 96

struct __align__(16) CellTiny {
  float4 plane[MBK_TINY_P];
  float4 c0[4];
  float4 cof[MBK_TINY_T];
  int pnb[MBK_TINY_P];
  uchar4 ver[MBK_TINY_T];
  unsigned adj[MBK_TINY_P];            // directed dual-edge bit matrix of the cavity (24 planes: one word per row)
  unsigned char bnext[32];             // cavity boundary successor per plane (16-byte aligned, cleared with vector stores)
  unsigned char cyc[32];               // the boundary cycle in walk order
  unsigned char rm[MBK_TINY_T];        // slots of the vertices removed by the current clip, ascending
  unsigned char edge[MBK_TINY_E * 3];
};
static_assert(offsetof(CellTiny, bnext) % 16 == 0, "bnext must be 16-byte aligned");
static_assert(sizeof(CellTiny) % 16 == 0, "cells are 16-byte aligned");

// PT = true: per-tet candidate lists (grid-kNN mode), hole-filling clip.  PT = false: per-site neighbour lists
// (given-neighbours mode): the same batch-synchronous structure, but the order-defining bookkeeping of clip_by_plane --
// swap partition (convex_cell.cu:706-721) and compute_boundary (:618-678) -- is replayed literally by one lane per
// group, so array positions, and with them the records, stay byte-identical to the reference.
template <int G, int NB, bool PT>
__global__ void __launch_bounds__(128, NB) k_clip_tiny(ClipArgs A) {
  constexpr int KP = MBK_TINY_P, KT = MBK_TINY_T, KE = MBK_TINY_E;
  constexpr int NG = 32 / G;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CellTiny* cells = reinterpret_cast<CellTiny*>(smem_raw);
  const int lane = threadIdx.x % G;
  const int wl = threadIdx.x & 31;
  const int gi = wl / G;               // group index in the warp
  const int gshift = gi * G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gshift);
  const unsigned glow = (G == 32) ? 0xffffffffu : ((1u << G) - 1u);
  const int src = gshift;
  CellTiny& S = cells[threadIdx.x / G];
  __shared__ unsigned long long blk_cnt[16];
  if (threadIdx.x < 16) blk_cnt[threadIdx.x] = 0;
  if (PT)
    for (int i = lane; i < KP; i += G) S.adj[i] = 0u;
  __syncthreads();

  const long long NP = A.n_pairs_dev ? min((long long)*A.n_pairs_dev, A.n_pairs) : A.n_pairs;
  int GRAB = A.grab;
  if (GRAB == 0) {
    long long g = NP / ((long long)gridDim.x * 4 * 16);
    g = (g / NG) * NG;
    GRAB = (int)max((long long)NG, min((long long)(8 * NG), g));
  }
  unsigned long long chunk_at = 0;
  unsigned chunk_left = 0;
  unsigned n_clips = 0, n_culled = 0, n_valid = 0, n_exact = 0;

  for (;;) {
    unsigned long long b = 0;
    if (wl == 0) b = atomicAdd(&A.counters[CNT_WORK_CURSOR_IDX], (unsigned long long)GRAB);
    b = __shfl_sync(0xffffffffu, b, 0);
    if ((long long)b >= NP) break;
    const long long q_end = min((long long)b + GRAB, NP);
    for (long long q0 = (long long)b; q0 < q_end; q0 += NG) {
      const long long pair = q0 + gi;
      const bool alive = pair < q_end;
      // ================= init (ConvexCell ctor, convex_cell.cu:116-214) =================================
      int t = 0, seed_id = 0, list_len = 0;
      float4 seed = make_float4(0, 0, 0, 0);
      unsigned hf4 = 0;
      unsigned long long e6 = 0;
      const int* list = nullptr;
      bool cull_ok = false, flagged = false;
      float pm_tet = 0.f, pm_bis = 0.f;
      int nb_v = 4, nb_p = 4, nb_e = 6, status = ST_success;
      if (alive) {
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
          const float4 pl = A.tet_geo[(size_t)t * 8 + lane];
          const float4 c = A.tet_geo[(size_t)t * 8 + 4 + lane];
          S.plane[lane] = pl;
          S.c0[lane] = c;
          S.cof[lane] = c;
          S.pnb[lane] = lane == 0 ? fid.x : (lane == 1 ? fid.y : (lane == 2 ? fid.z : fid.w));
          const unsigned char w = (unsigned char)((vadj4 >> (8 * lane)) & 0xffu);
          S.ver[lane] = lane == 0 ? make_uchar4(1, 3, 2, w)
                                  : (lane == 1 ? make_uchar4(0, 2, 3, w)
                                               : (lane == 2 ? make_uchar4(0, 3, 1, w) : make_uchar4(0, 1, 2, w)));
          ok0 = (c.w < 0.f) && isfinite(c.x) && isfinite(c.y) && isfinite(c.z) && isfinite(c.w);
          pm_tet = fmaxf(fabsf(pl.x), fmaxf(fabsf(pl.y), fabsf(pl.z)));
        }
        // edges (2,3)(1,3)(1,2)(0,3)(0,2)(0,1) with the e_adj of vertex pairs (0,1)(0,2)(0,3)(1,2)(1,3)(2,3)
        for (int q = lane; q < 6; q += G) {
          const unsigned char ea = q < 3 ? (q == 0 ? 2 : 1) : 0;
          const unsigned char eb = q == 0 ? 3 : (q == 1 ? 3 : (q == 2 ? 2 : (q == 3 ? 3 : (q == 4 ? 2 : 1))));
          S.edge[3 * q + 0] = ea;
          S.edge[3 * q + 1] = eb;
          S.edge[3 * q + 2] = (unsigned char)((e6 >> (8 * q)) & 0xff);
        }
        pm_tet = fmaxf(pm_tet, __shfl_xor_sync(gmask, pm_tet, 1));
        pm_tet = fmaxf(pm_tet, __shfl_xor_sync(gmask, pm_tet, 2));
        pm_tet = __shfl_sync(gmask, pm_tet, src);
        cull_ok = group_ballot<G>(gmask, gshift, !ok0) == 0 && !A.no_cull;
        if (PT) {
          const int tl = A.pair_local[pair];
          list = A.nbr + (size_t)tl * A.nbr_stride;
          list_len = A.nbr_cnt[tl];
        } else {
          list = A.nbr + (size_t)seed_id * A.nbr_stride;
          list_len = A.nbr_stride;
        }
      }
      __syncwarp();
      // ================= scan batches of G candidates, clip by the survivors =============================
      const int len_max = __reduce_max_sync(0xffffffffu, list_len);
      for (int base = 0; base < len_max; base += G) {
        unsigned todo = 0;
        int nb = -1;
        float4 eqn = make_float4(0, 0, 0, 0);
        if (alive && status == ST_success && base < list_len) {
          const int j = base + lane;
          nb = (j < list_len) ? list[j] : -1;
          // PT: the list is the tet's candidate set, skip the seed.  !PT: the first -1 terminates the list (:1259)
          bool cand = PT ? (nb >= 0 && nb != seed_id) : (nb >= 0);
          const bool is_end = !PT && (j < list_len) && (nb == -1);
          const bool valid_nb = cand;
          if (cand) {
            eqn = bisector_exact(seed, A.site4[nb]);
            if (cull_ok) {
              const float n1 = fabsf(eqn.x) + fabsf(eqn.y) + fabsf(eqn.z);
              const float nmax = fmaxf(fabsf(eqn.x), fmaxf(fabsf(eqn.y), fabsf(eqn.z)));
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
                cand = false;
                n_culled++;
              }
            }
          }
          todo = group_ballot<G>(gmask, gshift, cand);
          unsigned valid = group_ballot<G>(gmask, gshift, valid_nb);
          if (!PT) {
            const unsigned end_mask = group_ballot<G>(gmask, gshift, is_end);
            if (end_mask) {
              const unsigned before = (1u << (__ffs(end_mask) - 1)) - 1u;
              todo &= before;
              valid &= before;
              list_len = 0;  // nothing beyond the terminator
            }
          }
          // a plane beyond the compact cap: let the full-caps pass decide (it applies the reference's 64-plane rule)
          if (nb_p + __popc(todo) > KP || (nb_p >= KP && valid)) {
            status = ST_vertex_overflow;
            todo = 0;
          }
        }
        // ---- clip loop: one plane per group and iteration, all groups together ------------------------
        while (__ballot_sync(0xffffffffu, todo != 0)) {
          const bool c_act = todo != 0;
          const int k = c_act ? (__ffs(todo) - 1) : 0;
          const int nbk = __shfl_sync(0xffffffffu, nb, src + k);
          float4 e;
          e.x = __shfl_sync(0xffffffffu, eqn.x, src + k);
          e.y = __shfl_sync(0xffffffffu, eqn.y, src + k);
          e.z = __shfl_sync(0xffffffffu, eqn.z, src + k);
          e.w = __shfl_sync(0xffffffffu, eqn.w, src + k);
          if (c_act) {
            todo &= todo - 1;
            if (lane == 0) n_clips++;
          }
          // ---- C1: conflict flags; removed vertices -> slot list + directed dual edges in the bit matrix
          unsigned f0 = 0;
          int nb_r = 0;
          {
            const int vmax = __reduce_max_sync(0xffffffffu, c_act ? nb_v : 0);
            const float nmax = fmaxf(fabsf(e.x), fmaxf(fabsf(e.y), fabsf(e.z)));
            const float m_b = fmaxf(pm_bis, nmax);
            const float eps_b = filter_eps_upper(m_b), eps_t = filter_eps_upper(fmaxf(pm_tet, m_b));
            if (c_act) pm_bis = m_b;
            for (int vb = 0; vb < vmax; vb += G) {
              const int v = vb + lane;
              bool cf = false;
              uchar4 tv = make_uchar4(0, 0, 0, 0);
              if (c_act && v < nb_v) {
                tv = S.ver[v];
                const float4 c = S.cof[v];
                const float cw = c.w * e.w;
                const float s = fmaf(c.x, e.x, fmaf(c.y, e.y, fmaf(c.z, e.z, cw)));
                const float T = fmaf(fabsf(c.x) + fabsf(c.y) + fabsf(c.z), nmax, fabsf(cw));
                const float eps_up = min(tv.x, min(tv.y, tv.z)) < 4 ? eps_t : eps_b;
                if (fabsf(s) > fmaxf(4e-6f * T, fmaf(4e-7f, T, eps_up))) {
                  cf = s > 0.f;
                } else {
                  const int r2 = conflict_exact_flag_ool(S.plane[tv.x], S.plane[tv.y], S.plane[tv.z], e);
                  cf = (r2 & 1) != 0;
                  flagged = flagged || (r2 & 2) != 0;
                  n_exact++;
                }
              }
              const unsigned m = (__ballot_sync(0xffffffffu, cf) >> gshift) & glow;
              if (PT && cf) {
                S.rm[nb_r + __popc(m & ((1u << lane) - 1u))] = (unsigned char)v;
                atomicOr(&S.adj[tv.x], 1u << tv.y);
                atomicOr(&S.adj[tv.y], 1u << tv.z);
                atomicOr(&S.adj[tv.z], 1u << tv.x);
              }
              nb_r += __popc(m);
              f0 |= m << vb;
            }
          }
          // 0: nothing removed: the plane is dropped (:741-744)   1: clip   2: everything removed (:746-749)
          const int act2 = !c_act ? 0 : (nb_r == nb_v ? 2 : (nb_r != 0 ? 1 : 0));
          const bool a1 = act2 == 1;
          const int cur_p = nb_p;
          if (a1) {
            if (lane == 0) {
              S.plane[cur_p] = e;
              S.pnb[cur_p] = nbk;
            }
            if (lane < 2) reinterpret_cast<uint4*>(S.bnext)[lane] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
          }
          __syncwarp();
          int L = 0, st2 = ST_success;
          if (!PT) {
            // ---- C2 (given-neighbours mode): one lane per group replays the reference's serial bookkeeping
            if (a1 && lane == 0) {
              // swap partition, convex_cell.cu:706-721 (the cofactor cache follows the vertices)
              int nv = nb_v, i = 0;
              unsigned f = f0;
              while (i < nv) {
                if ((f >> i) & 1u) {
                  nv--;
                  const unsigned fn = (f >> nv) & 1u;
                  const uchar4 tmp = S.ver[i];
                  S.ver[i] = S.ver[nv];
                  S.ver[nv] = tmp;
                  S.cof[i] = S.cof[nv];
                  f = (f & ~(1u << i)) | (fn << i);
                } else
                  i++;
              }
              // cavity boundary, compute_boundary convex_cell.cu:618-678
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
                  if (++fails >= r) {  // a full round without progress: the reference spins until nb_iter > 65535
                    st2 = ST_inconsistent_boundary;
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
          } else {
          // ---- C2: cavity boundary = directed edges a->b of removed triangles whose twin b->a is absent
          const int rmax = __reduce_max_sync(0xffffffffu, c_act ? nb_r : 0);
          int nbnd = 0, first = MBK_END;
          for (int rb = 0; rb < rmax; rb += G) {
            const int r = rb + lane;
            if (a1 && r < nb_r) {
              const uchar4 tv = S.ver[S.rm[r]];
              const unsigned char pl[3] = {tv.x, tv.y, tv.z};
#pragma unroll
              for (int q = 0; q < 3; q++) {
                const int a = pl[q], bb = pl[(q + 1) % 3];
                if (!((S.adj[bb] >> a) & 1u)) {
                  S.bnext[a] = (unsigned char)bb;
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
          for (int rb = 0; rb < rmax; rb += G) {  // the matrix is all-zero between clips
            const int r = rb + lane;
            if (c_act && r < nb_r) {
              const uchar4 tv = S.ver[S.rm[r]];
              S.adj[tv.x] = 0u;
              S.adj[tv.y] = 0u;
              S.adj[tv.z] = 0u;
            }
          }
          // the cycle, from its smallest plane id; a boundary that is not ONE simple cycle is the reference's
          // inconsistent_boundary (:626-629)
          if (a1 && lane == 0 && first != MBK_END) {
            int cir = first;
            do {
              S.cyc[L++] = (unsigned char)cir;
              cir = S.bnext[cir];
            } while (cir != first && cir != MBK_END && L < nbnd && L < KP);
            if (cir != first || L != nbnd) st2 = ST_inconsistent_boundary;
          }
          }
          __syncwarp();
          L = __shfl_sync(0xffffffffu, L, src);
          st2 = __shfl_sync(0xffffffffu, st2, src);
          bool do_new = false;
          int nv_after = nb_v;
          if (act2 == 2) {
            status = ST_no_intersection;
            todo = 0;
          } else if (a1) {
            nb_p++;
            nv_after = nb_v - nb_r + L;
            if (st2 != ST_success)
              status = st2;
            else if (L != 0) {
              if (nb_e + L > KE)
                status = ST_edge_overflow;
              else if (nv_after + 1 > KT)
                status = ST_triangle_overflow;
              else
                do_new = true;
            }
          }
          // ---- C3: new edges (:756-762) and vertices (:764-773); vertex jj fills the jj-th hole, then the tail
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
                // PT: the jj-th hole, then the tail.  !PT: the partition has compacted the survivors: append
                const int slot = PT ? (jj < nb_r ? (int)S.rm[jj] : nb_v + (jj - nb_r)) : (nb_v - nb_r) + jj;
                S.ver[slot] = make_uchar4(cur_p, cir, nxt, w);
                S.cof[slot] = cofactors_f32(minors_exact(e, p2, p3));
                // is_vertex_perturb (:274-316): w-component of the vertex == 0
                const float wdet = det3_exact(e.x, e.y, e.z, p2.x, p2.y, p2.z, p3.x, p3.y, p3.z);
                pj = (wdet == 0.f);
              }
              if (__ballot_sync(0xffffffffu, pj) & gmask) perturb = true;
            }
            if (do_new) {
              nb_e += L;
              if (perturb) status = ST_needs_perturb;
              if (PT && L < nb_r && lane == 0) {
                // more vertices removed than created: move live tail vertices into the holes rm[L .. nb_r) that lie
                // below the new count (holes ascending, sources descending: they never cross)
                int srcv = nb_v - 1;
                for (int h = L; h < nb_r; h++) {
                  const int hole = S.rm[h];
                  if (hole >= nv_after) break;
                  // last live slot: not one of the remaining holes (which are exactly the removed slots >= hole)
                  while ((f0 >> srcv) & 1u) srcv--;
                  S.ver[hole] = S.ver[srcv];
                  S.cof[hole] = S.cof[srcv];
                  srcv--;
                }
              }
              nb_v = nv_after;
            }
          }
          if (a1 && status != ST_success) todo = 0;
          __syncwarp();
        }
      }
      // ================= write the record (copy(), convex_cell.cu:933-949) ================================
      if (alive) {
        const unsigned fbit = group_ballot<G>(gmask, gshift, flagged) ? MB_FLAG_BIT : 0u;
        long long blob_at = -1;
        int words = 0;
        const bool redo = status == ST_triangle_overflow || status == ST_vertex_overflow || status == ST_edge_overflow;
        if (status == ST_success) {
          words = compact_words(nb_v, nb_p, nb_e);
          unsigned long long at = 0;
          if (lane == 0) {
            if ((unsigned)words > chunk_left) {
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
            const uint32_t tail_mask = (3 * nb_e) & 3 ? (0xffffffffu >> (8 * (4 - ((3 * nb_e) & 3)))) : 0xffffffffu;
            for (int i = lane; i < ew; i += G) o[i] = (i == ew - 1) ? (se[i] & tail_mask) : se[i];
          }
        }
        if (lane == 0) {
          if (redo) {
            const unsigned long long at = atomicAdd(&A.counters[CNT_REDO], 1ull);
            A.redo_out[at] = (int)pair;
            A.pair_status[pair] = (signed char)ST_early_return;
            A.pair_blob[pair] = -1;
            A.pair_words[pair] = 0;
          } else {
            A.pair_status[pair] = (signed char)status;
            A.pair_blob[pair] = blob_at;
            A.pair_words[pair] = (int)(((blob_at >= 0) ? (unsigned)(words | (nb_p << 16)) : 0u) | fbit);
            if (status == ST_success && blob_at >= 0) n_valid++;
            if (fbit) {
              atomicAdd(&A.counters[CNT_FLAG_PAIRS], 1ull);
              if (status == ST_success && blob_at >= 0) atomicAdd(&A.counters[CNT_FLAG_CELLS], 1ull);
            }
            if (status != ST_success) atomicAdd(&blk_cnt[CNT_HIST + status + 1], 1ull);
          }
        }
      }
      __syncwarp();
    }
  }
  atomicAdd(&blk_cnt[CNT_CULLED], (unsigned long long)n_culled);
  atomicAdd(&blk_cnt[CNT_EXACT], (unsigned long long)n_exact);
  if (lane == 0) {
    atomicAdd(&blk_cnt[CNT_CLIPS], (unsigned long long)n_clips);
    atomicAdd(&blk_cnt[CNT_VALID], (unsigned long long)n_valid);
    atomicAdd(&blk_cnt[CNT_HIST + ST_success + 1], (unsigned long long)n_valid);
  }
  __syncthreads();
  if (threadIdx.x >= 1 && threadIdx.x < 16 && blk_cnt[threadIdx.x]) atomicAdd(&A.counters[threadIdx.x], blk_cnt[threadIdx.x]);
}
