// K3: sub-warp-cooperative convex-cell clipping (one group of G lanes per (tet, site) cell).
//
// Replaces clipped_voro_cell_test_GPU_param_tet (reference src/rpd3d/convex_cell.cu:1166-1337),
// which runs ONE THREAD per cell with 3 KB of shared memory per thread (64 threads / SM).
// Here G in {8,16,32} lanes share one 2.3 KB polytope in shared memory:
//   * lanes scan the neighbour list G at a time: float4 gather of the neighbour sphere, exact
//     bisector, then a conservative FP32 filter against the 4 vertices of the initial tet (any
//     plane that provably removes nothing is skipped -- the reference would pop it again,
//     convex_cell.cu:741-744);
//   * surviving planes are handled in list order: lanes evaluate the reference's FP64 det4x4
//     predicate on the cell vertices in parallel (bit-exact), one lane replays the reference's
//     order-defining bookkeeping (swap partition :706-721, cavity boundary walk :618-678) on the
//     shared-memory state, then lanes create the new edges / vertices in parallel (:756-773).
// Array positions therefore equal the reference's, so records are byte-identical on the
// defined entries in given-neighbours mode.
#pragma once

#include "rpd_device.cuh"

struct ClipArgs {
  // mesh
  const float4* vert4;
  const int4* tet_idx;
  const int4* tet_fadj;
  const int4* tet_fid;
  const uint2* tet_e6;
  // sites
  const float4* site4;
  int n_site;
  // neighbour lists
  const int* nbr;        // given mode: [n_site][nbr_stride]; grid mode: [local tet][nbr_stride]
  int nbr_stride;        // given mode: site_k; grid mode: kcap
  const int* nbr_cnt;    // grid mode: #candidates per local tet (nullptr in given mode)
  int tet_first;         // first tet of the processed range (grid-mode lists are relative to it)
  // pairs
  const int* pair_tet;
  const int* pair_site;
  long long n_pairs;
  // outputs
  signed char* pair_status;
  long long* pair_blob;
  int* pair_words;
  uint32_t* scratch;
  unsigned long long scratch_words;
  unsigned long long* counters;  // RpdCounters as u64 array
};

// indices into RpdCounters viewed as u64[]
#define CNT_BLOB 0
#define CNT_CLIPS 1
#define CNT_CULLED 2
#define CNT_VALID 3
#define CNT_CANDOVF 4
#define CNT_HIST 5

template <int G>
__device__ __forceinline__ unsigned group_ballot(unsigned gmask, int gshift, bool pred) {
  unsigned m = __ballot_sync(gmask, pred);
  return (m >> gshift) & (G == 32 ? 0xffffffffu : ((1u << G) - 1u));
}

// z of the edge between planes a < b (closed form of what new_edge stored, convex_cell.cu:604-616,
// and of the 6 initial tet edges :194-210)
__device__ __forceinline__ unsigned char edge_z(int a, int b, const float* hf, unsigned long long e6) {
  if (b < 4) {
    // face pair -> index in e_adj6 order: (2,3)->0 (1,3)->1 (1,2)->2 (0,3)->3 (0,2)->4 (0,1)->5
    int idx = (a == 2) ? 0 : (a == 1 ? (b == 3 ? 1 : 2) : (b == 3 ? 3 : (b == 2 ? 4 : 5)));
    return (unsigned char)((e6 >> (8 * idx)) & 0xff);
  }
  const float ha = a == 0 ? hf[0] : (a == 1 ? hf[1] : (a == 2 ? hf[2] : (a == 3 ? hf[3] : 1.f)));
  return (unsigned char)fmaxf(ha, 1.f);
}

#define CLIP_CHUNK_WORDS 1024u  // scratch is bump-allocated per group in 4 KB chunks

template <int G>
__device__ void clip_cell(const ClipArgs& A, long long pair, CellS& S, int lane, unsigned gmask,
                          int gshift, unsigned long long& chunk_at, unsigned& chunk_left,
                          unsigned long long* blk_cnt) {
  const int t = A.pair_tet[pair];
  const int seed_id = A.pair_site[pair];
  const int src = gshift;  // warp lane of the group's rank 0

  // ---- load tet + seed (all lanes; same addresses -> one transaction, broadcast) -------------
  const int4 vi = A.tet_idx[t];
  const float4 q0 = A.vert4[vi.x], q1 = A.vert4[vi.y], q2 = A.vert4[vi.z], q3 = A.vert4[vi.w];
  const int4 fadj = A.tet_fadj[t];
  const int4 fid = A.tet_fid[t];
  const uint2 e6u = A.tet_e6[t];
  const unsigned long long e6 = ((unsigned long long)e6u.y << 32) | e6u.x;
  const float4 seed = A.site4[seed_id];
  float hf[4] = {(float)fadj.x, (float)fadj.y, (float)fadj.z, (float)fadj.w};

  // ---- initial polytope: ConvexCell ctor, convex_cell.cu:116-214 ----------------------------------
  if (lane < 4) {
    // face i is opposite local vertex i: {2,1,3},{0,2,3},{1,0,3},{0,1,2} (convex_cell.h:30-31)
    float3 p[4] = {make_float3(q0.x, q0.y, q0.z), make_float3(q1.x, q1.y, q1.z),
                   make_float3(q2.x, q2.y, q2.z), make_float3(q3.x, q3.y, q3.z)};
    const int f0 = lane == 0 ? 2 : (lane == 1 ? 0 : (lane == 2 ? 1 : 0));
    const int f1 = lane == 0 ? 1 : (lane == 1 ? 2 : (lane == 2 ? 0 : 1));
    const int f2 = lane == 3 ? 2 : 3;
    float3 a = f0 == 0 ? p[0] : (f0 == 1 ? p[1] : p[2]);
    float3 b = f1 == 0 ? p[0] : (f1 == 1 ? p[1] : p[2]);
    float3 c = f2 == 2 ? p[2] : p[3];
    S.plane[lane] = tri2plane_exact(a, b, c);
    S.pnb[lane] = lane == 0 ? fid.x : (lane == 1 ? fid.y : (lane == 2 ? fid.z : fid.w));
    // dual triangles (1,3,2) (0,2,3) (0,3,1) (0,1,2) with w = (uchar)v_adjs (:186-189)
    const float4 qq = lane == 0 ? q0 : (lane == 1 ? q1 : (lane == 2 ? q2 : q3));
    const unsigned char w = (unsigned char)__float_as_int(qq.w);
    S.ver[lane] = lane == 0 ? make_uchar4(1, 3, 2, w)
                            : (lane == 1 ? make_uchar4(0, 2, 3, w)
                                         : (lane == 2 ? make_uchar4(0, 3, 1, w) : make_uchar4(0, 1, 2, w)));
  }
  if (lane < 6) {
    // edges (2,3)(1,3)(1,2)(0,3)(0,2)(0,1) with the e_adj of vertex pairs (0,1)(0,2)(0,3)(1,2)(1,3)(2,3)
    const unsigned char ea = lane < 3 ? (lane == 0 ? 2 : 1) : 0;
    const unsigned char eb = lane == 0 ? 3 : (lane == 1 ? 3 : (lane == 2 ? 2 : (lane == 3 ? 3 : (lane == 4 ? 2 : 1))));
    S.edge[3 * lane + 0] = ea;
    S.edge[3 * lane + 1] = eb;
    S.edge[3 * lane + 2] = (unsigned char)((e6 >> (8 * lane)) & 0xff);
  }
  __syncwarp(gmask);

  // ---- filter data: cofactor vectors of the 4 initial vertices -------------------------------
  bool cull_ok = true;
  if (lane < 4) {
    const uchar4 v = S.ver[lane];
    const Minors m = minors_exact(S.plane[v.x], S.plane[v.y], S.plane[v.z]);
    // det4x4(p1,p2,p3,e) = m234*e.x - m134*e.y + m124*e.z - m123*e.w
    const float4 c = make_float4((float)m.m234, (float)(-m.m134), (float)m.m124, (float)(-m.m123));
    S.c0[lane] = c;
    S.a0[lane] = fabsf(c.x) + fabsf(c.y) + fabsf(c.z);
    // a proper vertex has c.w < 0 (conflict <=> det > 0 <=> vertex on the negative side)
    cull_ok = (c.w < 0.f) && isfinite(c.x) && isfinite(c.y) && isfinite(c.z) && isfinite(c.w);
  }
  cull_ok = group_ballot<G>(gmask, gshift, !cull_ok) == 0;
  __syncwarp(gmask);
  float4 c0[4] = {S.c0[0], S.c0[1], S.c0[2], S.c0[3]};
  float a0[4] = {S.a0[0], S.a0[1], S.a0[2], S.a0[3]};

  int nb_v = 4, nb_p = 4, nb_e = 6;
  int status = ST_success;
  unsigned long long n_clips = 0, n_culled = 0;

  // ---- neighbour list --------------------------------------------------------------------------
  const int* list;
  int list_len;
  if (A.nbr_cnt) {
    list = A.nbr + (size_t)(t - A.tet_first) * A.nbr_stride;
    list_len = A.nbr_cnt[t - A.tet_first];
  } else {
    list = A.nbr + (size_t)seed_id * A.nbr_stride;
    list_len = A.nbr_stride;
  }

  bool done = false;
  for (int base = 0; base < list_len && !done; base += G) {
    // ---- stage 1: G neighbours in parallel ---------------------------------------------------
    const int j = base + lane;
    int nb = (j < list_len) ? list[j] : -1;
    bool is_end = false, cand = false, valid_nb = false;
    float4 eqn = make_float4(0, 0, 0, 0);
    if (A.nbr_cnt) {
      // grid mode: the list is the tet's candidate set; skip the seed itself
      cand = (nb >= 0 && nb != seed_id);
      valid_nb = cand;
    } else {
      // given mode: the first -1 terminates the list (convex_cell.cu:1259)
      is_end = (j < list_len) && (nb == -1);
      cand = (nb >= 0);
      valid_nb = cand;
    }
    if (cand) {
      const float4 B = A.site4[nb];
      eqn = bisector_exact(seed, B);
      if (cull_ok) {
        const float n1 = fabsf(eqn.x) + fabsf(eqn.y) + fabsf(eqn.z);
        const float nmax = fmaxf(fabsf(eqn.x), fmaxf(fabsf(eqn.y), fabsf(eqn.z)));
        bool all_out = true;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const float cw = c0[i].w * eqn.w;
          const float s = fmaf(c0[i].x, eqn.x, fmaf(c0[i].y, eqn.y, fmaf(c0[i].z, eqn.z, cw)));
          const float T = fmaf(a0[i], nmax, fabsf(cw));
          const float margin = fmaxf(4e-6f * T, 1e-3f * n1 * fabsf(c0[i].w));
          all_out = all_out && (s < -margin);
        }
        if (all_out) {
          cand = false;  // certainly removes no vertex: the reference would pop this plane
          n_culled++;
        }
      }
    }
    unsigned end_mask = group_ballot<G>(gmask, gshift, is_end);
    unsigned todo = group_ballot<G>(gmask, gshift, cand);
    // every listed neighbour (culled or not) makes the reference call new_plane, which refuses
    // a 65th plane (vertex_overflow, convex_cell.cu:562-565)
    unsigned valid = group_ballot<G>(gmask, gshift, valid_nb);
    if (end_mask) {
      const unsigned before = (1u << (__ffs(end_mask) - 1)) - 1u;
      todo &= before;
      valid &= before;
      done = true;
    }
    if (nb_p >= MBK_MAX_P && valid) {
      status = ST_vertex_overflow;
      break;
    }
    // ---- stage 2: survivors in list order ----------------------------------------------------
    while (todo) {
      const int k = __ffs(todo) - 1;
      todo &= todo - 1;
      if (nb_p >= MBK_MAX_P) {
        status = ST_vertex_overflow;
        break;
      }
      const int nbk = __shfl_sync(gmask, nb, src + k);
      float4 e;
      e.x = __shfl_sync(gmask, eqn.x, src + k);
      e.y = __shfl_sync(gmask, eqn.y, src + k);
      e.z = __shfl_sync(gmask, eqn.z, src + k);
      e.w = __shfl_sync(gmask, eqn.w, src + k);
      n_clips++;
      // conflict flags of all vertices (lanes strided over vertices)
      unsigned long long f0 = 0;
      unsigned f1 = 0;
      int nb_r = 0;
      for (int vb = 0; vb < nb_v; vb += G) {
        const int v = vb + lane;
        bool cf = false;
        if (v < nb_v) {
          const uchar4 tv = S.ver[v];
          cf = conflict_exact(S.plane[tv.x], S.plane[tv.y], S.plane[tv.z], e);
        }
        const unsigned m = group_ballot<G>(gmask, gshift, cf);
        nb_r += __popc(m);
        if (vb < 64)  // G divides 64: a round never straddles the two words
          f0 |= (unsigned long long)m << vb;
        else
          f1 |= m << (vb - 64);
      }
      if (nb_r == 0) continue;  // plane removes nothing: dropped (:741-744)
      if (nb_r == nb_v) {
        status = ST_no_intersection;  // :746-749
        break;
      }
      // ---- one lane replays the order-defining serial bookkeeping ---------------------------
      int L = 0;
      int st2 = ST_success;
      if (lane == 0) {
        const int cur_p = nb_p;
        S.plane[cur_p] = e;
        S.pnb[cur_p] = nbk;
        // swap partition, convex_cell.cu:706-721
        int nv = nb_v, i = 0;
        while (i < nv) {
          const bool fi = i < 64 ? ((f0 >> i) & 1ull) : ((f1 >> (i - 64)) & 1u);
          if (fi) {
            nv--;
            const bool fn = nv < 64 ? ((f0 >> nv) & 1ull) : ((f1 >> (nv - 64)) & 1u);
            const uchar4 tmp = S.ver[i];
            S.ver[i] = S.ver[nv];
            S.ver[nv] = tmp;
            if (i < 64)
              f0 = (f0 & ~(1ull << i)) | ((unsigned long long)fn << i);
            else
              f1 = (f1 & ~(1u << (i - 64))) | ((unsigned)fn << (i - 64));
          } else
            i++;
        }
        // cavity boundary, compute_boundary convex_cell.cu:618-678
        for (int p = 0; p <= cur_p; p++) S.bnext[p] = MBK_END;
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
          } while (cir != first && cir != MBK_END && L < MBK_MAX_P);
        }
      }
      L = __shfl_sync(gmask, L, src);
      st2 = __shfl_sync(gmask, st2, src);
      __syncwarp(gmask);
      const int cur_p = nb_p;
      nb_p++;
      nb_v -= nb_r;
      if (st2 != ST_success) {
        status = st2;
        break;
      }
      if (L == 0) continue;  // first_boundary_ == END_OF_LIST (:754)
      // ---- new edges (:756-762) and new vertices (:764-773), lanes over the cycle -----------
      if (nb_e + L > MBK_MAX_E) {
        status = ST_edge_overflow;
        break;
      }
      bool perturb = false;
      for (int jb = 0; jb < L; jb += G) {
        const int jj = jb + lane;
        bool pj = false;
        if (jj < L) {
          const int cir = S.cyc[jj];
          const int nxt = S.cyc[jj + 1 == L ? 0 : jj + 1];
          const unsigned char z1 = edge_z(cir, cur_p, hf, e6);
          S.edge[3 * (nb_e + jj) + 0] = (unsigned char)cir;
          S.edge[3 * (nb_e + jj) + 1] = (unsigned char)cur_p;
          S.edge[3 * (nb_e + jj) + 2] = z1;
          const unsigned char z2 = edge_z(nxt, cur_p, hf, e6);
          const unsigned char z3 = edge_z(min(cir, nxt), max(cir, nxt), hf, e6);
          const unsigned char w = max(max(z1, z2), z3);
          if (nb_v + jj + 1 < MBK_MAX_T) S.ver[nb_v + jj] = make_uchar4(cur_p, cir, nxt, w);
          // is_vertex_perturb (:274-316): w-component of the vertex == 0
          const float4 p1 = e, p2 = S.plane[cir], p3 = S.plane[nxt];
          const float wdet = det3_exact(p1.x, p1.y, p1.z, p2.x, p2.y, p2.z, p3.x, p3.y, p3.z);
          pj = (wdet == 0.f);
        }
        const unsigned pm = group_ballot<G>(gmask, gshift, pj);
        if (pm && !perturb) {
          perturb = true;
          const int jp = jb + __ffs(pm) - 1;          // first perturbed vertex
          const int jo = MBK_MAX_T - 1 - nb_v;        // first overflowing vertex (if < L)
          status = (jo < L && jo < jp) ? ST_triangle_overflow : ST_needs_perturb;
        }
      }
      if (!perturb && nb_v + L + 1 > MBK_MAX_T) status = ST_triangle_overflow;  // nb_v+1 >= 96 (:525)
      nb_e += L;
      nb_v += L;
      __syncwarp(gmask);
      if (status != ST_success) break;
      // a later listed neighbour in this chunk would be refused by new_plane
      if (nb_p >= MBK_MAX_P && (valid & ~((2u << k) - 1u))) {
        status = ST_vertex_overflow;
        break;
      }
    }
    if (status != ST_success) break;
  }
  __syncwarp(gmask);

  // ---- output: copy() convex_cell.cu:933-949 into the compact record ---------------------------
  long long blob_at = -1;
  int words = 0;
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
        o[0] = (uint32_t)t;
        o[1] = (uint32_t)seed_id;
        o[2] = (uint32_t)nb_v | ((uint32_t)nb_p << 8) | ((uint32_t)nb_e << 16) | ((uint32_t)status << 24);
        o[3] = __float_as_uint(seed.w);
      }
      o += 4;
      const uint32_t* sv = reinterpret_cast<const uint32_t*>(S.ver);
      for (int i = lane; i < nb_v; i += G) o[i] = sv[i];
      o += nb_v;
      // planes are written word-wise: the record is only 4-byte aligned
      for (int i = lane; i < 4 * nb_p; i += G)
        o[i] = reinterpret_cast<const uint32_t*>(S.plane)[i];
      o += 4 * nb_p;
      for (int i = lane; i < nb_p; i += G) {
        int ida, idb;
        float h;
        if (i < 4) {
          ida = S.pnb[i];
          idb = -1;
          h = hf[i];
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
      for (int i = lane; i < ew; i += G) o[i] = se[i];
    }
  }
  if (lane == 0) {
    A.pair_status[pair] = (signed char)status;
    A.pair_blob[pair] = blob_at;
    A.pair_words[pair] = (blob_at >= 0) ? words : 0;
    atomicAdd(&blk_cnt[CNT_HIST + status + 1], 1ull);
    if (status == ST_success && blob_at >= 0) atomicAdd(&blk_cnt[CNT_VALID], 1ull);
  }
  // per-lane statistics
  for (int o = G / 2; o > 0; o >>= 1) {
    n_culled += __shfl_down_sync(gmask, n_culled, o, G);
  }
  if (lane == 0) {
    atomicAdd(&blk_cnt[CNT_CLIPS], n_clips);
    atomicAdd(&blk_cnt[CNT_CULLED], n_culled);
  }
  __syncwarp(gmask);
}

template <int G>
__global__ void __launch_bounds__(128) k_clip(ClipArgs A) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CellS* cells = reinterpret_cast<CellS*>(smem_raw);
  constexpr int GROUPS_PER_BLOCK = 128 / G;
  const int g_in_block = threadIdx.x / G;
  const int lane = threadIdx.x % G;
  const int wl = threadIdx.x & 31;
  const int gshift = (wl / G) * G;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gshift);
  CellS& S = cells[g_in_block];
  __shared__ unsigned long long blk_cnt[16];  // block-aggregated counters (RpdCounters layout)
  if (threadIdx.x < 16) blk_cnt[threadIdx.x] = 0;
  __syncthreads();
  unsigned long long chunk_at = 0;
  unsigned chunk_left = 0;
  const long long n_groups = (long long)gridDim.x * GROUPS_PER_BLOCK;
  for (long long pair = (long long)blockIdx.x * GROUPS_PER_BLOCK + g_in_block; pair < A.n_pairs;
       pair += n_groups) {
    clip_cell<G>(A, pair, S, lane, gmask, gshift, chunk_at, chunk_left, blk_cnt);
  }
  __syncthreads();
  if (threadIdx.x >= 1 && threadIdx.x < 16 && blk_cnt[threadIdx.x])
    atomicAdd(&A.counters[threadIdx.x], blk_cnt[threadIdx.x]);
}
