// K6: power-cell topology summary on the device -- the part of update_power_cells that follows
// get_all_voro_info in every MATTopo iteration (reference src/rpd3d_base/rpd_update.cxx:303-316, 439-521,
// 568-639) and what check_cc_and_euler consumes (src/matfun_fix/fix_topo.cxx:81-144, 150-235, 310-380):
//
//   cell_neighbors   two cells of ONE power cell are neighbours iff they share a tet-face id
//                    (tfid_to_cells, rpd_update.cxx:303-316; first and last cell of the id's set)
//   cc_cells         connected components of a power cell's cells        (update_pc_cc_info :497-503)
//   facet_cc_cells   per half-plane (site, neigh): components of the cells that carry that facet, walking
//                    cell_neighbors inside the set only (get_CC_given_neighbors, common_cxx.h:447-489;
//                    update_pc_facet_cc_info :439-470)
//   euler            sum over the power cell's cells, ascending cell id, of cal_cell_euler (double sum of the
//                    float per-cell values), minus the number of cells (is_to_fix_voro_euler fix_topo.cxx:117-144)
//
// The reference does this per sphere with std::map<int, std::set<int>> and a BFS.  Here: ONE radix sort of the
// K4 facets by (kind, site, key) puts the tet-face facets of a power cell with equal face id next to each other
// (= the neighbour pairs) and groups the half-plane facets by (site, neigh); components are found with a
// lock-free union-find (compare-and-swap the larger root under the smaller, then resolve), so every label is the
// SMALLEST member of its component -- deterministic.
#include <cub/cub.cuh>

#include "mb_internal.h"

namespace {

__global__ void k_topo_keys(const uint32_t* __restrict__ blob, const long long* __restrict__ cell_off,
                            const int* __restrict__ f_cell, const int* __restrict__ f_key,
                            const unsigned char* __restrict__ f_istet, long n_facets,
                            unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  const long f = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_facets) return;
  const int c = f_cell[f];
  const unsigned site = blob[cell_off[c] / 4 + 1];
  const unsigned long long kind = f_istet[f] ? 0ull : 1ull;
  keys[f] = (kind << 63) | ((unsigned long long)site << 32) | (unsigned)f_key[f];
  vals[f] = (int)f;
}

__global__ void k_iota(int* __restrict__ p, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int)i;
}

// Lock-free union-find: roots are only ever changed by a compare-and-swap that still sees them as roots, the
// larger root goes under the smaller one, and the structure is read with volatile loads (L2, never a stale L1
// line of another SM's update).  No path compression while unions are in flight; labels are resolved by a
// separate kernel into a separate array.  Components are small (the cells of one power cell), so are the trees.
__device__ __forceinline__ int uf_find(const int* parent, int x) {
  const volatile int* vp = parent;
  int p;
  while ((p = vp[x]) != x) x = p;
  return x;
}

__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  for (;;) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a > b) {
      const int t = a;
      a = b;
      b = t;
    }
    if (atomicCAS(&parent[b], b, a) == b) return;  // b was still a root: hooked under the smaller root a
  }
}

// sorted tet-face part [0, n_tet_facets): a run of equal (site, face id) is tfid_to_cells[face id] of that power
// cell, ascending cell id; its first and last cell become neighbours (rpd_update.cxx:307-316)
__global__ void k_topo_cell_pairs(const unsigned long long* __restrict__ keys, const int* __restrict__ facet_of,
                                  const int* __restrict__ f_cell, long n_tet_facets, int* __restrict__ parent_cell,
                                  int* __restrict__ adj_a, int* __restrict__ adj_b) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tet_facets) return;
  adj_a[i] = -1;
  adj_b[i] = -1;
  const unsigned long long k = keys[i];
  if (i > 0 && keys[i - 1] == k) return;  // not a run start
  long e = i;
  while (e + 1 < n_tet_facets && keys[e + 1] == k) e++;
  if (e == i) return;  // boundary face of the power cell: no neighbour
  const int c1 = f_cell[facet_of[i]], c2 = f_cell[facet_of[e]];
  if (c1 == c2) return;
  adj_a[i] = c1;
  adj_b[i] = c2;
  uf_union(parent_cell, c1, c2);
}

__global__ void k_topo_labels(const int* __restrict__ parent, long n, const unsigned char* __restrict__ skip_if_tet,
                              int* __restrict__ label) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  label[i] = (skip_if_tet && skip_if_tet[i]) ? -1 : uf_find(parent, (int)i);
}

// first facet of every cell (K4 emits the facets cell by cell)
__global__ void k_topo_facet_begin(const int* __restrict__ f_cell, long n_facets, long n_cells, int* __restrict__ begin) {
  const long f = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_facets) return;
  const int c = f_cell[f];
  if (f == 0 || f_cell[f - 1] != c) begin[c] = (int)f;
  if (f == n_facets - 1) begin[n_cells] = (int)n_facets;
}

// cells without facets cannot occur (a valid cell has >= 4 active planes); fill gaps defensively
__global__ void k_topo_facet_begin_fix(int* __restrict__ begin, long n_cells) {
  // serial back-fill by one thread per 1024 cells would race; n_cells is small enough for a single pass
  const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  if (begin[c] < 0) {
    long d = c + 1;
    while (d < n_cells && begin[d] < 0) d++;
    begin[c] = begin[d];  // begin[n_cells] is always set
  }
}

// for every neighbour pair (c1, c2): half-plane facets with the same neighbour site are connected
__global__ void k_topo_facet_pairs(const int* __restrict__ adj_a, const int* __restrict__ adj_b, long n_tet_facets,
                                   const int* __restrict__ begin, const int* __restrict__ f_key,
                                   const unsigned char* __restrict__ f_istet, int* __restrict__ parent_facet) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tet_facets) return;
  const int c1 = adj_a[i];
  if (c1 < 0) return;
  const int c2 = adj_b[i];
  const int b1 = begin[c1], e1 = begin[c1 + 1], b2 = begin[c2], e2 = begin[c2 + 1];
  for (int f1 = b1; f1 < e1; f1++) {
    if (f_istet[f1]) continue;
    const int n = f_key[f1];
    for (int f2 = b2; f2 < e2; f2++)
      if (!f_istet[f2] && f_key[f2] == n) {
        uf_union(parent_facet, f1, f2);
        break;
      }
  }
}

// edges (both planes half-planes, key = sorted neighbour pair): for every neighbour pair (c1, c2) the edges with the
// same key are connected -- edge_cc_cells of update_pc_edge_cc_info (rpd_update.cxx:507-521)
__global__ void k_topo_edge_pairs(const int* __restrict__ adj_a, const int* __restrict__ adj_b, long n_tet_facets,
                                  const int* __restrict__ ebegin, const int* __restrict__ e_key2,
                                  int* __restrict__ parent_edge) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tet_facets) return;
  const int c1 = adj_a[i];
  if (c1 < 0) return;
  const int c2 = adj_b[i];
  const int b1 = ebegin[c1], e1 = ebegin[c1 + 1], b2 = ebegin[c2], e2 = ebegin[c2 + 1];
  for (int x = b1; x < e1; x++) {
    const int k0 = e_key2[2 * x], k1 = e_key2[2 * x + 1];
    for (int y = b2; y < e2; y++)
      if (e_key2[2 * y] == k0 && e_key2[2 * y + 1] == k1) {
        uf_union(parent_edge, x, y);
        break;
      }
  }
}

// first edge of every cell (K4 emits the edges cell by cell; a cell may have none)
__global__ void k_topo_edge_begin(const int* __restrict__ e_cell, long n_edges, long n_cells, int* __restrict__ begin) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int c = e_cell[e];
  if (e == 0 || e_cell[e - 1] != c) begin[c] = (int)e;
  if (e == n_edges - 1) begin[n_cells] = (int)n_edges;
}

__global__ void k_topo_cell_sites(const uint32_t* __restrict__ blob, const long long* __restrict__ cell_off, long n_cells,
                                  int* __restrict__ site, int* __restrict__ idx) {
  const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cells) return;
  site[c] = (int)blob[cell_off[c] / 4 + 1];
  idx[c] = (int)c;
}

// one thread per site over its cells in ascending cell id (stable sort): count, components, Euler sum
__global__ void k_topo_site_stats(const int* __restrict__ sorted_site, const int* __restrict__ sorted_cell, long n_cells,
                                  int n_site, const int* __restrict__ cell_cc, const float* __restrict__ c_euler,
                                  int* __restrict__ site_n_cells, int* __restrict__ site_n_cc,
                                  double* __restrict__ site_euler_sum) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_site) return;
  long lo = 0, hi = n_cells;  // lower bound of s
  while (lo < hi) {
    const long mid = (lo + hi) >> 1;
    if (sorted_site[mid] < s) lo = mid + 1; else hi = mid;
  }
  int n = 0, ncc = 0;
  double sum = 0.0;
  for (long i = lo; i < n_cells && sorted_site[i] == s; i++) {
    const int c = sorted_cell[i];
    n++;
    ncc += (cell_cc[c] == c);
    sum += (double)c_euler[c];  // msphere.euler_sum += convex_cell.euler (fix_topo.cxx:128)
  }
  site_n_cells[s] = n;
  site_n_cc[s] = ncc;
  site_euler_sum[s] = sum;
}

// sorted half-plane part: a run of equal (site, neigh) = facet_neigh_to_cells[neigh] of that power cell
__global__ void k_topo_pair_flags(const unsigned long long* __restrict__ keys, long first, long n, int* __restrict__ flag) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == n) {
    flag[i] = 0;
    return;
  }
  flag[i] = (i == 0 || keys[first + i - 1] != keys[first + i]) ? 1 : 0;
}

__global__ void k_topo_pair_write(const unsigned long long* __restrict__ keys, const int* __restrict__ facet_of, long first,
                                  long n, const int* __restrict__ flag, const int* __restrict__ pos,
                                  const int* __restrict__ facet_cc, int* __restrict__ pair_site,
                                  int* __restrict__ pair_neigh, int* __restrict__ pair_n_cc) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  const unsigned long long k = keys[first + i];
  int ncc = 0;
  for (long j = i; j < n && keys[first + j] == k; j++) {
    const int f = facet_of[first + j];
    ncc += (facet_cc[f] == f);
  }
  const int o = pos[i];
  pair_site[o] = (int)((k >> 32) & 0x7fffffffu);
  pair_neigh[o] = (int)(k & 0xffffffffu);
  pair_n_cc[o] = ncc;
}

__global__ void k_topo_count_tet(const unsigned long long* __restrict__ keys, long n, int* __restrict__ out) {
  // number of keys with the top bit clear = size of the tet-face part (binary search by one thread)
  long lo = 0, hi = n;
  while (lo < hi) {
    const long mid = (lo + hi) >> 1;
    if (keys[mid] >> 63) hi = mid; else lo = mid + 1;
  }
  *out = (int)lo;
}

}  // namespace

static inline unsigned nblk(long n, int b) { return (unsigned)((n + b - 1) / b); }

void rpd_topology(mb_ctx* ctx, mb_rpd_result* res) {
  cudaStream_t s = ctx->stream;
  const long nc = res->n_cells, nf = res->emit_counts.n_facets;
  const int n_site = res->n_site;
  res->topo_done = false;  // set only once the work below has succeeded
  res->topo_pairs = 0;
  res->t_site_n_cells.reserve((size_t)n_site + 1);
  res->t_site_n_cc.reserve((size_t)n_site + 1);
  res->t_site_euler.reserve((size_t)n_site + 1);
  MB_CUDA(cudaMemsetAsync(res->t_site_n_cells.p, 0, sizeof(int) * (size_t)n_site, s));
  MB_CUDA(cudaMemsetAsync(res->t_site_n_cc.p, 0, sizeof(int) * (size_t)n_site, s));
  MB_CUDA(cudaMemsetAsync(res->t_site_euler.p, 0, sizeof(double) * (size_t)n_site, s));
  if (nc == 0 || nf == 0) {
    MB_CUDA(cudaStreamSynchronize(s));
    res->topo_done = true;
    return;
  }
  DevBuf<unsigned long long> k_in, k_out;
  DevBuf<int> v_in, v_out, adj_a, adj_b, begin, c_site, c_idx, cs_site, cs_idx, flag, pos, scal, par_c, par_f;
  k_in.reserve(nf); k_out.reserve(nf); v_in.reserve(nf); v_out.reserve(nf);
  res->t_cell_cc.reserve(nc + 1);
  res->t_facet_cc.reserve(nf + 1);
  // ---- one sort: (kind, site, key) -> facet -----------------------------------------------------------------
  ctx->n_launches++;
  k_topo_keys<<<nblk(nf, 256), 256, 0, s>>>(res->blob.p, res->cell_off.p, res->f_cell.p, res->f_key.p, res->f_istet.p, nf,
                                            k_in.p, v_in.p);
  {
    size_t tmp = 0;
    MB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in.p, k_out.p, v_in.p, v_out.p, (int)nf, 0, 64, s));
    ctx->cub_tmp.reserve(tmp);
    ctx->n_launches += 9;  // histogram + 8 onesweep passes
    MB_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp, k_in.p, k_out.p, v_in.p, v_out.p, (int)nf, 0, 64, s));
  }
  scal.reserve(4);
  ctx->n_launches++;
  k_topo_count_tet<<<1, 1, 0, s>>>(k_out.p, nf, scal.p);
  int n_tetf = 0;
  MB_CUDA(cudaMemcpyAsync(&n_tetf, scal.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  // ---- cells: neighbour pairs + components -------------------------------------------------------------------
  ctx->n_launches += 2;
  par_c.reserve(nc + 1);
  par_f.reserve(nf + 1);
  k_iota<<<nblk(nc, 256), 256, 0, s>>>(par_c.p, nc);
  k_iota<<<nblk(nf, 256), 256, 0, s>>>(par_f.p, nf);
  MB_CUDA(cudaStreamSynchronize(s));
  const long ntf = n_tetf, nhp = nf - n_tetf;
  adj_a.reserve(ntf + 1); adj_b.reserve(ntf + 1); begin.reserve(nc + 2);
  MB_CUDA(cudaMemsetAsync(begin.p, 0xff, sizeof(int) * (size_t)(nc + 1), s));
  if (ntf > 0) {
    ctx->n_launches++;
    k_topo_cell_pairs<<<nblk(ntf, 256), 256, 0, s>>>(k_out.p, v_out.p, res->f_cell.p, ntf, par_c.p, adj_a.p, adj_b.p);
  }
  ctx->n_launches += 3;
  k_topo_labels<<<nblk(nc, 256), 256, 0, s>>>(par_c.p, nc, nullptr, res->t_cell_cc.p);
  k_topo_facet_begin<<<nblk(nf, 256), 256, 0, s>>>(res->f_cell.p, nf, nc, begin.p);
  k_topo_facet_begin_fix<<<nblk(nc, 256), 256, 0, s>>>(begin.p, nc);
  // ---- half-plane facets: components inside every (site, neigh) set --------------------------------------------
  if (ntf > 0) {
    ctx->n_launches++;
    k_topo_facet_pairs<<<nblk(ntf, 256), 256, 0, s>>>(adj_a.p, adj_b.p, ntf, begin.p, res->f_key.p, res->f_istet.p,
                                                     par_f.p);
  }
  ctx->n_launches++;
  k_topo_labels<<<nblk(nf, 256), 256, 0, s>>>(par_f.p, nf, res->f_istet.p, res->t_facet_cc.p);
  // ---- edges between two half-planes: components inside every (site, neigh_min, neigh_max) set ----------------
  const long ne = res->emit_counts.n_edges;
  res->t_edge_cc.reserve((size_t)ne + 1);
  if (ne > 0) {
    DevBuf<int> par_e, ebegin;
    par_e.reserve(ne + 1);
    ebegin.reserve(nc + 2);
    MB_CUDA(cudaMemsetAsync(ebegin.p, 0xff, sizeof(int) * (size_t)(nc + 1), s));
    ctx->n_launches += 3;
    k_iota<<<nblk(ne, 256), 256, 0, s>>>(par_e.p, ne);
    k_topo_edge_begin<<<nblk(ne, 256), 256, 0, s>>>(res->e_cell.p, ne, nc, ebegin.p);
    k_topo_facet_begin_fix<<<nblk(nc, 256), 256, 0, s>>>(ebegin.p, nc);  // cells without edges: empty range
    if (ntf > 0) {
      ctx->n_launches++;
      k_topo_edge_pairs<<<nblk(ntf, 256), 256, 0, s>>>(adj_a.p, adj_b.p, ntf, ebegin.p, res->e_key2.p, par_e.p);
    }
    ctx->n_launches++;
    k_topo_labels<<<nblk(ne, 256), 256, 0, s>>>(par_e.p, ne, nullptr, res->t_edge_cc.p);
    MB_CUDA(cudaStreamSynchronize(s));
    par_e.release();
    ebegin.release();
  }
  // ---- per-site statistics ---------------------------------------------------------------------------------------
  c_site.reserve(nc); c_idx.reserve(nc); cs_site.reserve(nc); cs_idx.reserve(nc);
  ctx->n_launches++;
  k_topo_cell_sites<<<nblk(nc, 256), 256, 0, s>>>(res->blob.p, res->cell_off.p, nc, c_site.p, c_idx.p);
  {
    size_t tmp = 0;
    MB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, c_site.p, cs_site.p, c_idx.p, cs_idx.p, (int)nc, 0, 32, s));
    ctx->cub_tmp.reserve(tmp);
    ctx->n_launches += 5;
    MB_CUDA(cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tmp, c_site.p, cs_site.p, c_idx.p, cs_idx.p, (int)nc, 0, 32, s));
  }
  ctx->n_launches++;
  k_topo_site_stats<<<nblk(n_site, 128), 128, 0, s>>>(cs_site.p, cs_idx.p, nc, n_site, res->t_cell_cc.p, res->c_euler.p,
                                                      res->t_site_n_cells.p, res->t_site_n_cc.p, res->t_site_euler.p);
  // ---- one entry per half-plane (site, neigh): number of facet components -----------------------------------------
  if (nhp > 0) {
    flag.reserve(nhp + 1); pos.reserve(nhp + 1);
    ctx->n_launches++;
    k_topo_pair_flags<<<nblk(nhp + 1, 256), 256, 0, s>>>(k_out.p, ntf, nhp, flag.p);
    {
      size_t tmp = 0;
      MB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, flag.p, pos.p, (int)(nhp + 1), s));
      ctx->cub_tmp.reserve(tmp);
      ctx->n_launches += 2;
      MB_CUDA(cub::DeviceScan::ExclusiveSum(ctx->cub_tmp.p, tmp, flag.p, pos.p, (int)(nhp + 1), s));
    }
    int n_pairs = 0;
    MB_CUDA(cudaMemcpyAsync(&n_pairs, pos.p + nhp, sizeof(int), cudaMemcpyDeviceToHost, s));
    MB_CUDA(cudaStreamSynchronize(s));
    res->topo_pairs = n_pairs;
    res->t_pair_site.reserve((size_t)n_pairs + 1);
    res->t_pair_neigh.reserve((size_t)n_pairs + 1);
    res->t_pair_ncc.reserve((size_t)n_pairs + 1);
    ctx->n_launches++;
    k_topo_pair_write<<<nblk(nhp, 256), 256, 0, s>>>(k_out.p, v_out.p, ntf, nhp, flag.p, pos.p, res->t_facet_cc.p,
                                                    res->t_pair_site.p, res->t_pair_neigh.p, res->t_pair_ncc.p);
  }
  MB_CUDA(cudaGetLastError());
  MB_CUDA(cudaStreamSynchronize(s));
  res->topo_done = true;
}
