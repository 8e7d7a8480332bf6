// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// Host build of the reference's own device clipping code.  The reference file
// src/rpd3d/convex_cell.cu is #included IN PLACE from /root/reference (found via -I, never
// copied) after a small macro shim that turns the CUDA qualifiers into no-ops, so that
// g++ compiles the reference's ConvexCell / clip_by_plane / compute_boundary / copy and the
// kernel body clipped_voro_cell_test_GPU_param_tet (convex_cell.cu:1166-1337) for the host.
// This is the "rpd3d_base ConvexCellHost CPU path" of BASELINE.md section 3: the reference has
// no other CPU clipper.  Built by oracle/Makefile into oracle/_ref/libref_rpd.so.
//
// Build flags: -O2 -DNDEBUG -ffp-contract=off (x86-64 SSE2: IEEE float/double, no FMA
// contraction) -- the arithmetic contract the product kernels reproduce bit-for-bit.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>  // host-side vector types (float4, uchar4, make_*) only

#undef __device__
#undef __global__
#undef __shared__
#undef __constant__
#undef __host__
#define __device__
#define __global__
#define __host__
#define __constant__ static const
#define __shared__ static thread_local
struct ShimDim3 {
  unsigned x = 0, y = 0, z = 0;
};
static thread_local ShimDim3 threadIdx, blockIdx, blockDim;
using std::max;
using std::min;
static inline int __float2int_rn(float f) { return (int)lrintf(f); }
static inline float atomicAdd(float* a, float v) {
  float o = *a;
  *a += v;
  return o;
}

#ifdef USE_ARITHMETIC_FILTER
// libref_rpd_filter.so: the same file with the reference's own static-filter branch switched on
// (voronoi_common.h:27-32 "Uncomment to activate arithmetic filters"; convex_cell.cu:479-497): a cell with a
// conflict |det| under the predicate_generator bound ends with status needs_exact_predicates -- the reference's
// definition of the flagged class.  Its print_info() chatter for such cells (:1275-1277) is silenced.
#define printf(...) ((void)0)
#endif
#include "convex_cell.cu"  // the reference file, unmodified (-I/root/reference/src/rpd3d)
#ifdef USE_ARITHMETIC_FILTER
#undef printf
#endif

#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {

int ref_rpd_record_bytes() { return (int)sizeof(ConvexCellTransfer); }

// offsets of the ConvexCellTransfer fields, so that the Python side can decode the records
// without hard-coding the layout (order: status thread_id voro_id tet_id weight is_active
// nb_v nb_p nb_e ver clip id2 edge euler cell_vol id)
void ref_rpd_record_layout(int* off) {
  ConvexCellTransfer* p = nullptr;
#define OFF(f) (int)(size_t)(&(p->f))
  off[0] = OFF(status);
  off[1] = OFF(thread_id);
  off[2] = OFF(voro_id);
  off[3] = OFF(tet_id);
  off[4] = OFF(weight);
  off[5] = OFF(is_active);
  off[6] = OFF(nb_v);
  off[7] = OFF(nb_p);
  off[8] = OFF(nb_e);
  off[9] = OFF(ver_data_trans);
  off[10] = OFF(clip_data_trans);
  off[11] = OFF(clip_id2_data_trans);
  off[12] = OFF(edge_data);
  off[13] = OFF(euler);
  off[14] = OFF(cell_vol);
  off[15] = OFF(id);
  off[16] = (int)sizeof(float5);
#undef OFF
}

// Run the reference kernel body once per (tet, site) pair.
//
// Every call uses a 4-vertex local mesh for that tet (probe host_shim3): the
// kernel reads nothing else about the mesh, so the record equals the full-mesh call's record
// once tet_id is patched.  The dense e_adjs table (get_edge_idx, convex_cell.h:46-66) is built
// locally (11 ints) from the compact per-tet-edge counts e_adj6, given in the
// (0,1)(0,2)(0,3)(1,2)(1,3)(2,3) local-vertex-pair order of convex_cell.cu:194-207.
//
// records : n_pairs * sizeof(ConvexCellTransfer); slots the kernel does not write keep
//           status = early_return (the reference leaves them uninitialised).
// stat    : n_pairs, the gpu_stat value (final *cc.status, incl. the post-copy
//           no_intersection for |vol| < 0.1, convex_cell.cu:1040).
// vol/bary: per-site accumulators (n_site, 3*n_site SoA), may be NULL.
// returns seconds spent in the pair loop (steady clock), or <0 on error.
double ref_rpd_run_pairs(const float* verts_aos, const int* idx_aos, int n_tet, const int* v_adjs,
                         const int* e_adj6, const int* f_adjs, const int* f_ids,
                         const float* site_soa, const float* site_w, const unsigned* site_flags,
                         int n_site, const int* site_knn, int site_k, const int* pair_tet,
                         const int* pair_site, long n_pairs, void* records, int* stat,
                         float* site_vol, float* site_bary, int n_threads) {
  ConvexCellTransfer* out = reinterpret_cast<ConvexCellTransfer*>(records);
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
  int nt = omp_get_max_threads();
#else
  int nt = 1;
#endif
  std::vector<std::vector<float>> vol_t(nt), bary_t(nt);
  for (int t = 0; t < nt; t++) {
    vol_t[t].assign(n_site, 0.f);
    bary_t[t].assign(3 * (size_t)n_site, 0.f);
  }
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
#pragma omp parallel for schedule(dynamic, 256)
  for (long p = 0; p < n_pairs; p++) {
#ifdef _OPENMP
    int th = omp_get_thread_num();
#else
    int th = 0;
#endif
    const int t = pair_tet[p];
    float vert[12];
    int lidx[4] = {0, 1, 2, 3};
    int lv_adjs[4];
    for (int l = 0; l < 4; l++) {
      int v = idx_aos[4 * t + l];
      vert[l] = verts_aos[3 * v];
      vert[4 + l] = verts_aos[3 * v + 1];
      vert[8 + l] = verts_aos[3 * v + 2];
      lv_adjs[l] = v_adjs[v];
    }
    int le_adjs[11];
    for (int i = 0; i < 11; i++) le_adjs[i] = -1;
    int e = 0;
    for (int a = 0; a < 4; a++)
      for (int b = a + 1; b < 4; b++) le_adjs[get_edge_idx(a, b, 4)] = e_adj6[6 * t + e++];
    int tet_knn[1] = {pair_site[p]};
    Status st = security_radius_not_reached;  // initial fill of gpu_stat, voronoi.cu:668
    out[p].status = early_return;
    blockDim.x = 1;
    threadIdx.x = 0;
    blockIdx.x = 0;
    clipped_voro_cell_test_GPU_param_tet(
        site_soa, n_site, (size_t)n_site, site_w, site_flags, site_knn, (size_t)n_site, site_k,
        vert, 4, 4, lidx, 1, 1, lv_adjs, le_adjs, f_adjs + 4 * t, f_ids + 4 * t, tet_knn, 1, 1, &st,
        nullptr, out + p, bary_t[th].data(), (size_t)n_site, vol_t[th].data());
    if (out[p].status != early_return) {
      out[p].tet_id = t;
      out[p].thread_id = (int)p;
    }
    if (stat) stat[p] = (int)st;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (site_vol)
    for (int s = 0; s < n_site; s++) {
      float a = 0;
      for (int t = 0; t < nt; t++) a += vol_t[t][s];
      site_vol[s] = a;
    }
  if (site_bary)
    for (size_t s = 0; s < 3 * (size_t)n_site; s++) {
      float a = 0;
      for (int t = 0; t < nt; t++) a += bary_t[t][s];
      site_bary[s] = a;
    }
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

int ref_rpd_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
