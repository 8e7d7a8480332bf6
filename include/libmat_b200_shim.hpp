// libmat_b200 -- C++ shim with the reference's exact host signature for the RPD path.
//
// Header-only; compile it INSIDE LibMAT (it includes LibMAT's own voronoi_defs.h, so the
// ConvexCellHost it fills is the caller's type, whatever its std::vector ABI) and link
// libmat_b200.so.  It replaces the body of
//     compute_clipped_voro_diagram_GPU      reference src/rpd3d/voronoi.cu:455-795
//                                           (declaration src/rpd3d/voronoi.h:52-61)
// and is what RPD3D_GPU::calculate / calculate_partial call (src/rpd3d_api/rpd_api.cxx:117-123,
// 275-281).  Everything below the signature goes through the C ABI of include/libmat_b200.h.
//
// Behaviour kept: returns the valid cells sorted by (tet, site) with ConvexCellHost::id = index
// (voronoi.cu:744-769) filled by the copy_cc rule (:433-449: is_active = true, euler/weight/
// counts/arrays copied, cell_vol left at its default); never throws; site_cell_vol is resized and
// zero-filled like the reference leaves it (:501-502, never written back).  v2tets,
// num_itr_global, nb_Lloyd_iter and preferred_tet_k are accepted and unused, as in the reference.
// Behaviour changed on purpose: no process exit() on CUDA errors (an empty vector is returned and
// the message is printed to stderr), no record.csv appended in the CWD (:571-579), no device 0
// hard-wiring (MB_DEVICE environment variable, default: current device).
#pragma once

#if defined(__linux__)
#include <sys/mman.h>
#endif

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <vector>

#include "libmat_b200.h"
#include "voronoi_defs.h"  // LibMAT's own header (src/rpd3d_base/voronoi_defs.h)

#include "libmat_b200_shim_ctx.hpp"

namespace libmat_b200 {

// the caller's ConvexCellHost must have the capacities and element sizes the records were produced for
// (src/rpd3d_base/voronoi_defs.h:71-99, voronoi_common.h _MAX_P_/_MAX_T_/_MAX_E_, common_cxx.h:23-43)
static_assert(sizeof(cuchar4) == 4, "cuchar4 must be 4 packed bytes (vertices are copied as 32-bit words)");
static_assert(sizeof(ConvexCellHost::ver_data_trans) == MB_MAX_T * sizeof(cuchar4), "_MAX_T_ != 96");
static_assert(sizeof(ConvexCellHost::clip_data_trans) / sizeof(cfloat5) == MB_MAX_P, "_MAX_P_ != 64");
static_assert(sizeof(ConvexCellHost::clip_id2_data_trans) / sizeof(cint2) == MB_MAX_P, "_MAX_P_ != 64");
static_assert(sizeof(ConvexCellHost::edge_data) / sizeof(cuchar3) == MB_MAX_E, "_MAX_E_ != 152");

// compact record (see DESIGN.md "compact cell record") -> ConvexCellHost, the copy_cc rule
inline void expand_cell(const uint32_t* w, int id, ConvexCellHost& c) {
  const int nb_v = w[2] & 0xff, nb_p = (w[2] >> 8) & 0xff, nb_e = (w[2] >> 16) & 0xff;
  c.is_active = true;
  c.status = static_cast<Status>((int)((w[2] >> 24) & 0xf));  // bits 29 / 31: transport format, bit 30: flagged class
  c.thread_id = id;
  c.voro_id = (int)w[1];
  c.tet_id = (int)w[0];
  c.euler = -1.f;
  std::memcpy(&c.weight, &w[3], 4);
  c.nb_v = (uchar)nb_v;
  c.nb_p = (uchar)nb_p;
  c.nb_e = (uchar)nb_e;
  const uint32_t* p = w + 4;
  std::memcpy(c.ver_data_trans, p, 4 * (size_t)nb_v);
  p += nb_v;
  const float* pl = reinterpret_cast<const float*>(p);
  p += 4 * nb_p;
  for (int i = 0; i < nb_p; i++) {
    float h;
    std::memcpy(&h, &p[3 * i + 2], 4);
    c.clip_data_trans[i] = cmake_float5(pl[4 * i], pl[4 * i + 1], pl[4 * i + 2], pl[4 * i + 3], h);
    c.clip_id2_data_trans[i] = cmake_int2((int)p[3 * i], (int)p[3 * i + 1]);
  }
  p += 3 * nb_p;
  // the compact record packs edges as 3 bytes; the host cuchar3 is aligned(4) (common_cxx.h:29-31)
  const unsigned char* eb = reinterpret_cast<const unsigned char*>(p);
  for (int i = 0; i < nb_e; i++) c.edge_data[i] = cmake_uchar3(eb[3 * i], eb[3 * i + 1], eb[3 * i + 2]);
  c.id = id;
}

// wall milliseconds of the stages of this thread's last compute_clipped_voro_diagram_GPU call:
// [0] mb_set_tetmesh + mb_rpd_upload_sites (H2D)  [1] mb_rpd_run_to_host (kernels + streamed D2H)
// [2] std::vector<ConvexCellHost> construction + expansion of the compact records  [3] whole call
inline double* last_call_ms() {
  static thread_local double ms[4] = {0, 0, 0, 0};
  return ms;
}
inline double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace libmat_b200

#ifndef LIBMAT_B200_NO_REFERENCE_NAMES
inline std::vector<ConvexCellHost> compute_clipped_voro_diagram_GPU(
    const int num_itr_global, const std::vector<float>& vertices, const std::vector<int>& indices,
    const std::map<int, std::set<int>>& v2tets, const std::vector<int>& v_adjs,
    const std::vector<int>& e_adjs, const std::vector<int>& f_adjs, const std::vector<int>& f_ids,
    std::vector<float>& site, const int n_site, const std::vector<float>& site_weights,
    const std::vector<uint>& site_flags, const std::vector<int>& site_knn, const int site_k,
    std::vector<float>& site_cell_vol, const bool site_is_transposed, int nb_Lloyd_iter = 1,
    int preferred_tet_k = 0) {
  (void)num_itr_global; (void)v2tets; (void)site_is_transposed; (void)nb_Lloyd_iter; (void)preferred_tet_k;
  std::vector<ConvexCellHost> out;
  double* ms = libmat_b200::last_call_ms();
  const double t0 = libmat_b200::now_ms();
  site_cell_vol.assign((size_t)n_site, 0.f);  // voronoi.cu:501-502
  mb_ctx* ctx = libmat_b200::thread_ctx();
  if (!ctx) return out;
  const int n_vert = (int)(vertices.size() / 3), n_tet = (int)(indices.size() / 4);
  auto fail = [&](const char* what) {
    std::fprintf(stderr, "[libmat_b200] %s: %s\n", what, mb_last_error(ctx));
    return std::vector<ConvexCellHost>();
  };
  // the reference's dense triangular e_adjs (io.cxx:264) is accepted as given
  if (mb_set_tetmesh(ctx, vertices.data(), n_vert, indices.data(), n_tet, v_adjs.data(), e_adjs.data(), nullptr,
                     f_adjs.data(), f_ids.data()))
    return fail("mb_set_tetmesh");
  mb_rpd_result* res = nullptr;
  // Candidate (tet, site) pairs come from the uniform-grid search (8x faster than the dense relation
  // predicate of voronoi.cu:154-193); cells are still clipped by exactly the listed neighbours, so the
  // returned cells are identical whenever site_knn holds every true power neighbour -- the CGAL
  // regular-triangulation lists of rpd_api.cxx do.  MB_CANDIDATES=reference restores the dense predicate.
  mb_rpd_opts opts = {0, 0, 0, 1, 0};
  if (const char* cm = std::getenv("MB_CANDIDATES"))
    if (std::strcmp(cm, "reference") == 0) opts.grid_candidates = 0;
  // an empty site_knn selects the library's own uniform-grid neighbour search
  const int* knn = site_knn.empty() ? nullptr : site_knn.data();
  if (mb_rpd_upload_sites(ctx, site.data(), site_weights.data(), site_flags.data(), n_site, knn, site_k))
    return fail("mb_rpd_upload_sites");
  const double t1 = libmat_b200::now_ms();
  // streamed run: the compact records of tet span c cross PCIe while span c+1 is clipped; on return the whole
  // result sits in the library's pinned host memory (replaces the D2H of n_tet*tet_k ConvexCellTransfer
  // records + the std::map dedup of voronoi.cu:717-769)
  const void* blob_v = nullptr;
  const long* offs = nullptr;
  if (mb_rpd_run_to_host(ctx, &opts, 0, &res, &blob_v, &offs)) return fail("mb_rpd_run_to_host");
  // A per-tet candidate list of the grid search holds at most 96 sites; a tet with more TRUE candidates would lose
  // cells (mb_rpd_stats[4], never silent).  The drop-in call then falls back to the reference's own relation
  // predicate, which has no such cap.
  long st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  mb_rpd_stats(res, st);
  if (st[4] != 0 && opts.grid_candidates && knn) {
    std::fprintf(stderr, "[libmat_b200] %ld tets with more than 96 candidate sites: rerunning with the reference's "
                         "relation predicate (MB_CANDIDATES=reference)\n", st[4]);
    mb_rpd_free(res);
    res = nullptr;
    opts.grid_candidates = 0;
    if (mb_rpd_run_to_host(ctx, &opts, 0, &res, &blob_v, &offs)) return fail("mb_rpd_run_to_host");
  } else if (st[4] != 0) {
    std::fprintf(stderr, "[libmat_b200] WARNING: %ld tets with more than 96 candidate sites were truncated\n", st[4]);
  }
  const double t2 = libmat_b200::now_ms();
  long n_cells = 0;
  mb_rpd_count(res, &n_cells, nullptr, nullptr);
  const uint32_t* blob = static_cast<const uint32_t*>(blob_v);
  // 3 776 B per ConvexCellHost: the vector is GBs of fresh pages (config 2: 2.9 GB) whose first touch -- one page fault
  // and one kernel-side zero fill each -- would otherwise happen serially inside resize().  Touch them from all
  // threads first (transparent huge pages where the kernel grants them), then construct.
  out.reserve((size_t)n_cells);
  if (n_cells > 4096) {
    char* base = reinterpret_cast<char*>(out.data());
    const size_t bytes = (size_t)n_cells * sizeof(ConvexCellHost);
#if defined(__linux__) && defined(MADV_HUGEPAGE)
    {
      const size_t a = ((size_t)base + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
      if (a < (size_t)base + bytes) madvise(reinterpret_cast<void*>(a), (((size_t)base + bytes - a) >> 21) << 21, MADV_HUGEPAGE);
    }
#endif
#pragma omp parallel for schedule(static)
    for (long off = 0; off < (long)bytes; off += 4096) base[off] = 0;
  }
  out.resize((size_t)n_cells);
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n_cells; i++) libmat_b200::expand_cell(blob + offs[i] / 4, (int)i, out[(size_t)i]);
  mb_rpd_free(res);
  const double t3 = libmat_b200::now_ms();
  ms[0] = t1 - t0;
  ms[1] = t2 - t1;
  ms[2] = t3 - t2;
  ms[3] = t3 - t0;
  return out;
}
#endif
