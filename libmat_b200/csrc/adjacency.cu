// f4: sparse counterpart of load_tet_adj_info (reference src/IO/IO_CXX/io.cxx:238-335), host code.
//
// The reference builds, per mesh load, std::map<int, std::set<int>> vertex -> tets and intersects those sets for
// every tet edge and face; its edge table e_adjs is the dense triangular matrix n_vert (n_vert + 1) / 2 + 1
// (io.cxx:264): 2.6 GB at 36 k vertices, an int overflow at 46 341.  Same outputs here by sorting edge / face keys:
//   v_adjs[v]      number of tets at vertex v                                   (:256-260)
//   e_adj6[t][e]   number of tets around tet edge e, e in the (0,1)(0,2)(0,3)(1,2)(1,3)(2,3) local-vertex order of
//                  convex_cell.cu:194-207 -- the 6 entries of the dense table a tet's cells read   (:262-280)
//   f_adjs[t][i]   1 (boundary) or 2 (interior) for face i in tet_faces_lvid order ({2,1,3},{0,2,3},{1,0,3},{0,1,2},
//                  convex_cell.h:30-31)                                                    (:299-304)
//   f_ids[t][i]    boundary face: its surface-mesh facet id (given by the caller in (tet, local face) scan order, or
//                  0 .. n_b-1 in that order); interior face: n_sf_facets + rank of the face's first visit in
//                  (tet, local face) scan order                                                    (:286-326)
#include <algorithm>
#include <cstdint>
#include <vector>

#include "mb_internal.h"

namespace {
struct Key3 {
  int a, b, c, slot;
  bool operator<(const Key3& o) const {
    if (a != o.a) return a < o.a;
    if (b != o.b) return b < o.b;
    if (c != o.c) return c < o.c;
    return slot < o.slot;
  }
};
}  // namespace

void tet_adjacency(const int* idx, int n_tet, int n_vert, const int* boundary_sf_fids, int n_sf_facets, int* v_adjs,
                   int* e_adj6, int* f_adjs, int* f_ids, int* n_boundary) {
  static const int ep[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
  static const int fl[4][3] = {{2, 1, 3}, {0, 2, 3}, {1, 0, 3}, {0, 1, 2}};
  for (long i = 0; i < 4L * n_tet; i++)
    MB_REQUIRE(idx[i] >= 0 && idx[i] < n_vert, MB_ERR_ARG, "tet index out of range");
  if (v_adjs) {
    std::fill(v_adjs, v_adjs + n_vert, 0);
    for (long i = 0; i < 4L * n_tet; i++) v_adjs[idx[i]]++;
  }
  if (e_adj6) {
    std::vector<std::pair<uint64_t, int>> ek((size_t)6 * n_tet);
#pragma omp parallel for schedule(static)
    for (long t = 0; t < n_tet; t++)
      for (int e = 0; e < 6; e++) {
        const uint64_t a = (uint64_t)idx[4 * t + ep[e][0]], b = (uint64_t)idx[4 * t + ep[e][1]];
        ek[(size_t)6 * t + e] = {std::min(a, b) * (uint64_t)n_vert + std::max(a, b), (int)(6 * t + e)};
      }
    std::sort(ek.begin(), ek.end());
    for (size_t i = 0; i < ek.size();) {
      size_t j = i;
      while (j < ek.size() && ek[j].first == ek[i].first) j++;
      for (size_t q = i; q < j; q++) e_adj6[ek[q].second] = (int)(j - i);
      i = j;
    }
  }
  if (f_adjs || f_ids) {
    std::vector<Key3> fk((size_t)4 * n_tet);
#pragma omp parallel for schedule(static)
    for (long t = 0; t < n_tet; t++)
      for (int i = 0; i < 4; i++) {
        int v[3] = {idx[4 * t + fl[i][0]], idx[4 * t + fl[i][1]], idx[4 * t + fl[i][2]]};
        std::sort(v, v + 3);
        fk[(size_t)4 * t + i] = {v[0], v[1], v[2], (int)(4 * t + i)};
      }
    std::sort(fk.begin(), fk.end());
    // per face slot: multiplicity and the smallest slot of its run (= the face's first visit in scan order)
    std::vector<int> mult((size_t)4 * n_tet), first((size_t)4 * n_tet);
    for (size_t i = 0; i < fk.size();) {
      size_t j = i;
      while (j < fk.size() && fk[j].a == fk[i].a && fk[j].b == fk[i].b && fk[j].c == fk[i].c) j++;
      MB_REQUIRE(j - i <= 2, MB_ERR_ARG, "a face shared by more than two tets: not a manifold tet mesh (io.cxx:302)");
      for (size_t q = i; q < j; q++) {
        mult[(size_t)fk[q].slot] = (int)(j - i);
        first[(size_t)fk[q].slot] = fk[i].slot;  // slots ascend inside a run
      }
      i = j;
    }
    int nb = 0, next_interior = n_sf_facets;
    std::vector<int> id_of_first((size_t)4 * n_tet, -1);
    for (long s = 0; s < 4L * n_tet; s++) {
      if (f_adjs) f_adjs[s] = mult[(size_t)s];
      int id;
      if (mult[(size_t)s] == 1) {
        id = boundary_sf_fids ? boundary_sf_fids[nb] : nb;
        nb++;
      } else if (first[(size_t)s] == (int)s) {
        id = next_interior++;
        id_of_first[(size_t)s] = id;
      } else {
        id = id_of_first[(size_t)first[(size_t)s]];
      }
      if (f_ids) f_ids[s] = id;
    }
    if (n_boundary) *n_boundary = nb;
  }
}
