// TEST INFRASTRUCTURE ONLY.  Stand-in for src/inputs/input_types.h + include/common_geogram.h (both need geogram
// v1.7.5, an un-vendored download of the reference: cmake/rpdDownloadExternal.cmake:17-48, absent here): just the
// declarations that src/rpd3d_base/rpd_update.cxx uses, so that file compiles IN PLACE (oracle/ref_shim_update.cpp).
// Semantics that matter for the comparison: a 3-vector of doubles built from floats (common_geogram.h:20,36), the
// facet count and facet adjacency of the surface mesh, GEO::parallel_for as a plain loop.
#pragma once
#include <array>
#include <cstddef>
#include <map>
#include <set>
#include <utility>
#include <vector>

#include "common.h"      // the reference's own (aint2 ..., no third-party dependency)
#include "common_cxx.h"  // the reference's own (cfloat4, get_CC_given_neighbors, set_intersection, to_set ...)

namespace GEO {
typedef unsigned int index_t;
const index_t NO_FACET = index_t(-1);
struct vec3 {
  double x, y, z;
  vec3() : x(0), y(0), z(0) {}
  vec3(double a, double b, double c) : x(a), y(b), z(c) {}
  double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  const double& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <class F>
inline void parallel_for(std::size_t from, std::size_t to, F f) {
  for (std::size_t i = from; i < to; i++) f((int)i);
}
}  // namespace GEO
typedef GEO::vec3 Vector3;
typedef std::pair<Vector3, int> v2int;
inline Vector3 to_vec(const cfloat4& v) { return Vector3(v.x, v.y, v.z); }

enum EdgeType { UE = -1, SE = 1, CE = 2 };  // src/inputs/input_types.h:13-17

// the part of SurfaceMesh (src/inputs/input_types.h:383-475, a GEO::Mesh) that rpd_update.cxx reads
struct SurfaceMesh {
  struct Facets {
    std::vector<int> adj;  // 3 neighbours per triangle, -1 = none
    GEO::index_t nb() const { return (GEO::index_t)(adj.size() / 3); }
    GEO::index_t nb_vertices(GEO::index_t) const { return 3; }
    GEO::index_t adjacent(GEO::index_t f, GEO::index_t le) const {
      const int a = adj[3 * (std::size_t)f + le];
      return a < 0 ? GEO::NO_FACET : (GEO::index_t)a;
    }
  } facets;
  std::set<aint2> fe_sf_fs_pairs;
  std::map<int, std::set<int>> sf_fid_neighs_no_cross;
};
