// TEST INFRASTRUCTURE ONLY.  Stand-in for src/matbase/medial_sphere.h (needs geogram): the containers of PowerCell
// (:145-277) that src/rpd3d_base/rpd_update.cxx fills, with the reference's element types, and the few MedialSphere
// members that file touches.  topo_clear restates medial_sphere.cxx:732-762 as "start from an empty power cell".
#pragma once
#include "input_types.h"

enum SphereType { T_UNK = -1, T_2 = 2, T_3_MORE = 3 };  // the two values rpd_update.cxx compares (:279-297)

struct PowerCell {
  int voro_id = -1;
  std::set<int> cell_ids, tet_ids;
  std::map<int, std::set<int>> cell_neighbors, cell_to_tfids, facet_neigh_to_cells;
  std::map<int, std::vector<v2int>> cell_to_surfv2fid;
  std::vector<std::set<int>> cc_cells;
  std::vector<std::vector<v2int>> cc_surf_v2fids, surf_v2fid_in_groups;
  std::map<int, std::vector<std::set<int>>> facet_cc_cells;
  std::map<int, std::vector<std::vector<v2int>>> facet_cc_surf_v2fids;
  std::map<aint2, std::set<int>> e_to_cells;
  std::map<aint2, std::vector<std::set<int>>> edge_cc_cells;
  std::map<aint2, std::vector<std::array<aint2, 2>>> edge_2endvertices;
  std::map<aint2, aint3> vertex_2id;
  std::map<aint2, v2int> vertex_2pos;
  std::set<aint5> se_covered_lvids, ce_covered_lvids;
  std::map<aint4, Vector3> se_line_endpos;
};

struct MedialSphere {
  int id = -1;
  bool is_deleted = false;
  int type = T_UNK;
  PowerCell pcell;
  void topo_clear() { pcell = PowerCell(); }
  void update_sphere_covered_sf_fids(const SurfaceMesh&, bool) {}  // bookkeeping outside the compared outputs
};
