// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// The reference's power-cell update compiled IN PLACE: src/rpd3d_base/rpd_update.cxx (get_all_voro_info :80-340,
// update_pc_cc_info :497-503, update_pc_facet_cc_info :439-470, update_pc_edge_cc_info :507-521, driven by
// update_power_cells :568-639) on top of the reference's ConvexCellHost (voronoi_defs.cxx: reload_active,
// reload_pc_explicit, cal_cell_euler).  geogram / matbase / inputs headers are replaced by the stand-ins in
// oracle/stubs_update (a 3-vector, the PowerCell containers, facet count + adjacency of the surface mesh).
// This is what pins K4 (k_emit) and K6 (rpd_topo.cu) -- and the C restatement orc_emit / oracle.topology -- against
// the reference's own code instead of against a restatement of it.
#include <cstdio>
#include <cstring>
#include <vector>

#include "voronoi_defs.cxx"  // -I/root/reference/src/rpd3d_base
#define printf(...) ((void)0)  // "updating power cells ..." chatter
#include "rpd_update.cxx"
#undef printf

#include "oracle.h"

static void expand(const orc_record& r, ConvexCellHost& h, int id) {
  h.is_active = true;
  h.status = (Status)r.status;
  h.thread_id = r.thread_id;
  h.voro_id = r.voro_id;
  h.tet_id = r.tet_id;
  h.euler = r.euler;
  h.weight = r.weight;
  h.nb_v = r.nb_v;
  h.nb_p = r.nb_p;
  h.nb_e = r.nb_e;
  for (int i = 0; i < r.nb_v; i++)
    h.ver_data_trans[i] = cmake_uchar4(r.ver[i][0], r.ver[i][1], r.ver[i][2], r.ver[i][3]);
  for (int i = 0; i < r.nb_p; i++) {
    h.clip_data_trans[i] = cmake_float5(r.clip[i].x, r.clip[i].y, r.clip[i].z, r.clip[i].w, r.clip[i].h);
    h.clip_id2_data_trans[i] = cmake_int2(r.id2[i][0], r.id2[i][1]);
  }
  for (int i = 0; i < r.nb_e; i++) h.edge_data[i] = cmake_uchar3(r.edge[i][0], r.edge[i][1], r.edge[i][2]);
  h.id = id;
}

static std::vector<MedialSphere> g_spheres;
static std::vector<ConvexCellHost> g_cells;

template <class T>
static long dump(const std::vector<T>& v, T* out, long cap) {
  if (out && (long)v.size() <= cap) memcpy(out, v.data(), sizeof(T) * v.size());
  return (long)v.size();
}

extern "C" {

// Run update_power_cells on n success records (ConvexCellTransfer layout, id = index).  fe_map: n_fe rows of
// (tet, lf_min, lf_max, fe_type, fe_id, fe_line_id) = TetMesh::tet_es2fe_map.  Results stay in static storage until
// the next call; fetch them with the ref_update_* getters below.  Returns 0.
int ref_update_run(const orc_record* recs, long n, int n_site, int max_surf_fid, const int* fe_map, long n_fe) {
  g_cells.assign((size_t)n, ConvexCellHost());
  for (long i = 0; i < n; i++) expand(recs[i], g_cells[(size_t)i], (int)i);
  g_spheres.assign((size_t)n_site, MedialSphere());
  for (int s = 0; s < n_site; s++) g_spheres[(size_t)s].id = s;
  SurfaceMesh sf;
  sf.facets.adj.assign(3 * (size_t)(max_surf_fid + 1), -1);  // facet count = max_surf_fid + 1 (rpd_update.cxx:606)
  std::map<aint3, aint3> es2fe;
  for (long i = 0; i < n_fe; i++)
    es2fe[{{fe_map[6 * i], fe_map[6 * i + 1], fe_map[6 * i + 2]}}] = {{fe_map[6 * i + 3], fe_map[6 * i + 4], fe_map[6 * i + 5]}};
  update_power_cells(sf, g_cells, g_spheres, es2fe, false);
  return 0;
}

// per-cell Euler value as the reference computes it (ConvexCellHost::cal_cell_euler after reload_active)
void ref_update_cell_euler(float* out) {
  for (size_t i = 0; i < g_cells.size(); i++) out[i] = (float)g_cells[i].cal_cell_euler();
}

// Every getter flattens one PowerCell container over all spheres in ascending sphere id and returns the number of
// rows (call with out = NULL / cap = 0 to size).  Row layouts:
//   facets      (site, neigh, cell)                       facet_neigh_to_cells
//   tfids       (site, cell, tet-face id)                 cell_to_tfids
//   surf        (site, cell, surf_fid) + centroid xyz     cell_to_surfv2fid (list order kept)
//   vertices    (site, cell, lvid, k0, k1, k2, surf_fid) + pos xyz      vertex_2id / vertex_2pos
//   edges       (site, n_min, n_max, cell, lv_a, lv_b)    edge_2endvertices (list order kept)
//   e2cells     (site, n_min, n_max, cell)                e_to_cells
//   neighbours  (site, cell, neighbour cell)              cell_neighbors
//   cc          (site, component index, cell)             cc_cells
//   facet_cc    (site, neigh, component index, cell)      facet_cc_cells (sorted by size, :447-451)
//   edge_cc     (site, n_min, n_max, component index, cell)             edge_cc_cells
//   fe          (site, kind 1=SE 2=CE, cell, lv1, lv2, fe_line_id, fe_id)   se_covered_lvids / ce_covered_lvids
//   fe_end      (site, cell, lvid, neigh, se_line_id) + pos xyz         se_line_endpos
long ref_update_get(const char* what, int* rows, double* vals, long cap) {
  std::vector<int> r;
  std::vector<double> v;
  const std::string w(what);
  int width = 0, vw = 0;
  for (const MedialSphere& ms : g_spheres) {
    const PowerCell& pc = ms.pcell;
    const int s = ms.id;
    if (w == "facets") {
      width = 3;
      for (const auto& kv : pc.facet_neigh_to_cells)
        for (int c : kv.second) r.insert(r.end(), {s, kv.first, c});
    } else if (w == "tfids") {
      width = 3;
      for (const auto& kv : pc.cell_to_tfids)
        for (int f : kv.second) r.insert(r.end(), {s, kv.first, f});
    } else if (w == "surf") {
      width = 3;
      vw = 3;
      for (const auto& kv : pc.cell_to_surfv2fid)
        for (const v2int& p : kv.second) {
          r.insert(r.end(), {s, kv.first, p.second});
          v.insert(v.end(), {p.first.x, p.first.y, p.first.z});
        }
    } else if (w == "vertices") {
      width = 7;
      vw = 3;
      for (const auto& kv : pc.vertex_2id) {
        const v2int& p = pc.vertex_2pos.at(kv.first);
        r.insert(r.end(), {s, kv.first[0], kv.first[1], kv.second[0], kv.second[1], kv.second[2], p.second});
        v.insert(v.end(), {p.first.x, p.first.y, p.first.z});
      }
    } else if (w == "edges") {
      width = 6;
      for (const auto& kv : pc.edge_2endvertices)
        for (const auto& e : kv.second) r.insert(r.end(), {s, kv.first[0], kv.first[1], e[0][0], e[0][1], e[1][1]});
    } else if (w == "e2cells") {
      width = 4;
      for (const auto& kv : pc.e_to_cells)
        for (int c : kv.second) r.insert(r.end(), {s, kv.first[0], kv.first[1], c});
    } else if (w == "neighbours") {
      width = 3;
      for (const auto& kv : pc.cell_neighbors)
        for (int c : kv.second) r.insert(r.end(), {s, kv.first, c});
    } else if (w == "cc") {
      width = 3;
      for (size_t k = 0; k < pc.cc_cells.size(); k++)
        for (int c : pc.cc_cells[k]) r.insert(r.end(), {s, (int)k, c});
    } else if (w == "facet_cc") {
      width = 4;
      for (const auto& kv : pc.facet_cc_cells)
        for (size_t k = 0; k < kv.second.size(); k++)
          for (int c : kv.second[k]) r.insert(r.end(), {s, kv.first, (int)k, c});
    } else if (w == "edge_cc") {
      width = 5;
      for (const auto& kv : pc.edge_cc_cells)
        for (size_t k = 0; k < kv.second.size(); k++)
          for (int c : kv.second[k]) r.insert(r.end(), {s, kv.first[0], kv.first[1], (int)k, c});
    } else if (w == "fe") {
      width = 7;
      for (const aint5& a : pc.se_covered_lvids) r.insert(r.end(), {s, 1, a[0], a[1], a[2], a[3], a[4]});
      for (const aint5& a : pc.ce_covered_lvids) r.insert(r.end(), {s, 2, a[0], a[1], a[2], a[3], a[4]});
    } else if (w == "fe_end") {
      width = 5;
      vw = 3;
      for (const auto& kv : pc.se_line_endpos) {
        r.insert(r.end(), {s, kv.first[0], kv.first[1], kv.first[2], kv.first[3]});
        v.insert(v.end(), {kv.second.x, kv.second.y, kv.second.z});
      }
    } else {
      return -1;
    }
  }
  const long n = width ? (long)r.size() / width : 0;
  if (rows && n <= cap) memcpy(rows, r.data(), sizeof(int) * r.size());
  if (vals && vw && n <= cap) memcpy(vals, v.data(), sizeof(double) * v.size());
  return n;
}

}  // extern "C"
