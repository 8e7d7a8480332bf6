// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// The reference's ConvexCellHost record and its CPU post-processing
// (src/rpd3d_base/voronoi_defs.cxx: reload_active :76-106, reload_pc_explicit :108-188,
// cal_cell_euler :190-222, compute_vertex_coordinates :33-74) compiled IN PLACE.
// Input records use the ConvexCellTransfer layout (3456 B, see oracle.h orc_record); they are
// expanded with the copy_cc rule of src/rpd3d/voronoi.cu:433-449 (restated here because
// voronoi.cu itself needs nvcc).
#include <cstring>

#include "voronoi_defs.cxx"  // found via -I/root/reference/src/rpd3d_base

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

extern "C" {

int ref_host_sizeof_cell() { return (int)sizeof(ConvexCellHost); }

// offsets of the POD tail of ConvexCellHost
void ref_host_layout(int* off) {
  ConvexCellHost* p = nullptr;
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
  off[16] = (int)sizeof(cfloat5);
#undef OFF
}

// reload_active + cal_cell_euler per record
void ref_host_reload_active(const orc_record* recs, long n, unsigned char* active_planes,
                            unsigned char* active_edges, float* euler) {
#pragma omp parallel for
  for (long i = 0; i < n; i++) {
    ConvexCellHost h;
    expand(recs[i], h, (int)i);
    h.reload_active();
    memset(active_planes + i * _MAX_P_, 0, _MAX_P_);
    memset(active_edges + i * _MAX_E_, 0, _MAX_E_);
    for (int p = 0; p < h.nb_p; p++) active_planes[i * _MAX_P_ + p] = h.active_clipping_planes[p] > 0;
    for (int e = 0; e < h.nb_e; e++) active_edges[i * _MAX_E_ + e] = h.active_edges[e] > 0;
    if (euler) euler[i] = (float)h.cal_cell_euler();
  }
}

// compute_vertex_coordinates per vertex: out n*96*4 floats (x/w, y/w, z/w, 1)
void ref_host_vertex_coordinates(const orc_record* recs, long n, float* out) {
#pragma omp parallel for
  for (long i = 0; i < n; i++) {
    ConvexCellHost h;
    expand(recs[i], h, (int)i);
    for (int t = 0; t < h.nb_v; t++) {
      cfloat4 v = h.compute_vertex_coordinates(cmake_uchar3(h.ver_trans(t)));
      float* o = out + (i * _MAX_T_ + t) * 4;
      o[0] = v.x;
      o[1] = v.y;
      o[2] = v.z;
      o[3] = v.w;
    }
  }
}

// reload_pc_explicit: face loops. loops: n*64*16 ints (-1 padded: up to 16 vertices per face kept),
// lf2active: n*64 ints
void ref_host_face_loops(const orc_record* recs, long n, int* loops, int* lf2active) {
#pragma omp parallel for
  for (long i = 0; i < n; i++) {
    ConvexCellHost h;
    expand(recs[i], h, (int)i);
    h.reload_pc_explicit();
    for (int p = 0; p < _MAX_P_; p++) {
      lf2active[i * _MAX_P_ + p] = p < h.nb_p ? h.pc_lf2active_map[p] : -1;
      for (int k = 0; k < 16; k++) loops[(i * _MAX_P_ + p) * 16 + k] = -1;
      if (p < h.nb_p && h.pc_lf2active_map[p] >= 0) {
        const auto& f = h.pc_local_active_faces[h.pc_lf2active_map[p]];
        for (size_t k = 0; k < f.size() && k < 16; k++) loops[(i * _MAX_P_ + p) * 16 + k] = f[k];
      }
    }
  }
}

}  // extern "C"
