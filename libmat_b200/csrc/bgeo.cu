// The IO_CUDA result format, written straight from the compact cell records: Houdini .bgeo V5, big-endian,
// one polygon per shown cell facet, point attribute P, primitive attribute "PrimAttr" = site id.
//
// Replaces, byte for byte (tests/test_bgeo.py compares with the reference's own writer compiled in place):
//   save_convex_cells_houdini          reference src/IO/IO_CUDA/io_cuda.cxx:152-187  (is_slice_plane = false)
//   get_one_convex_cell_faces_const    io_cuda.cxx:21-148   vertices of every cell, facet loops of the active planes
//   GeometryWriter::OutputGeometry     io_utils.cpp:106-231 header, points (x, y, z, 1), primitive attribute
//                                      table, polygons with 16-bit point indices when n_points <= 65536 else 32-bit
//   ConvexCellHost::compute_vertex_coordinates / reload_active   src/rpd3d_base/voronoi_defs.cxx:33-74, 76-106
// Host code only (no kernel): the records come from mb_rpd_run_to_host / mb_rpd_fetch_compact (full or lean
// format) or are passed in the ConvexCellTransfer layout.  Compiled without FMA contraction, so the vertex
// coordinates round exactly like the reference's host build.
#include <cmath>
#include <algorithm>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

#include "mb_internal.h"

namespace {

struct CellRef {  // one cell, whatever the container it came from
  int site, nb_v, nb_p;
  const unsigned char* ver;  // nb_v x 4 bytes (3 plane ids + adjacency)
  const float* plane;        // nb_p equations, `plane_stride` floats apart
  int plane_stride;
  const int* id2;            // nb_p x (x, y), `id2_stride` ints apart
  int id2_stride;
};

inline float det2(float a11, float a12, float a21, float a22) { return a11 * a22 - a12 * a21; }
inline float det3(float a11, float a12, float a13, float a21, float a22, float a23, float a31, float a32, float a33) {
  return a11 * det2(a22, a23, a32, a33) - a21 * det2(a12, a13, a32, a33) + a31 * det2(a12, a13, a22, a23);
}

struct BgeoBuilder {
  std::vector<float> points;        // x, y, z per point
  std::vector<unsigned> poly_idx;   // concatenated polygon point indices
  std::vector<int> poly_len, poly_site;
  int max_sf_fid = 0;
  bool boundary_only = false;

  // io_cuda.cxx:21-148 with is_triangle = false
  void add_cell(const CellRef& c) {
    const unsigned row = (unsigned)(points.size() / 3);
    for (int i = 0; i < c.nb_v; i++) {
      const float* p1 = c.plane + (size_t)c.ver[4 * i + 0] * c.plane_stride;
      const float* p2 = c.plane + (size_t)c.ver[4 * i + 1] * c.plane_stride;
      const float* p3 = c.plane + (size_t)c.ver[4 * i + 2] * c.plane_stride;
      const float x = -det3(p1[3], p1[1], p1[2], p2[3], p2[1], p2[2], p3[3], p3[1], p3[2]);
      const float y = -det3(p1[0], p1[3], p1[2], p2[0], p2[3], p2[2], p3[0], p3[3], p3[2]);
      const float z = -det3(p1[0], p1[1], p1[3], p2[0], p2[1], p2[3], p3[0], p3[1], p3[3]);
      const float w = det3(p1[0], p1[1], p1[2], p2[0], p2[1], p2[2], p3[0], p3[1], p3[2]);
      const float vx = x / w, vy = y / w, vz = z / w;
      if (std::isnan(vx) || std::isnan(vy) || std::isnan(vz)) return;  // the reference reports and drops the rest (:36-46)
      points.push_back(vx);
      points.push_back(vy);
      points.push_back(vz);
    }
    // reload_active (voronoi_defs.cxx:76-91): a plane is active iff a vertex refers to it
    unsigned long long active = 0;
    for (int i = 0; i < c.nb_v; i++)
      active |= (1ull << c.ver[4 * i]) | (1ull << c.ver[4 * i + 1]) | (1ull << c.ver[4 * i + 2]);
    int tab_v[MB_MAX_T], tab_lp[MB_MAX_T];
    for (int plane = 0; plane < c.nb_p; plane++) {
      if (!((active >> plane) & 1ull)) continue;
      const int hx = c.id2[(size_t)plane * c.id2_stride], hy = c.id2[(size_t)plane * c.id2_stride + 1];
      if (boundary_only && !(hy != -1 || hx < max_sf_fid)) continue;  // :60-69
      int n = 0;
      for (int t = 0; t < c.nb_v; t++)
        for (int l = 0; l < 3; l++)
          if ((int)c.ver[4 * t + l] == plane) {
            tab_v[n] = t;
            tab_lp[n] = l;
            n++;
            break;
          }
      // the facet loop (:96-111): after vertex i comes the vertex j whose previous plane is i's next plane
      int i = 0, len = 0;
      const size_t start = poly_idx.size();
      while (len < n) {
        const int ind_i = (tab_lp[i] + 1) % 3;
        const unsigned char want = c.ver[4 * tab_v[i] + ind_i];
        int j = 0;
        for (; j < n; j++)
          if (c.ver[4 * tab_v[j] + (tab_lp[j] + 2) % 3] == want) break;
        if (j == n) break;  // open loop: cannot occur for a valid cell (the reference would not terminate)
        poly_idx.push_back(row + (unsigned)tab_v[i]);
        len++;
        i = j;
      }
      if (len != n) {
        poly_idx.resize(start);
        continue;
      }
      poly_len.push_back(n);
      poly_site.push_back(c.site);
    }
  }

  static void put32(std::vector<unsigned char>& o, uint32_t v) {
    o.push_back((unsigned char)(v >> 24));
    o.push_back((unsigned char)(v >> 16));
    o.push_back((unsigned char)(v >> 8));
    o.push_back((unsigned char)v);
  }
  static void put16(std::vector<unsigned char>& o, uint16_t v) {
    o.push_back((unsigned char)(v >> 8));
    o.push_back((unsigned char)v);
  }
  static uint32_t fbits(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
  }

  // io_utils.cpp:106-231
  void write(const char* path) const {
    const int n_points = (int)(points.size() / 3), n_prims = (int)poly_len.size();
    MB_REQUIRE(n_points > 0, MB_ERR_ARG, "Geometry data does not contain any particles");  // :110-112
    std::vector<unsigned char> o;
    o.reserve(64 + (size_t)n_points * 16 + poly_idx.size() * 4 + (size_t)n_prims * 13);
    put32(o, (((('B' << 8) | 'g') << 8 | 'e') << 8) | 'o');
    o.push_back('V');
    put32(o, 5);
    put32(o, (uint32_t)n_points);
    put32(o, (uint32_t)n_prims);
    put32(o, 0);  // point groups
    put32(o, 0);  // primitive groups
    put32(o, 0);  // point attributes besides P
    put32(o, 0);  // vertex attributes
    put32(o, 1);  // primitive attributes
    put32(o, 0);  // detail attributes
    for (int p = 0; p < n_points; p++) {
      put32(o, fbits(points[3 * (size_t)p]));
      put32(o, fbits(points[3 * (size_t)p + 1]));
      put32(o, fbits(points[3 * (size_t)p + 2]));
      put32(o, fbits(1.0f));
    }
    // primitive attribute table: name, size 1, type INT (= 1), default 0  (WriteHoudiniStr + DefineAttribute)
    const char* name = "PrimAttr";
    put16(o, (uint16_t)strlen(name));
    o.insert(o.end(), name, name + strlen(name));
    put16(o, 1);
    put32(o, 1);
    put32(o, 0);
    const bool wide = n_points > (1 << 16);
    size_t at = 0;
    for (int k = 0; k < n_prims; k++) {
      put32(o, 1);  // polygon
      put32(o, (uint32_t)poly_len[k]);
      o.push_back('<');
      for (int q = 0; q < poly_len[k]; q++) {
        const unsigned idx = poly_idx[at++];
        if (wide)
          put32(o, idx);
        else
          put16(o, (uint16_t)idx);
      }
      put32(o, (uint32_t)poly_site[k]);
    }
    o.push_back(0x00);
    o.push_back(0xff);
    FILE* f = fopen(path, "wb");
    MB_REQUIRE(f != nullptr, MB_ERR_ARG, std::string("cannot open ") + path);
    const size_t wr = fwrite(o.data(), 1, o.size(), f);
    fclose(f);
    MB_REQUIRE(wr == o.size(), MB_ERR_ARG, std::string("short write to ") + path);
  }
};

}  // namespace

static void add_records(BgeoBuilder& B, const unsigned char* recs, long n) {
  for (long i = 0; i < n; i++) {
    const unsigned char* r = recs + (size_t)i * MB_RECORD_BYTES;
    CellRef c;
    memcpy(&c.site, r + 8, 4);
    c.nb_v = r[21];
    c.nb_p = r[22];
    c.ver = r + 24;
    c.plane = reinterpret_cast<const float*>(r + 416);
    c.plane_stride = 8;
    c.id2 = reinterpret_cast<const int*>(r + 2464);
    c.id2_stride = 2;
    B.add_cell(c);
  }
}

// records in the ConvexCellTransfer layout (MB_RECORD_BYTES each)
void bgeo_write_records(const unsigned char* recs, long n, int max_sf_fid, bool boundary_only, const char* path,
                        long* n_points, long* n_polys) {
  BgeoBuilder B;
  B.max_sf_fid = max_sf_fid;
  B.boundary_only = boundary_only;
  add_records(B, recs, n);
  B.write(path);
  if (n_points) *n_points = (long)(B.points.size() / 3);
  if (n_polys) *n_polys = (long)B.poly_len.size();
}

// records produced slice by slice (compact -> ConvexCellTransfer expansion of `count` cells starting at `first`
// into `dst`): the 3 456-byte layout never exists for more than one slice at a time
void bgeo_write_sliced(long n, const std::function<void(long first, long count, unsigned char* dst)>& expand,
                       int max_sf_fid, bool boundary_only, const char* path, long* n_points, long* n_polys) {
  BgeoBuilder B;
  B.max_sf_fid = max_sf_fid;
  B.boundary_only = boundary_only;
  const long slice = 16384;
  std::vector<unsigned char> buf((size_t)std::min(slice, std::max(n, 1L)) * MB_RECORD_BYTES);
  for (long first = 0; first < n; first += slice) {
    const long count = std::min(slice, n - first);
    expand(first, count, buf.data());
    add_records(B, buf.data(), count);
  }
  B.write(path);
  if (n_points) *n_points = (long)(B.points.size() / 3);
  if (n_polys) *n_polys = (long)B.poly_len.size();
}
