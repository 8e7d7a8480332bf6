/* TEST INFRASTRUCTURE ONLY -- see oracle.h.  Plain-C restatement of the reference's RPD3D
 * clipping path.  Pinned against oracle/_ref/libref_rpd.so (the reference's own
 * convex_cell.cu compiled for the host) in tests/test_oracle_vs_ref.py and against the
 * committed golden vectors (tests/golden/kat1_*.npz). */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  float x, y, z, w;
} f4;

/* ---- include/common_cuda.h:103-162, 235-254 (float helpers, literal operation order) ---- */
static f4 minus4(f4 A, f4 B) { return (f4){A.x - B.x, A.y - B.y, A.z - B.z, A.w - B.w}; }
static f4 plus4(f4 A, f4 B) { return (f4){A.x + B.x, A.y + B.y, A.z + B.z, A.w + B.w}; }
static float dot4(f4 A, f4 B) { return A.x * B.x + A.y * B.y + A.z * B.z + A.w * B.w; }
static float dot3(f4 A, f4 B) { return A.x * B.x + A.y * B.y + A.z * B.z; }
static f4 mul3(float s, f4 A) { return (f4){s * A.x, s * A.y, s * A.z, 1.f}; }
static f4 cross3(f4 A, f4 B) {
  return (f4){A.y * B.z - A.z * B.y, A.z * B.x - A.x * B.z, A.x * B.y - A.y * B.x, 0.f};
}
static float det2x2f(float a11, float a12, float a21, float a22) { return a11 * a22 - a12 * a21; }
static float det3x3f(float a11, float a12, float a13, float a21, float a22, float a23, float a31,
                     float a32, float a33) {
  return a11 * det2x2f(a22, a23, a32, a33) - a21 * det2x2f(a12, a13, a32, a33) +
         a31 * det2x2f(a12, a13, a22, a23);
}
/* common_cuda.h:195-213 */
static double det4x4d(double a11, double a12, double a13, double a14, double a21, double a22,
                      double a23, double a24, double a31, double a32, double a33, double a34,
                      double a41, double a42, double a43, double a44) {
  double m12 = a21 * a12 - a11 * a22;
  double m13 = a31 * a12 - a11 * a32;
  double m14 = a41 * a12 - a11 * a42;
  double m23 = a31 * a22 - a21 * a32;
  double m24 = a41 * a22 - a21 * a42;
  double m34 = a41 * a32 - a31 * a42;
  double m123 = m23 * a13 - m13 * a23 + m12 * a33;
  double m124 = m24 * a13 - m14 * a23 + m12 * a43;
  double m134 = m34 * a13 - m14 * a33 + m13 * a43;
  double m234 = m34 * a23 - m24 * a33 + m23 * a43;
  return (m234 * a14 - m134 * a24 + m124 * a34 - m123 * a44);
}

/* ---- per-cell state: the reference keeps it in __shared__ arrays (convex_cell.h:73-92) ---- */
typedef struct {
  uint8_t ver[ORC_MAX_T][4];
  f4 clip[ORC_MAX_P];
  float clip_h[ORC_MAX_P], clip_g[ORC_MAX_P], clip_k[ORC_MAX_P]; /* float7 h,g,k */
  uint8_t edge[ORC_MAX_E][3];
  uint8_t bnext[ORC_MAX_P];
  int status;
  int flagged; /* some conflict |det| fell under the static-filter bound (the USE_ARITHMETIC_FILTER test) */
  int nb_v, nb_r, nb_p, nb_e;
  uint8_t first_boundary;
  int voro_id, tet_id;
  f4 seed;
  const float* pts;
  int pts_pitch;
  const float* w;
} cell_t;

#define END_OF_LIST 255

/* convex_cell.cu:116-214 (ctor in use) with a compact per-tet e_adj6 instead of the dense
 * get_e_adj lookup (convex_cell.h:46-66): same values, (0,1)(0,2)(0,3)(1,2)(1,3)(2,3) order */
static void cell_init(cell_t* c, int seed, const float* pts, int pitch, const float* w, int tid,
                      const float* verts_aos, const int* idx4, const int* v_adjs,
                      const int* e_adj6, const int* f_adjs4, const int* f_ids4) {
  static const int faces[4][3] = {{2, 1, 3}, {0, 2, 3}, {1, 0, 3}, {0, 1, 2}}; /* convex_cell.h:30 */
  c->pts = pts;
  c->pts_pitch = pitch;
  c->w = w;
  c->first_boundary = END_OF_LIST;
  memset(c->bnext, END_OF_LIST, sizeof c->bnext);
  c->voro_id = seed;
  c->seed = (f4){pts[seed], pts[seed + pitch], pts[seed + 2 * pitch], w[seed]};
  c->tet_id = tid;
  c->status = ORC_success;
  c->flagged = 0;
  for (int i = 0; i < 4; i++) {
    f4 v[3];
    for (int j = 0; j < 3; j++) {
      const float* p = verts_aos + 3 * (size_t)idx4[faces[i][j]];
      v[j] = (f4){p[0], p[1], p[2], 0.f};
    }
    /* tri2plane, common_cuda.h:248-254 */
    f4 n = cross3(minus4(v[1], v[0]), minus4(v[2], v[0]));
    c->clip[i] = (f4){n.x, n.y, n.z, -dot3(v[0], n)};
    c->clip_h[i] = (float)f_adjs4[i];
    c->clip_g[i] = (float)f_ids4[i];
    c->clip_k[i] = -1.f;
  }
  c->nb_p = 4;
  static const uint8_t v0[4][3] = {{1, 3, 2}, {0, 2, 3}, {0, 3, 1}, {0, 1, 2}}; /* :186-189 */
  for (int l = 0; l < 4; l++) {
    c->ver[l][0] = v0[l][0];
    c->ver[l][1] = v0[l][1];
    c->ver[l][2] = v0[l][2];
    c->ver[l][3] = (uint8_t)v_adjs[idx4[l]];
  }
  c->nb_v = 4;
  static const uint8_t e0[6][2] = {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}}; /* :202-207 */
  for (int e = 0; e < 6; e++) {
    c->edge[e][0] = e0[e][0];
    c->edge[e][1] = e0[e][1];
    c->edge[e][2] = (uint8_t)e_adj6[e];
  }
  c->nb_e = 6;
}

/* convex_cell.cu:561-592 */
static int new_plane(cell_t* c, int seed_id) {
  if (c->nb_p >= ORC_MAX_P) {
    c->status = ORC_vertex_overflow;
    return -1;
  }
  f4 B = {c->pts[seed_id], c->pts[seed_id + c->pts_pitch], c->pts[seed_id + 2 * c->pts_pitch],
          c->w[seed_id]};
  f4 dir = minus4(c->seed, B);
  f4 ave2 = plus4(c->seed, B);
  float dot = dot3(ave2, dir) + (B.w - c->seed.w);
  int lo = c->voro_id, hi = seed_id;
  if (seed_id < c->voro_id) {
    lo = seed_id;
    hi = c->voro_id;
  }
  c->clip[c->nb_p] = (f4){dir.x, dir.y, dir.z, -dot / 2.f};
  c->clip_h[c->nb_p] = 1.f; /* F_CELL_ADJ_DEFAULT, common.h:61 */
  c->clip_g[c->nb_p] = (float)lo;
  c->clip_k[c->nb_p] = (float)hi;
  c->nb_p++;
  return c->nb_p - 1;
}

static int f4eq(f4 a, f4 b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }

static double dmax4(double a, double b, double c, double d) { return fmax(fmax(a, b), fmax(c, d)); }

/* convex_cell.cu:437-500.  USE_ARITHMETIC_FILTER is off in the live build (voronoi_common.h:32), so the decision is
 * the FP64 sign; the filter test of :479-497 (bound from predicate_generator/main.cpp:52-78) is evaluated on the
 * side and only RECORDED in c->flagged -- the flagged class of SURVEY 8a. */
static int in_conflict(cell_t* c, const uint8_t* v, f4 e) {
  f4 p1 = c->clip[v[0]], p2 = c->clip[v[1]], p3 = c->clip[v[2]];
  if (f4eq(e, p1) || f4eq(e, p2) || f4eq(e, p3)) return 0;
  double det = det4x4d(p1.x, p2.x, p3.x, e.x, p1.y, p2.y, p3.y, e.y, p1.z, p2.z, p3.z, e.z, p1.w,
                       p2.w, p3.w, e.w);
  double maxx = dmax4(fabs((double)p1.x), fabs((double)p2.x), fabs((double)p3.x), fabs((double)e.x));
  double maxy = dmax4(fabs((double)p1.y), fabs((double)p2.y), fabs((double)p3.y), fabs((double)e.y));
  double maxz = dmax4(fabs((double)p1.z), fabs((double)p2.z), fabs((double)p3.z), fabs((double)e.z));
  double eps = 1.2466136531027298e-13 * maxx * maxy * maxz;
  double max_max = fmax(fmax(maxx, maxy), maxz);
  eps *= (max_max * max_max);
  if (fabs(det) < eps) c->flagged = 1;
  return det > 0.0;
}

/* convex_cell.cu:319-351 */
static f4 vertex_coordinates(cell_t* c, const uint8_t* v, int persp) {
  f4 p1 = c->clip[v[0]], p2 = c->clip[v[1]], p3 = c->clip[v[2]];
  f4 r;
  r.x = -det3x3f(p1.w, p1.y, p1.z, p2.w, p2.y, p2.z, p3.w, p3.y, p3.z);
  r.y = -det3x3f(p1.x, p1.w, p1.z, p2.x, p2.w, p2.z, p3.x, p3.w, p3.z);
  r.z = -det3x3f(p1.x, p1.y, p1.w, p2.x, p2.y, p2.w, p3.x, p3.y, p3.w);
  r.w = det3x3f(p1.x, p1.y, p1.z, p2.x, p2.y, p2.z, p3.x, p3.y, p3.z);
  if (r.w == 0.f) c->status = ORC_needs_perturb;
  if (persp) return (f4){r.x / r.w, r.y / r.w, r.z / r.w, 1.f};
  return r;
}

static void swap_ver(cell_t* c, int a, int b) {
  uint8_t t[4];
  memcpy(t, c->ver[a], 4);
  memcpy(c->ver[a], c->ver[b], 4);
  memcpy(c->ver[b], t, 4);
}

/* convex_cell.cu:618-678 */
static void compute_boundary(cell_t* c) {
  memset(c->bnext, END_OF_LIST, sizeof c->bnext);
  c->first_boundary = END_OF_LIST;
  int nb_iter = 0;
  int t = c->nb_v;
  while (c->nb_r > 0) {
    if (nb_iter++ > 65535) {
      c->status = ORC_inconsistent_boundary;
      return;
    }
    int in_border[3], next_is_opp[3];
    const uint8_t* tv = c->ver[t];
    for (int e = 0; e < 3; e++) in_border[e] = (c->bnext[tv[e]] != END_OF_LIST);
    for (int e = 0; e < 3; e++) next_is_opp[e] = (c->bnext[tv[(e + 1) % 3]] == tv[e]);
    int simple = 1;
    for (int e = 0; e < 3; e++)
      if (!next_is_opp[e] && !next_is_opp[(e + 1) % 3] && in_border[(e + 1) % 3]) simple = 0;
    if (!next_is_opp[0] && !next_is_opp[1] && !next_is_opp[2]) {
      if (c->first_boundary == END_OF_LIST) {
        for (int e = 0; e < 3; e++) c->bnext[tv[e]] = tv[(e + 1) % 3];
        c->first_boundary = tv[0];
      } else
        simple = 0;
    }
    if (!simple) {
      t++;
      if (t == c->nb_v + c->nb_r) t = c->nb_v;
      continue;
    }
    for (int e = 0; e < 3; e++)
      if (!next_is_opp[e]) c->bnext[tv[e]] = tv[(e + 1) % 3];
    for (int e = 0; e < 3; e++)
      if (next_is_opp[e] && next_is_opp[(e + 1) % 3]) {
        if (c->first_boundary == tv[(e + 1) % 3]) c->first_boundary = c->bnext[tv[(e + 1) % 3]];
        c->bnext[tv[(e + 1) % 3]] = END_OF_LIST;
      }
    swap_ver(c, t, c->nb_v + c->nb_r - 1);
    t = c->nb_v;
    c->nb_r--;
  }
}

/* convex_cell.cu:511-522 */
static int find_edge_id(const cell_t* c, uint8_t a, uint8_t b) {
  uint8_t lo = a < b ? a : b, hi = a < b ? b : a;
  for (int e = 0; e < c->nb_e; e++)
    if (c->edge[e][0] == lo && c->edge[e][1] == hi) return e;
  return -1;
}

/* convex_cell.cu:680-774 */
static void clip_by_plane(cell_t* c, int neigh) {
  int cur_p = new_plane(c, neigh);
  if (c->status == ORC_vertex_overflow) return;
  f4 eqn = c->clip[cur_p];
  c->nb_r = 0;
  int i = 0;
  while (i < c->nb_v) {
    if (in_conflict(c, c->ver[i], eqn)) {
      c->nb_v--;
      swap_ver(c, i, c->nb_v);
      c->nb_r++;
    } else
      i++;
  }
  if (c->nb_r == 0) {
    c->nb_p--;
    return;
  }
  if (c->nb_v == 0) {
    c->status = ORC_no_intersection;
    return;
  }
  compute_boundary(c);
  if (c->status != ORC_success) return;
  if (c->first_boundary == END_OF_LIST) return;
  /* new_edge, :604-616 */
  uint8_t cir = c->first_boundary;
  do {
    if (c->nb_e >= ORC_MAX_E) {
      c->status = ORC_edge_overflow;
      return;
    }
    uint8_t lo = cur_p < cir ? cur_p : cir, hi = cur_p < cir ? cir : cur_p;
    float h = c->clip_h[cur_p] > c->clip_h[cir] ? c->clip_h[cur_p] : c->clip_h[cir];
    c->edge[c->nb_e][0] = lo;
    c->edge[c->nb_e][1] = hi;
    c->edge[c->nb_e][2] = (uint8_t)h;
    c->nb_e++;
    cir = c->bnext[cir];
  } while (cir != c->first_boundary);
  /* new_vertex :524-554 + is_vertex_perturb :274-316 */
  cir = c->first_boundary;
  do {
    uint8_t nv[3] = {(uint8_t)cur_p, cir, c->bnext[cir]};
    if (c->nb_v + 1 >= ORC_MAX_T) {
      c->status = ORC_triangle_overflow;
    } else {
      int e1 = find_edge_id(c, nv[0], nv[1]);
      int e2 = find_edge_id(c, nv[0], nv[2]);
      int e3 = find_edge_id(c, nv[1], nv[2]);
      uint8_t a = 0;
      if (e1 >= 0 && c->edge[e1][2] > a) a = c->edge[e1][2];
      if (e2 >= 0 && c->edge[e2][2] > a) a = c->edge[e2][2];
      if (e3 >= 0 && c->edge[e3][2] > a) a = c->edge[e3][2];
      c->ver[c->nb_v][0] = nv[0];
      c->ver[c->nb_v][1] = nv[1];
      c->ver[c->nb_v][2] = nv[2];
      c->ver[c->nb_v][3] = a;
      c->nb_v++;
    }
    {
      f4 p1 = c->clip[nv[0]], p2 = c->clip[nv[1]], p3 = c->clip[nv[2]];
      float w = det3x3f(p1.x, p1.y, p1.z, p2.x, p2.y, p2.z, p3.x, p3.y, p3.z);
      if (w == 0.f) c->status = ORC_needs_perturb;
    }
    if (c->status != ORC_success) return;
    cir = c->bnext[cir];
  } while (cir != c->first_boundary);
}

/* common_cuda.h:235-240 */
static f4 project_on_plane(f4 P, f4 plane) {
  f4 n = {plane.x, plane.y, plane.z, 0.f};
  float n_2 = dot4(n, n);
  float lambda = n_2 > 1e-2 ? (dot4(n, P) + plane.w) / n_2 : 0.0f;
  return plus4(P, mul3(-lambda, n));
}

/* convex_cell.cu:986-1069 (atomic_add_bary_and_volume); returns 0 if |vol| < 0.1 */
static int bary_and_volume(cell_t* c, float* bary, int pitch, float* vol) {
  f4 bary_sum = {0, 0, 0, 0};
  float cur = 0;
  f4 P[6];
  f4 C = c->seed;
  for (int t = 0; t < c->nb_v; t++) {
    f4 A = vertex_coordinates(c, c->ver[t], 1);
    f4 A2 = vertex_coordinates(c, c->ver[t], 1); /* :989, same value */
    (void)A2;
    for (int i = 0; i < 3; i++) P[2 * i] = project_on_plane(C, c->clip[c->ver[t][i]]);
    for (int i = 0; i < 3; i++) {
      f4 n = cross3(minus4(P[2 * i], C), minus4(P[(2 * (i + 1)) % 6], C));
      f4 pl = {n.x, n.y, n.z, -dot3(C, n)};
      P[2 * i + 1] = project_on_plane(A, pl);
    }
    for (int i = 0; i < 6; i++) {
      f4 a = minus4(P[i], A), b = minus4(P[(i + 1) % 6], A), cc = minus4(C, A);
      float tv = (float)(-det3x3f(a.x, a.y, a.z, b.x, b.y, b.z, cc.x, cc.y, cc.z) / 6.);
      f4 q0 = P[i], q1 = P[(i + 1) % 6];
      f4 tb = {.25f * (q0.x + q1.x + C.x + A.x), .25f * (q0.y + q1.y + C.y + A.y),
               .25f * (q0.z + q1.z + C.z + A.z), 1.0f};
      bary_sum = plus4(bary_sum, mul3(tv, tb));
      cur += tv;
    }
  }
  if (fabsf(cur) < 0.1) {
    c->status = ORC_no_intersection;
    return 0;
  }
  if (bary) {
    bary[c->voro_id] += bary_sum.x;
    bary[c->voro_id + pitch] += bary_sum.y;
    bary[c->voro_id + 2 * pitch] += bary_sum.z;
  }
  if (vol) vol[c->voro_id] += cur;
  return 1;
}

/* convex_cell.cu:240-268 (the weighted variant): has the cell's security radius been reached by neighbour B?
 * v_dist = largest squared distance of a cell vertex from the seed; d2 = squared distance from the seed to the point
 * where the bisector of (seed, B) crosses the segment seed-B; reached iff d2 > 4 v_dist. */
static int security_radius_reached(cell_t* c, f4 B) {
  float v_dist = 0;
  for (int i = 0; i < c->nb_v; i++) {
    f4 pc = vertex_coordinates(c, c->ver[i], 1);
    f4 diff = minus4(pc, c->seed);
    float d2 = dot3(diff, diff);
    v_dist = d2 > v_dist ? d2 : v_dist;
  }
  f4 diff = minus4(c->seed, B);
  float r2_diff = c->seed.w - B.w;
  float w = (dot3(diff, diff) - r2_diff) / (2 * dot3(diff, diff));
  f4 ph = plus4((f4){w * diff.x, w * diff.y, w * diff.z, w * diff.w}, B);
  f4 vp = minus4(ph, c->seed);
  float d2 = dot3(vp, vp);
  return d2 > 4 * v_dist;
}

/* the security-radius option of the kernel body, restored from the blocks the live reference comments out
 * (convex_cell.cu:1285-1296: leave the neighbour loop once the radius is reached; :1304-1316: a cell whose LAST listed
 * neighbour does not reach it ends as security_radius_not_reached).  Only meaningful for distance-sorted lists. */
static int g_security_radius = 0;
void orc_set_security_radius(int on) { g_security_radius = on; }

/* kernel body convex_cell.cu:1166-1337 + copy :933-949 */
static void run_pair(cell_t* c, const float* verts_aos, const int* idx_aos, const int* v_adjs,
                     const int* e_adj6, const int* f_adjs, const int* f_ids, const float* site_soa,
                     const float* site_w, const unsigned* site_flags, int n_site,
                     const int* site_knn, int site_k, int t, int seed, long p, orc_record* out,
                     int* stat, float* bary, float* vol) {
  out->status = ORC_early_return;
  if (stat) *stat = ORC_security_radius_not_reached; /* voronoi.cu:668 */
  if (seed < 0 || seed >= n_site) return;
  if (site_flags[seed] == 0) return;
  cell_init(c, seed, site_soa, n_site, site_w, t, verts_aos, idx_aos + 4 * (size_t)t, v_adjs,
            e_adj6 + 6 * (size_t)t, f_adjs + 4 * (size_t)t, f_ids + 4 * (size_t)t);
  int last_nb = -1, reached = 0;
  for (int v = 0; v <= site_k - 1; v++) {
    int nb = site_knn[seed + (size_t)v * n_site];
    if (nb == -1) break;
    last_nb = nb;
    clip_by_plane(c, nb);
    if (c->status != ORC_success) {
      if (stat) *stat = c->status;
      return; /* record stays early_return, :1279-1283 */
    }
    if (g_security_radius &&
        security_radius_reached(c, (f4){site_soa[nb], site_soa[nb + n_site], site_soa[nb + 2 * (size_t)n_site], site_w[nb]})) {
      reached = 1;
      break; /* :1285-1296 */
    }
  }
  if (g_security_radius && !reached && c->status != ORC_no_intersection) { /* :1304-1316 */
    if (last_nb < 0 || !security_radius_reached(c, (f4){site_soa[last_nb], site_soa[last_nb + n_site],
                                                        site_soa[last_nb + 2 * (size_t)n_site], site_w[last_nb]})) {
      c->status = ORC_security_radius_not_reached;
      if (stat) *stat = c->status;
      return; /* not a success record: dropped by the host filter (voronoi.cu:749) */
    }
  }
  if (c->status != ORC_no_intersection) {
    out->is_active = 1;
    out->status = c->status;
    out->thread_id = (int)p;
    out->voro_id = c->voro_id;
    out->tet_id = c->tet_id;
    out->euler = -1.f;
    out->weight = c->seed.w;
    out->nb_v = (uint8_t)c->nb_v;
    out->nb_p = (uint8_t)c->nb_p;
    out->nb_e = (uint8_t)c->nb_e;
    for (int i = 0; i < c->nb_v; i++) memcpy(out->ver[i], c->ver[i], 4);
    for (int i = 0; i < c->nb_p; i++) {
      out->clip[i].x = c->clip[i].x;
      out->clip[i].y = c->clip[i].y;
      out->clip[i].z = c->clip[i].z;
      out->clip[i].w = c->clip[i].w;
      out->clip[i].h = c->clip_h[i];
      out->id2[i][0] = (int)lrintf(c->clip_g[i]); /* clip_id2, convex_cell.h:113-115 */
      out->id2[i][1] = (int)lrintf(c->clip_k[i]);
    }
    for (int i = 0; i < c->nb_e; i++) memcpy(out->edge[i], c->edge[i], 3);
    bary_and_volume(c, bary, n_site, vol);
  }
  if (stat) *stat = c->status;
}

double orc_rpd_run_pairs(const float* verts_aos, const int* idx_aos, int n_tet, const int* v_adjs,
                         const int* e_adj6, const int* f_adjs, const int* f_ids,
                         const float* site_soa, const float* site_w, const unsigned* site_flags,
                         int n_site, const int* site_knn, int site_k, const int* pair_tet,
                         const int* pair_site, long n_pairs, orc_record* records, int* stat,
                         float* site_vol, float* site_bary, int n_threads) {
  (void)n_tet;
  int nt = 1;
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
  nt = omp_get_max_threads();
#endif
  float* vol_t = (float*)calloc((size_t)nt * n_site, sizeof(float));
  float* bary_t = (float*)calloc((size_t)nt * 3 * n_site, sizeof(float));
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
#pragma omp parallel
  {
    int th = 0;
#ifdef _OPENMP
    th = omp_get_thread_num();
#endif
    cell_t* c = (cell_t*)malloc(sizeof(cell_t));
#pragma omp for schedule(dynamic, 256)
    for (long p = 0; p < n_pairs; p++)
      run_pair(c, verts_aos, idx_aos, v_adjs, e_adj6, f_adjs, f_ids, site_soa, site_w, site_flags,
               n_site, site_knn, site_k, pair_tet[p], pair_site[p], p, records + p,
               stat ? stat + p : 0, bary_t + (size_t)th * 3 * n_site, vol_t + (size_t)th * n_site);
    free(c);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (site_vol)
    for (int s = 0; s < n_site; s++) {
      float a = 0;
      for (int t = 0; t < nt; t++) a += vol_t[(size_t)t * n_site + s];
      site_vol[s] = a;
    }
  if (site_bary)
    for (long s = 0; s < 3L * n_site; s++) {
      float a = 0;
      for (int t = 0; t < nt; t++) a += bary_t[(size_t)t * 3 * n_site + s];
      site_bary[s] = a;
    }
  free(vol_t);
  free(bary_t);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* the flagged class per candidate pair: the clipping of run_pair with the static-filter test recorded.  Equals the
 * per-pair status needs_exact_predicates of the reference built with USE_ARITHMETIC_FILTER (oracle/_ref/
 * libref_rpd_filter.so): that build stops a cell at its first flagged clip, this one records it and goes on --
 * the tests up to that clip are the same ones. */
void orc_rpd_flagged_pairs(const float* verts_aos, const int* idx_aos, int n_tet, const int* v_adjs,
                           const int* e_adj6, const int* f_adjs, const int* f_ids, const float* site_soa,
                           const float* site_w, const unsigned* site_flags, int n_site, const int* site_knn,
                           int site_k, const int* pair_tet, const int* pair_site, long n_pairs, uint8_t* flagged) {
  (void)n_tet;
#pragma omp parallel
  {
    cell_t* c = (cell_t*)malloc(sizeof(cell_t));
    orc_record* rec = (orc_record*)malloc(sizeof(orc_record));
    float* vol = (float*)calloc((size_t)n_site, sizeof(float));
    float* bary = (float*)calloc(3 * (size_t)n_site, sizeof(float));
#pragma omp for schedule(dynamic, 256)
    for (long p = 0; p < n_pairs; p++) {
      c->flagged = 0;
      run_pair(c, verts_aos, idx_aos, v_adjs, e_adj6, f_adjs, f_ids, site_soa, site_w, site_flags, n_site, site_knn,
               site_k, pair_tet[p], pair_site[p], p, rec, 0, bary, vol);
      flagged[p] = (uint8_t)(c->flagged != 0);
    }
    free(c);
    free(rec);
    free(vol);
    free(bary);
  }
}

/* a3: knncuda.cu:21-94 (compute_distances: ssd += tmp*tmp over x,y,z then 13 zero terms)
 *     + :130-160 (dist_minus_weight);  a4: voronoi.cu:154-193 + host compaction :266-317.
 * Evaluated per (tet, site) with early exit over neighbours (same result: is_relate is an AND) */
static float pdist(const float* site_soa, int n_site, const float* w, int s, const float* p) {
  float ssd = 0.f, tmp;
  tmp = site_soa[s] - p[0];
  ssd += tmp * tmp;
  tmp = site_soa[s + n_site] - p[1];
  ssd += tmp * tmp;
  tmp = site_soa[s + 2 * n_site] - p[2];
  ssd += tmp * tmp;
  return ssd - w[s];
}

long orc_tet_sphere_relation(const float* verts_aos, const int* idx_aos, int n_tet,
                             const float* site_soa, const float* site_w, const unsigned* site_flags,
                             int n_site, const int* site_knn, int site_k, int* pair_tet,
                             int* pair_site, long cap) {
  long* cnt = (long*)calloc((size_t)n_tet + 1, sizeof(long));
  /* two passes: count, then fill (keeps (tet, site) order deterministic under OpenMP) */
  for (int pass = 0; pass < 2; pass++) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int t = 0; t < n_tet; t++) {
      const float* p[4];
      for (int l = 0; l < 4; l++) p[l] = verts_aos + 3 * (size_t)idx_aos[4 * (size_t)t + l];
      long k = 0;
      for (int s = 0; s < n_site; s++) {
        if (site_flags[s] != 1) continue;
        float pd_i[4];
        for (int l = 0; l < 4; l++) pd_i[l] = pdist(site_soa, n_site, site_w, s, p[l]);
        int ok = 1;
        for (int sm = 0; sm < site_k && ok; sm++) {
          int m = site_knn[s + (size_t)sm * n_site];
          if (m == -1) continue;
          int any = 0;
          for (int l = 0; l < 4 && !any; l++)
            if (pdist(site_soa, n_site, site_w, m, p[l]) > pd_i[l]) any = 1;
          if (!any) ok = 0;
        }
        if (ok) {
          if (pass == 1) {
            long o = cnt[t] + k;
            if (o < cap) {
              pair_tet[o] = t;
              pair_site[o] = s;
            }
          }
          k++;
        }
      }
      if (pass == 0) cnt[t + 1] = k;
    }
    if (pass == 0) {
      cnt[0] = 0;
      for (int t = 0; t < n_tet; t++) cnt[t + 1] += cnt[t];
      if (cnt[n_tet] > cap) {
        long need = cnt[n_tet];
        free(cnt);
        return -need;
      }
    }
  }
  long total = cnt[n_tet];
  free(cnt);
  return total;
}

/* voronoi_defs.cxx:76-106 (reload_active) + :190-222 (cal_cell_euler) */
static void reload_active_one(const orc_record* r, uint8_t* ap, uint8_t* ae, float* euler) {
  int cntp[ORC_MAX_P + 1];
  memset(cntp, 0, sizeof cntp);
  memset(ap, 0, ORC_MAX_P);
  memset(ae, 0, ORC_MAX_E);
  for (int t = 0; t < r->nb_v; t++)
    for (int k = 0; k < 3; k++) cntp[r->ver[t][k]]++;
  for (int p = 0; p < r->nb_p; p++) ap[p] = cntp[p] > 0;
  for (int e = 0; e < r->nb_e; e++) {
    int a = r->edge[e][0], b = r->edge[e][1];
    if (cntp[a] <= 0 || cntp[b] <= 0) continue;
    int shared = 0;
    for (int t = 0; t < r->nb_v; t++) {
      int ha = 0, hb = 0;
      for (int k = 0; k < 3; k++) {
        ha |= r->ver[t][k] == a;
        hb |= r->ver[t][k] == b;
      }
      shared += (ha && hb);
    }
    if (shared < 2) continue;
    ae[e] = 1;
  }
  if (euler) {
    double sv = 0, sf = 0, se = 0; /* Scalar = double, common.h:21 */
    for (int v = 0; v < r->nb_v; v++) sv += 1.f / (int)r->ver[v][3];
    for (int f = 0; f < r->nb_p; f++)
      if (ap[f]) sf += 1. / r->clip[f].h;
    for (int e = 0; e < r->nb_e; e++)
      if (ae[e]) se += 1. / r->edge[e][2];
    *euler = (float)(sv - se + sf);
  }
}

void orc_reload_active(const orc_record* recs, long n, uint8_t* active_planes,
                       uint8_t* active_edges, float* euler) {
#pragma omp parallel for
  for (long i = 0; i < n; i++)
    reload_active_one(recs + i, active_planes + i * ORC_MAX_P, active_edges + i * ORC_MAX_E,
                      euler ? euler + i : 0);
}

static int cmp_u8x4(const void* a, const void* b) { return memcmp(a, b, 4); }
static int cmp_u8x3(const void* a, const void* b) { return memcmp(a, b, 3); }

void orc_canonicalize(const orc_record* recs, long n, orc_record* out) {
#pragma omp parallel for
  for (long i = 0; i < n; i++) {
    const orc_record* r = recs + i;
    orc_record* o = out + i;
    memset(o, 0, sizeof *o);
    o->status = r->status;
    o->voro_id = r->voro_id;
    o->tet_id = r->tet_id;
    o->weight = r->weight;
    o->is_active = r->is_active;
    o->euler = -1.f;
    o->cell_vol = -1.f;
    o->id = r->id;
    if (r->status != ORC_success) continue;
    uint8_t ap[ORC_MAX_P], ae[ORC_MAX_E];
    reload_active_one(r, ap, ae, 0);
    /* plane relabelling: 0-3 stay; active bisectors sorted by (id2.x, id2.y) */
    int order[ORC_MAX_P], m = 0;
    for (int p = 4; p < r->nb_p; p++)
      if (ap[p]) order[m++] = p;
    for (int a = 1; a < m; a++) { /* insertion sort by id2 */
      int x = order[a], b = a - 1;
      while (b >= 0 && (r->id2[order[b]][0] > r->id2[x][0] ||
                        (r->id2[order[b]][0] == r->id2[x][0] && r->id2[order[b]][1] > r->id2[x][1]))) {
        order[b + 1] = order[b];
        b--;
      }
      order[b + 1] = x;
    }
    uint8_t map[ORC_MAX_P];
    memset(map, 255, sizeof map);
    int np = 4;
    for (int p = 0; p < 4; p++) map[p] = (uint8_t)p;
    for (int a = 0; a < m; a++) map[order[a]] = (uint8_t)np++;
    o->nb_p = (uint8_t)np;
    for (int p = 0; p < r->nb_p; p++) {
      if (map[p] == 255) continue;
      o->clip[map[p]] = r->clip[p];
      /* inactive tet faces keep their slot but are marked h = -h so that the comparison
       * still sees whether the face is active */
      if (p < 4 && !ap[p]) o->clip[map[p]].h = -r->clip[p].h;
      o->id2[map[p]][0] = r->id2[p][0];
      o->id2[map[p]][1] = r->id2[p][1];
    }
    o->nb_v = r->nb_v;
    for (int t = 0; t < r->nb_v; t++) {
      uint8_t a = map[r->ver[t][0]], b = map[r->ver[t][1]], c = map[r->ver[t][2]];
      /* rotate (orientation kept) so that the smallest label comes first */
      if (b < a && b < c) {
        uint8_t t0 = a;
        a = b;
        b = c;
        c = t0;
      } else if (c < a && c < b) {
        uint8_t t0 = c;
        c = b;
        b = a;
        a = t0;
      }
      o->ver[t][0] = a;
      o->ver[t][1] = b;
      o->ver[t][2] = c;
      o->ver[t][3] = r->ver[t][3];
    }
    qsort(o->ver, o->nb_v, 4, cmp_u8x4);
    int ne = 0;
    for (int e = 0; e < r->nb_e; e++) {
      if (!ae[e]) continue;
      uint8_t a = map[r->edge[e][0]], b = map[r->edge[e][1]];
      o->edge[ne][0] = a < b ? a : b;
      o->edge[ne][1] = a < b ? b : a;
      o->edge[ne][2] = r->edge[e][2];
      ne++;
    }
    qsort(o->edge, ne, 3, cmp_u8x3);
    o->nb_e = (uint8_t)ne;
  }
}

/* voronoi_defs.cxx:33-74 == convex_cell.cu:319-351 */
void orc_vertex_coordinates(const orc_record* recs, long n, float* out) {
#pragma omp parallel for
  for (long i = 0; i < n; i++) {
    const orc_record* r = recs + i;
    if (r->status != ORC_success) continue;
    for (int t = 0; t < r->nb_v; t++) {
      const orc_float5 *p1 = &r->clip[r->ver[t][0]], *p2 = &r->clip[r->ver[t][1]],
                       *p3 = &r->clip[r->ver[t][2]];
      float x = -det3x3f(p1->w, p1->y, p1->z, p2->w, p2->y, p2->z, p3->w, p3->y, p3->z);
      float y = -det3x3f(p1->x, p1->w, p1->z, p2->x, p2->w, p2->z, p3->x, p3->w, p3->z);
      float z = -det3x3f(p1->x, p1->y, p1->w, p2->x, p2->y, p2->w, p3->x, p3->y, p3->w);
      float w = det3x3f(p1->x, p1->y, p1->z, p2->x, p2->y, p2->z, p3->x, p3->y, p3->z);
      float* o = out + (i * ORC_MAX_T + t) * 4;
      o[0] = x / w;
      o[1] = y / w;
      o[2] = z / w;
      o[3] = w;
    }
  }
}

/* exact-ish cell volume in double: V = 1/3 * sum_faces dist(origin-shifted plane) * area, computed
 * as sum over faces of fan triangles around the face's vertex loop (order from the dual-triangle
 * orientation, like reload_pc_explicit voronoi_defs.cxx:108-188).  Property tests only. */
static void vtx_d(const orc_record* r, int t, double* o) {
  const orc_float5 *a = &r->clip[r->ver[t][0]], *b = &r->clip[r->ver[t][1]], *c = &r->clip[r->ver[t][2]];
  double A[3][3] = {{a->x, a->y, a->z}, {b->x, b->y, b->z}, {c->x, c->y, c->z}};
  double d[3] = {-(double)a->w, -(double)b->w, -(double)c->w};
  double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) -
               A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
               A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
  for (int k = 0; k < 3; k++) {
    double M[3][3];
    memcpy(M, A, sizeof M);
    for (int q = 0; q < 3; q++) M[q][k] = d[q];
    double dk = M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) -
                M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
                M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
    o[k] = dk / det;
  }
}

void orc_cell_volumes(const orc_record* recs, long n, double* vol) {
#pragma omp parallel for
  for (long i = 0; i < n; i++) {
    const orc_record* r = recs + i;
    vol[i] = 0;
    if (r->status != ORC_success) continue;
    double P[ORC_MAX_T][3];
    for (int t = 0; t < r->nb_v; t++) vtx_d(r, t, P[t]);
    double c0[3] = {P[0][0], P[0][1], P[0][2]};
    double V = 0;
    for (int p = 0; p < r->nb_p; p++) {
      /* vertices on plane p, ordered by walking: next(t) = the vertex u on p whose
       * plane-after-p equals t's plane-before-p ... simpler: angular sort around the centroid */
      int idx[ORC_MAX_T], m = 0;
      for (int t = 0; t < r->nb_v; t++)
        if (r->ver[t][0] == p || r->ver[t][1] == p || r->ver[t][2] == p) idx[m++] = t;
      if (m < 3) continue;
      double g[3] = {0, 0, 0};
      for (int a = 0; a < m; a++)
        for (int k = 0; k < 3; k++) g[k] += P[idx[a]][k] / m;
      double nrm[3] = {r->clip[p].x, r->clip[p].y, r->clip[p].z};
      double nl = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
      if (nl == 0) continue;
      for (int k = 0; k < 3; k++) nrm[k] /= nl;
      double u[3], w[3];
      int kmin = fabs(nrm[0]) < fabs(nrm[1]) ? (fabs(nrm[0]) < fabs(nrm[2]) ? 0 : 2)
                                             : (fabs(nrm[1]) < fabs(nrm[2]) ? 1 : 2);
      double e[3] = {0, 0, 0};
      e[kmin] = 1;
      u[0] = nrm[1] * e[2] - nrm[2] * e[1];
      u[1] = nrm[2] * e[0] - nrm[0] * e[2];
      u[2] = nrm[0] * e[1] - nrm[1] * e[0];
      w[0] = nrm[1] * u[2] - nrm[2] * u[1];
      w[1] = nrm[2] * u[0] - nrm[0] * u[2];
      w[2] = nrm[0] * u[1] - nrm[1] * u[0];
      double ang[ORC_MAX_T];
      for (int a = 0; a < m; a++) {
        double d[3] = {P[idx[a]][0] - g[0], P[idx[a]][1] - g[1], P[idx[a]][2] - g[2]};
        ang[a] = atan2(d[0] * w[0] + d[1] * w[1] + d[2] * w[2], d[0] * u[0] + d[1] * u[1] + d[2] * u[2]);
      }
      for (int a = 1; a < m; a++) {
        int x = idx[a];
        double xa = ang[a];
        int b = a - 1;
        while (b >= 0 && ang[b] > xa) {
          idx[b + 1] = idx[b];
          ang[b + 1] = ang[b];
          b--;
        }
        idx[b + 1] = x;
        ang[b + 1] = xa;
      }
      /* area vector of the polygon and the pyramid volume with apex c0 */
      double av[3] = {0, 0, 0};
      for (int a = 0; a < m; a++) {
        const double* q0 = P[idx[a]];
        const double* q1 = P[idx[(a + 1) % m]];
        double d0[3] = {q0[0] - g[0], q0[1] - g[1], q0[2] - g[2]};
        double d1[3] = {q1[0] - g[0], q1[1] - g[1], q1[2] - g[2]};
        av[0] += 0.5 * (d0[1] * d1[2] - d0[2] * d1[1]);
        av[1] += 0.5 * (d0[2] * d1[0] - d0[0] * d1[2]);
        av[2] += 0.5 * (d0[0] * d1[1] - d0[1] * d1[0]);
      }
      double area = sqrt(av[0] * av[0] + av[1] * av[1] + av[2] * av[2]);
      double hgt = fabs((g[0] - c0[0]) * nrm[0] + (g[1] - c0[1]) * nrm[1] + (g[2] - c0[2]) * nrm[2]);
      V += area * hgt / 3.0;
    }
    vol[i] = V;
  }
}

/* ---- a14 second half: reload_pc_explicit (voronoi_defs.cxx:108-188) + the per-cell part of
 * get_all_voro_info (rpd_update.cxx:112-301) ------------------------------------------------- */
/* face loop of an active plane: vertices referencing it, in the cyclic order of the walk at
 * voronoi_defs.cxx:163-182.  Returns the loop length. */
static int face_loop(const orc_record* r, int plane, int* loop) {
  int tab_v[ORC_MAX_T], tab_lp[ORC_MAX_T], m = 0;
  for (int t = 0; t < r->nb_v; t++) {
    if (r->ver[t][0] == plane) { tab_v[m] = t; tab_lp[m++] = 0; }
    else if (r->ver[t][1] == plane) { tab_v[m] = t; tab_lp[m++] = 1; }
    else if (r->ver[t][2] == plane) { tab_v[m] = t; tab_lp[m++] = 2; }
  }
  int i = 0, n = 0;
  while (n < m) {
    int ind_i = (tab_lp[i] + 1) % 3;
    int found = 0;
    for (int j = 0; j < m && !found; j++) {
      int ind_j = (tab_lp[j] + 2) % 3;
      if (r->ver[tab_v[i]][ind_i] == r->ver[tab_v[j]][ind_j]) {
        loop[n++] = tab_v[i];
        found = 1;
        i = j;
      }
    }
    if (!found) break; /* the reference would run off the table here */
  }
  return n;
}

static int cmp_int(const void* a, const void* b) { return *(const int*)a - *(const int*)b; }

long orc_emit(const orc_record* recs, long n, int max_surf_fid, long cap, long* counts,
              int* f_cell, int* f_key, uint8_t* f_istet, float* f_centroid3, int* v_cell,
              int* v_lvid, int* v_key3, float* v_pos3, int* v_surf, int* e_cell, int* e_key2,
              int* e_lvid2) {
  long nf = 0, nv = 0, ne = 0;
  for (long c = 0; c < n; c++) {
    const orc_record* r = recs + c;
    if (r->status != ORC_success) continue;
    uint8_t ap[ORC_MAX_P], ae[ORC_MAX_E];
    reload_active_one(r, ap, ae, 0);
    float pos[ORC_MAX_T][3];
    for (int t = 0; t < r->nb_v; t++) {
      const orc_float5 *p1 = &r->clip[r->ver[t][0]], *p2 = &r->clip[r->ver[t][1]],
                       *p3 = &r->clip[r->ver[t][2]];
      float x = -det3x3f(p1->w, p1->y, p1->z, p2->w, p2->y, p2->z, p3->w, p3->y, p3->z);
      float y = -det3x3f(p1->x, p1->w, p1->z, p2->x, p2->w, p2->z, p3->x, p3->w, p3->z);
      float z = -det3x3f(p1->x, p1->y, p1->w, p2->x, p2->y, p2->w, p3->x, p3->y, p3->w);
      float w = det3x3f(p1->x, p1->y, p1->z, p2->x, p2->y, p2->z, p3->x, p3->y, p3->z);
      pos[t][0] = x / w; pos[t][1] = y / w; pos[t][2] = z / w;
    }
    /* facets, rpd_update.cxx:121-140 */
    for (int p = 0; p < r->nb_p; p++) {
      if (!ap[p]) continue;
      if (nf < cap) {
        f_cell[nf] = (int)c;
        float cx = 0.f, cy = 0.f, cz = 0.f;
        if (r->id2[p][1] != -1) {
          f_key[nf] = r->id2[p][0] == r->voro_id ? r->id2[p][1] : r->id2[p][0];
          f_istet[nf] = 0;
        } else {
          f_key[nf] = r->id2[p][0];
          f_istet[nf] = 1;
          if (r->id2[p][0] <= max_surf_fid) { /* get_cell_v2surffid, rpd_update.cxx:20-42 */
            int loop[ORC_MAX_T];
            int m = face_loop(r, p, loop);
            for (int k = 0; k < m; k++) { cx = cx + pos[loop[k]][0]; cy = cy + pos[loop[k]][1]; cz = cz + pos[loop[k]][2]; }
            cx = cx / (float)m; cy = cy / (float)m; cz = cz / (float)m;
          }
        }
        f_centroid3[3 * nf] = cx; f_centroid3[3 * nf + 1] = cy; f_centroid3[3 * nf + 2] = cz;
      }
      nf++;
    }
    /* vertices, rpd_update.cxx:147-191 */
    for (int t = 0; t < r->nb_v; t++) {
      int seeds[3], ns = 0, surf = -1;
      for (int i = 0; i < 3; i++) {
        int lf = r->ver[t][i];
        if (r->id2[lf][1] != -1) {
          int nb = r->id2[lf][0] == r->voro_id ? r->id2[lf][1] : r->id2[lf][0];
          int dup = 0;
          for (int q = 0; q < ns; q++) dup |= seeds[q] == nb;
          if (!dup) seeds[ns++] = nb;
        } else if (r->id2[lf][0] <= max_surf_fid)
          surf = r->id2[lf][0];
      }
      if (ns < 2) continue;
      if (ns == 2) seeds[ns++] = -1;
      qsort(seeds, 3, sizeof(int), cmp_int);
      if (nv < cap) {
        v_cell[nv] = (int)c; v_lvid[nv] = t;
        memcpy(v_key3 + 3 * nv, seeds, sizeof seeds);
        memcpy(v_pos3 + 3 * nv, pos[t], 3 * sizeof(float));
        v_surf[nv] = surf;
      }
      nv++;
    }
    /* edges between two bisectors, rpd_update.cxx:195-200, 261-295 */
    for (int e = 0; e < r->nb_e; e++) {
      if (!ae[e]) continue;
      int a = r->edge[e][0], b = r->edge[e][1];
      if (r->id2[a][1] == -1 || r->id2[b][1] == -1) continue;
      int k[2] = {r->id2[a][0] == r->voro_id ? r->id2[a][1] : r->id2[a][0],
                  r->id2[b][0] == r->voro_id ? r->id2[b][1] : r->id2[b][0]};
      if (k[0] > k[1]) { int s = k[0]; k[0] = k[1]; k[1] = s; }
      /* end vertices = sorted intersection of the two face loops' vertex sets */
      int la[ORC_MAX_T], lb[ORC_MAX_T];
      int ma = face_loop(r, a, la), mb = face_loop(r, b, lb);
      qsort(la, ma, sizeof(int), cmp_int);
      qsort(lb, mb, sizeof(int), cmp_int);
      int ends[2] = {-1, -1}, m = 0;
      for (int i = 0, j = 0; i < ma && j < mb;) {
        if (la[i] < lb[j]) i++;
        else if (la[i] > lb[j]) j++;
        else { if (m < 2) ends[m] = la[i]; m++; i++; j++; }
      }
      if (ne < cap) {
        e_cell[ne] = (int)c;
        e_key2[2 * ne] = k[0]; e_key2[2 * ne + 1] = k[1];
        e_lvid2[2 * ne] = ends[0]; e_lvid2[2 * ne + 1] = ends[1];
      }
      ne++;
    }
  }
  counts[0] = nf; counts[1] = nv; counts[2] = ne;
  return (nf <= cap && nv <= cap && ne <= cap) ? 0 : -1;
}
