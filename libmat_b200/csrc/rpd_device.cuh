// Device-side building blocks of the RPD3D path (sm_100a).
//
// Arithmetic contract: every value that can influence a combinatorial decision or that is stored
// in the output record is computed with explicitly rounded, NON-fused IEEE operations
// (__fmul_rn / __fadd_rn / __dmul_rn / __dadd_rn ...) in the exact operation order of the
// reference expressions, so that the result is bit-identical to the reference's own code built
// for the host (oracle/_ref, g++ -O2 -ffp-contract=off).  Conservative *filters* that only ever
// skip work whose outcome is certain use fused arithmetic freely.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstddef>

#define MBK_MAX_P 64
#define MBK_MAX_T 96
#define MBK_MAX_E 152
#define MBK_END 255
#define MBK_CV 32  // vertices with an FP32 cofactor-filter cache entry (the rest take the FP64 path)

enum : int {
  ST_early_return = -1,
  ST_triangle_overflow = 0,
  ST_vertex_overflow = 1,
  ST_inconsistent_boundary = 2,
  ST_security_radius_not_reached = 3,
  ST_success = 4,
  ST_needs_exact_predicates = 5,
  ST_no_intersection = 6,
  ST_edge_overflow = 7,
  ST_needs_perturb = 8
};

// ---- exact (non-fused) float helpers: include/common_cuda.h:103-162 of the reference --------
__device__ __forceinline__ float xfsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xfadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xfmul(float a, float b) { return __fmul_rn(a, b); }

// dot3(A,B) = A.x*B.x + A.y*B.y + A.z*B.z (left-to-right), common_cuda.h:127-132
__device__ __forceinline__ float dot3_exact(float ax, float ay, float az, float bx, float by,
                                            float bz) {
  return xfadd(xfadd(xfmul(ax, bx), xfmul(ay, by)), xfmul(az, bz));
}

// det2x2 / det3x3 float, common_cuda.h:153-162
__device__ __forceinline__ float det2_exact(float a11, float a12, float a21, float a22) {
  return xfsub(xfmul(a11, a22), xfmul(a12, a21));
}
__device__ __forceinline__ float det3_exact(float a11, float a12, float a13, float a21, float a22,
                                            float a23, float a31, float a32, float a33) {
  return xfadd(xfsub(xfmul(a11, det2_exact(a22, a23, a32, a33)),
                   xfmul(a21, det2_exact(a12, a13, a32, a33))),
              xfmul(a31, det2_exact(a12, a13, a22, a23)));
}

// tri2plane, common_cuda.h:248-254 (with cross3 :143-146, plane_from_point_and_normal :150-152)
__device__ __forceinline__ float4 tri2plane_exact(float3 v0, float3 v1, float3 v2) {
  float ax = xfsub(v1.x, v0.x), ay = xfsub(v1.y, v0.y), az = xfsub(v1.z, v0.z);
  float bx = xfsub(v2.x, v0.x), by = xfsub(v2.y, v0.y), bz = xfsub(v2.z, v0.z);
  float nx = xfsub(xfmul(ay, bz), xfmul(az, by));
  float ny = xfsub(xfmul(az, bx), xfmul(ax, bz));
  float nz = xfsub(xfmul(ax, by), xfmul(ay, bx));
  float d = -dot3_exact(v0.x, v0.y, v0.z, nx, ny, nz);
  return make_float4(nx, ny, nz, d);
}

// power bisector of seed A and neighbour B, ConvexCell::new_plane convex_cell.cu:561-592:
// n = A - B, d = -(dot3(A+B, n) + (w_B - w_A)) / 2
__device__ __forceinline__ float4 bisector_exact(float4 A, float4 B) {
  float dx = xfsub(A.x, B.x), dy = xfsub(A.y, B.y), dz = xfsub(A.z, B.z);
  float sx = xfadd(A.x, B.x), sy = xfadd(A.y, B.y), sz = xfadd(A.z, B.z);
  float dot = xfadd(dot3_exact(sx, sy, sz, dx, dy, dz), xfsub(B.w, A.w));
  return make_float4(dx, dy, dz, __fdiv_rn(-dot, 2.f));
}

// ---- the FP64 conflict predicate: cc_vertex_is_in_conflict_double convex_cell.cu:437-500 with
// det4x4(double) common_cuda.h:195-213, literal operation order, no FMA -----------------------
__device__ __forceinline__ double xdmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xdsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xdadd(double a, double b) { return __dadd_rn(a, b); }

struct Minors {  // the 3x3 minors of the three vertex planes (depend on the vertex only)
  double m123, m124, m134, m234;
};

__device__ __forceinline__ Minors minors_exact(float4 p1, float4 p2, float4 p3) {
  const double a11 = p1.x, a12 = p2.x, a13 = p3.x;
  const double a21 = p1.y, a22 = p2.y, a23 = p3.y;
  const double a31 = p1.z, a32 = p2.z, a33 = p3.z;
  const double a41 = p1.w, a42 = p2.w, a43 = p3.w;
  const double m12 = xdsub(xdmul(a21, a12), xdmul(a11, a22));
  const double m13 = xdsub(xdmul(a31, a12), xdmul(a11, a32));
  const double m14 = xdsub(xdmul(a41, a12), xdmul(a11, a42));
  const double m23 = xdsub(xdmul(a31, a22), xdmul(a21, a32));
  const double m24 = xdsub(xdmul(a41, a22), xdmul(a21, a42));
  const double m34 = xdsub(xdmul(a41, a32), xdmul(a31, a42));
  Minors m;
  m.m123 = xdadd(xdsub(xdmul(m23, a13), xdmul(m13, a23)), xdmul(m12, a33));
  m.m124 = xdadd(xdsub(xdmul(m24, a13), xdmul(m14, a23)), xdmul(m12, a43));
  m.m134 = xdadd(xdsub(xdmul(m34, a13), xdmul(m14, a33)), xdmul(m13, a43));
  m.m234 = xdadd(xdsub(xdmul(m34, a23), xdmul(m24, a33)), xdmul(m23, a43));
  return m;
}

__device__ __forceinline__ double det4_from_minors(const Minors& m, float4 e) {
  // (m234*a14 - m134*a24 + m124*a34 - m123*a44)
  return xdsub(xdadd(xdsub(xdmul(m.m234, (double)e.x), xdmul(m.m134, (double)e.y)),
                   xdmul(m.m124, (double)e.z)),
              xdmul(m.m123, (double)e.w));
}

__device__ __forceinline__ bool plane_eq(float4 a, float4 b) {
  return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w;
}

__device__ __forceinline__ bool conflict_exact(float4 p1, float4 p2, float4 p3, float4 e) {
  if (plane_eq(e, p1) || plane_eq(e, p2) || plane_eq(e, p3)) return false;
  Minors m = minors_exact(p1, p2, p3);
  return det4_from_minors(m, e) > 0.0;
}

// ---- the flagged class: the static filter of src/predicate_generator (main.cpp:52-78), consumed by the reference's
// USE_ARITHMETIC_FILTER branch (convex_cell.cu:479-497; compiled out in the live build, voronoi_common.h:32):
//   eps = 1.2466136531027298e-13 * maxx * maxy * maxz * max(maxx, maxy, maxz)^2,   flagged <=> |det| < eps
// with maxx/y/z the largest |normal component| over the vertex's three planes and the new plane.  A cell with a
// flagged conflict test is one whose combinatorics the FP64 determinant does not certify: the reference's filter build
// drops it as needs_exact_predicates; here the decision stays the FP64 one (bit-identical to the live reference) and
// the cell carries a flag (record word 2 bit 30, mb_rpd_fetch_flags).
#define MBK_FILTER_BOUND_F64 1.2466136531027298e-13
__device__ __forceinline__ bool conflict_exact_flag(float4 p1, float4 p2, float4 p3, float4 e, bool& flagged) {
  if (plane_eq(e, p1) || plane_eq(e, p2) || plane_eq(e, p3)) return false;  // :444-451, before the determinant
  Minors m = minors_exact(p1, p2, p3);
  const double det = det4_from_minors(m, e);
  const double maxx = (double)fmaxf(fmaxf(fabsf(p1.x), fabsf(p2.x)), fmaxf(fabsf(p3.x), fabsf(e.x)));
  const double maxy = (double)fmaxf(fmaxf(fabsf(p1.y), fabsf(p2.y)), fmaxf(fabsf(p3.y), fabsf(e.y)));
  const double maxz = (double)fmaxf(fmaxf(fabsf(p1.z), fabsf(p2.z)), fmaxf(fabsf(p3.z), fabsf(e.z)));
  double eps = xdmul(xdmul(xdmul(MBK_FILTER_BOUND_F64, maxx), maxy), maxz);  // literal left-to-right order (:486)
  const double mm = fmax(maxx, fmax(maxy, maxz));
  eps = xdmul(eps, xdmul(mm, mm));                                            // :492
  flagged = flagged || (fabs(det) < eps);
  return det > 0.0;
}

// conservative FP32 upper bound of that eps from M >= every |normal component| involved: eps <= K * M^5
__device__ __forceinline__ float filter_eps_upper(float M) {
  return (M * M) * (M * M) * (M * 1.2467e-13f) * 1.0001f;
}

// ---- per-cell shared-memory state ---------------------------------------------------------------
// KP / KT / KE = capacity in planes / vertices (dual triangles) / edges.  The reference's caps are 64 / 96 / 152
// (_MAX_P_ / _MAX_T_ / _MAX_E_); the grid-kNN first pass runs with compact caps (more cells resident per SM) and
// hands the rare cell that outgrows them to a second pass at the reference's caps.
template <int KP, int KT, int KE>
struct __align__(16) CellT {
  float4 plane[KP];                 // plane equations (a,b,c,d)
  float4 c0[4];                     // cull filter: cofactor vectors of the 4 initial (tet) vertices
  float4 cof[MBK_CV];               // conflict filter: cofactor vector of live vertex v (v < MBK_CV)
  int pnb[KP];                      // p<4: tet-face id; p>=4: neighbour site id of the bisector
  uchar4 ver[KT];                   // dual triangles (3 plane ids, #adjacent cells)
  unsigned char bnext[KP];          // cavity boundary circular list (16-byte aligned: cleared with vector stores)
  unsigned char cyc[KP];            // the boundary cycle in walk order
  unsigned char edge[KE * 3];       // (plane a, plane b, #adjacent cells)
  unsigned long long adj[KP];       // grid-kNN mode: directed dual-edge bit matrix of the cavity (zero between clips)
};
typedef CellT<MBK_MAX_P, MBK_MAX_T, MBK_MAX_E> CellS;
static_assert(offsetof(CellS, bnext) % 16 == 0, "bnext must be 16-byte aligned");
// compact caps of the first pass: 2 640 B per cell -> 5 blocks of 16 cells per SM instead of 4
#define MBK_SMALL_P 48
#define MBK_SMALL_T 64   // <= 64: the conflict flags of a clip fit one 64-bit word (the second word's code is compiled out)
#define MBK_SMALL_E 120
typedef CellT<MBK_SMALL_P, MBK_SMALL_T, MBK_SMALL_E> CellSmall;
static_assert(offsetof(CellSmall, bnext) % 16 == 0, "bnext must be 16-byte aligned");
static_assert(MBK_SMALL_P % 16 == 0 && MBK_MAX_P % 16 == 0, "bnext is cleared 16 bytes at a time");

// compact record size in 4-byte words: header 4 + ver nb_v + planes 4*nb_p + (id2,h) 3*nb_p +
// edges ceil(3*nb_e/4)
__host__ __device__ __forceinline__ int compact_words(int nb_v, int nb_p, int nb_e) {
  return 4 + nb_v + 7 * nb_p + (3 * nb_e + 3) / 4;
}
