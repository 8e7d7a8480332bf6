/* TEST INFRASTRUCTURE ONLY -- see oracle.h.  Plain-C restatement of the reference's dist2mat
 * path (src/dist2mat/dist2mat.cu).  Pinned against oracle/_ref/libref_d2m.so (the reference's
 * own __host__ __device__ functions) and tests/golden/kat2_dist2mat.json (KAT-2 / KAT-2b). */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  float x, y, z;
} f3;
typedef struct {
  float x, y, z, w;
} f4;

static float dotf(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } /* helper_math:1015 */
static f3 sub3(f3 a, f3 b) { return (f3){a.x - b.x, a.y - b.y, a.z - b.z}; }
static float length3(f3 v) { return sqrtf(dotf(v, v)); } /* :1047 */
/* cuda_helper_math.h:932-934 clamp = fmaxf(a, fminf(f, b)) AS THE REFERENCE'S DEVICE BUILD EVALUATES IT: nvcc lowers
 * clamp(t, 0, 1) to a saturate modifier, which maps a NaN t to +0 (IEEE fminf/fmaxf would give 1).  t is NaN for
 * nested spheres (dist2mat.cu:61); verified against the device build on the B200 (tests/test_gpu_reference_build.py). */
static float clampf(float f, float a, float b) { return (f != f) ? a : fmaxf(a, fminf(f, b)); }
static float lerpf(float a, float b, float t) { return b + t * (a - b); } /* :911-913 */

/* dist2mat.cu:5-8 */
static float d_sphere(f3 p, f4 sp) { return length3(sub3(p, (f3){sp.x, sp.y, sp.z})) - sp.w; }

/* dist2mat.cu:10-15 */
static f4 bary_lerp(f4 m1, f4 m2, f4 m3, float t1, float t2) {
  float t3 = 1.f - t1 - t2;
  return (f4){m1.x * t1 + m2.x * t2 + m3.x * t3, m1.y * t1 + m2.y * t2 + m3.y * t3,
              m1.z * t1 + m2.z * t2 + m3.z * t3, m1.w * t1 + m2.w * t2 + m3.w * t3};
}

/* dist2mat.cu:17-36 */
static void solve_quadratic(float A, float B, float C, float* root) {
  root[0] = -1.f;
  root[1] = -1.f;
  if (A == 0.f) {
    if (B != 0.f) {
      root[0] = -C / B;
      root[1] = root[0];
    }
  } else {
    float delta = B * B - 4.f * A * C;
    if (delta < 0.f) return;
    root[0] = (-B - sqrtf(delta)) / (2 * A);
    root[1] = (-B + sqrtf(delta)) / (2 * A);
  }
}

/* dist2mat.cu:38-69 */
static float d_cone(f3 pos, f4 m1, f4 m2) {
  f3 c1 = {m1.x, m1.y, m1.z}, c2 = {m2.x, m2.y, m2.z};
  float r1 = m1.w, r2 = m2.w;
  if (r1 > r2) {
    f3 tmp = c1;
    c1 = c2;
    c2 = tmp;
    float tf = r1;
    r1 = r2;
    r2 = tf;
  }
  f3 c21 = sub3(c1, c2);
  f3 cq2 = sub3(c2, pos);
  float A = dotf(c21, c21);
  float D = 2.f * dotf(c21, cq2);
  float F = dotf(cq2, cq2);
  float R1 = r1 - r2;
  /* the literal 4.0 promotes the radicand to double (:62) */
  float t = -(A * D - R1 * R1 * D) - sqrtf((D * D - 4.0 * A * F) * (R1 * R1 - A) * R1 * R1);
  t /= 2.f * (A * A - A * R1 * R1);
  t = clampf(t, 0.f, 1.f);
  f3 sp = {lerpf(c1.x, c2.x, t), lerpf(c1.y, c2.y, t), lerpf(c1.z, c2.z, t)};
  float lr = lerpf(r1, r2, t);
  return length3(sub3(pos, sp)) - lr;
}

/* dist2mat.cu:71-193 */
static float d_slab(f3 pos, f4 m1, f4 m2, f4 m3) {
  f3 c31 = {m1.x - m3.x, m1.y - m3.y, m1.z - m3.z};
  f3 c32 = {m2.x - m3.x, m2.y - m3.y, m2.z - m3.z};
  f3 cm3 = {m3.x - pos.x, m3.y - pos.y, m3.z - pos.z};
  float R1 = m1.w - m3.w;
  float R2 = m2.w - m3.w;
  float A = dotf(c31, c31);
  float B = 2.f * dotf(c31, c32);
  float C = dotf(c32, c32);
  float D = 2.f * dotf(c31, cm3);
  float E = 2.f * dotf(c32, cm3);
  float F = dotf(cm3, cm3);
  float t1 = -1.f, t2 = -1.f;
  if (R1 == 0.f && R2 == 0.f) {
    float denom = 4.f * A * C - B * B;
    t1 = (B * E - 2.0 * C * D) / denom;
    t2 = (B * D - 2.0 * A * E) / denom;
  } else if (R1 != 0.f && R2 == 0.f) {
    float H2 = -B / (2.f * C);
    float K2 = -E / (2.f * C);
    float W1 = powf(2.f * A + B * H2, 2.f) - 4.f * R1 * R1 * (A + B * H2 + C * H2 * H2);
    float W2 = 2.f * (2.f * A + B * H2) * (B * K2 + D) -
               4.f * R1 * R1 * (B * K2 + 2.f * C * H2 * K2 + D + E * H2);
    float W3 = powf(B * K2 + D, 2.f) - 4.f * R1 * R1 * (C * K2 * K2 + E * K2 + F);
    float root[2];
    solve_quadratic(W1, W2, W3, root);
    float t21 = H2 * root[0] + K2;
    float t22 = H2 * root[1] + K2;
    float dis = d_sphere(pos, bary_lerp(m1, m2, m3, root[0], t21));
    t1 = root[0];
    t2 = t21;
    float dis2 = d_sphere(pos, bary_lerp(m1, m2, m3, root[1], t22));
    if (dis2 < dis) {
      t1 = root[1];
      t2 = t22;
    }
  } else if (R1 == 0.f && R2 != 0.f) {
    float H1 = -B / (2.f * A);
    float K1 = -D / (2.f * A);
    float W1 = powf(2.f * C + B * H1, 2.f) - 4.f * R2 * R2 * (C + B * H1 + A * H1 * H1);
    float W2 = 2.f * (2.f * C + B * H1) * (B * K1 + E) -
               4.f * R2 * R2 * (B * K1 + 2.f * A * H1 * K1 + E + D * H1);
    float W3 = powf(B * K1 + E, 2.f) - 4.f * R2 * R2 * (A * K1 * K1 + D * K1 + F);
    float root[2];
    solve_quadratic(W1, W2, W3, root);
    float t11 = H1 * root[0] + K1;
    float t12 = H1 * root[1] + K1;
    float dis = d_sphere(pos, bary_lerp(m1, m2, m3, t11, root[0]));
    t1 = t11;
    t2 = root[0];
    float dis2 = d_sphere(pos, bary_lerp(m1, m2, m3, t12, root[1]));
    if (dis2 < dis) {
      t1 = t12;
      t2 = root[1];
    }
  } else {
    float L1 = 2.f * A * R2 - B * R1;
    float L2 = 2.f * C * R1 - B * R2;
    float L3 = E * R1 - D * R2;
    if (L1 == 0.f && L2 != 0.f) {
      t2 = -L3 / L2;
      float W1 = 4.f * A * A - 4.f * R1 * R1 * A;
      float W2 = 4.f * A * (B * t2 + D) - 4.f * R1 * R1 * (B * t2 + D);
      float W3 = powf(B * t2 + D, 2.f) - (C * t2 * t2 + E * t2 + F);
      float root[2];
      solve_quadratic(W1, W2, W3, root);
      float dis = d_sphere(pos, bary_lerp(m1, m2, m3, root[0], t2));
      t1 = root[0];
      if (d_sphere(pos, bary_lerp(m1, m2, m3, root[1], t2)) < dis) t1 = root[1];
    } else if (L1 != 0.f && L2 == 0.f) {
      t1 = L3 / L1;
      float W1 = 4.f * C * C - 4.f * R2 * R2 * C;
      float W2 = 4.f * C * (B * t1 + E) - 4.f * R2 * R2 * (B * t1 + E);
      float W3 = powf(B * t1 + E, 2.f) - (A * t1 * t1 + D * t1 + F);
      float root[2];
      solve_quadratic(W1, W2, W3, root);
      float dis = d_sphere(pos, bary_lerp(m1, m2, m3, t1, root[0]));
      t2 = root[0];
      if (d_sphere(pos, bary_lerp(m1, m2, m3, t1, root[1])) < dis) t2 = root[1];
    } else {
      float H3 = L2 / L1;
      float K3 = L3 / L1;
      float W1 = powf(2.f * C + B * H3, 2.f) - 4.f * R2 * R2 * (A * H3 * H3 + B * H3 + C);
      float W2 = 2.f * (2.f * C + B * H3) * (B * K3 + E) -
                 4.f * R2 * R2 * (2.f * A * H3 * K3 + B * K3 + D * H3 + E);
      float W3 = powf(B * K3 + E, 2.f) - 4.f * R2 * R2 * (A * K3 * K3 + D * K3 + F);
      float root[2];
      solve_quadratic(W1, W2, W3, root);
      float t11 = H3 * root[0] + K3;
      float t12 = H3 * root[1] + K3;
      float dis = d_sphere(pos, bary_lerp(m1, m2, m3, t11, root[0]));
      t1 = t11;
      t2 = root[0];
      if (d_sphere(pos, bary_lerp(m1, m2, m3, t12, root[1])) < dis) {
        t1 = t12;
        t2 = root[1];
      }
    }
  }
  if ((t1 + t2) < 1.f && t1 >= 0.f && t1 <= 1.f && t2 >= 0.f && t2 <= 1.f)
    return d_sphere(pos, bary_lerp(m1, m2, m3, t1, t2));
  float dis1 = d_cone(pos, m1, m3);
  float dis2 = d_cone(pos, m2, m3);
  float dis3 = d_cone(pos, m1, m2);
  return fminf(dis1, fminf(dis2, dis3));
}

static f3 ld3(const float* p) { return (f3){p[0], p[1], p[2]}; }
static f4 ld4(const float* p) { return (f4){p[0], p[1], p[2], p[3]}; }

float orc_distance_to_sphere(const float* pos, const float* sp) { return d_sphere(ld3(pos), ld4(sp)); }
float orc_distance_to_cone(const float* pos, const float* m1, const float* m2) {
  return d_cone(ld3(pos), ld4(m1), ld4(m2));
}
float orc_distance_to_slab(const float* pos, const float* m1, const float* m2, const float* m3) {
  return d_slab(ld3(pos), ld4(m1), ld4(m2), ld4(m3));
}

/* kernel dist2mat.cu:195-278: 32 lanes, lane l scans prims l, l+32, ... keeping its first strict
 * minimum; block min; winner = HIGHEST lane whose minimum is within 1e-10f of the block min.
 * second_best (may be NULL): smallest distance over prims other than the winner (tie flagging). */
double orc_dist2mat(const float* spheres, const float* samples, long n_samples,
                    const unsigned* offset, const unsigned* count, const int* prims, float* result,
                    int* closest_id, float* second_best, int n_threads) {
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
#pragma omp parallel for schedule(dynamic, 1024)
  for (long s = 0; s < n_samples; s++) {
    int num = (int)count[s];
    long off = offset[s];
    f3 pos = ld3(samples + 3 * s);
    float lane_min[32];
    int lane_id[32];
    for (int l = 0; l < 32; l++) {
      lane_min[l] = 1e16f;
      lane_id[l] = -1;
    }
    float best2 = 1e16f, best1 = 1e16f;
    for (int i = 0; i < num; i++) {
      const int* pr = prims + 3 * (off + i);
      float dist = 1e16f;
      if (pr[0] == -1 && pr[1] == -1)
        dist = d_sphere(pos, ld4(spheres + 4 * (long)pr[2]));
      else if (pr[0] == -1 && pr[1] != -1)
        dist = d_cone(pos, ld4(spheres + 4 * (long)pr[1]), ld4(spheres + 4 * (long)pr[2]));
      else if (pr[0] != -1)
        dist = d_slab(pos, ld4(spheres + 4 * (long)pr[0]), ld4(spheres + 4 * (long)pr[1]),
                      ld4(spheres + 4 * (long)pr[2]));
      int l = i & 31;
      if (dist < lane_min[l]) {
        lane_min[l] = fminf(lane_min[l], dist);
        lane_id[l] = i;
      }
      if (dist < best1) {
        best2 = best1;
        best1 = dist;
      } else if (dist < best2)
        best2 = dist;
    }
    float red = lane_min[0];
    for (int l = 1; l < 32; l++)
      if (lane_min[l] < red) red = lane_min[l];
    result[s] = red;
    int cid = closest_id[s];
    for (int l = 0; l < 32; l++)
      if (fabsf(lane_min[l] - red) < 1e-10f) cid = lane_id[l];
    closest_id[s] = cid;
    if (second_best) second_best[s] = best2;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
