// K5: dist2mat -- closest medial primitive (sphere / cone / slab) per sample.
//
// Replaces ClosestDistanceToLocalMat + compute_closest_dist2mat (reference
// src/dist2mat/dist2mat.cu:195-315): there, one 32-thread BLOCK per sample, a cub::BlockReduce and a
// serial lane-0 scan for the argmin, and 7 blocking H2D + 2 D2H per call.  Here: one WARP per sample
// in a persistent grid, lanes over the sample's primitive list (same lane <-> primitive
// assignment i = lane, lane+32, ... so the reference's tie rule is reproduced exactly), REDUX
// warp-min on order-preserving integer keys, ballot for the "highest lane within 1e-10" winner.
//
// Arithmetic: this TU is compiled with -fmad=false and the default IEEE sqrt/div, and keeps the
// reference's three double-promoted spots (dist2mat.cu:62, :89-90), so distances are bit-identical
// to the reference's own functions built for the host, except where the reference calls
// powf(x, 2.f) (glibc's powf is not always correctly rounded; x*x is).
#include "mb_internal.h"

namespace {

struct f3 {
  float x, y, z;
};

__device__ __forceinline__ float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 sub3(f3 a, f3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ float len3(f3 v) { return sqrtf(dot3(v, v)); }
// clamp(t, 0, 1) = fmaxf(0, fminf(t, 1)) (cuda_helper_math.h:932-934) as the reference's DEVICE build evaluates it:
// nvcc lowers it to a saturate modifier, and .sat maps a NaN t to +0 (IEEE fminf / fmaxf -- the reference's host
// compile -- would give 1).  t is NaN for nested spheres (dist2mat.cu:61): the reference then returns the distance to
// the larger sphere.  Written with explicit compares so that the semantics do not depend on the compiler's lowering.
__device__ __forceinline__ float clampf(float f, float a, float b) {
  const float m = (f < b) ? f : b;         // NaN f -> b here ...
  return (f != f) ? a : ((a > m) ? a : m);  // ... and a in the end, like .sat
}
// lerp(a,b,t) = b + t*(a-b) (cuda_helper_math.h:911-913)
__device__ __forceinline__ float lerpf(float a, float b, float t) { return b + t * (a - b); }
__device__ __forceinline__ float sq(float x) { return powf(x, 2.f); }  // literally, like dist2mat.cu:96,100,...

// dist2mat.cu:5-8
__device__ __forceinline__ float d_sphere(f3 p, float4 sp) {
  return len3(sub3(p, f3{sp.x, sp.y, sp.z})) - sp.w;
}

// dist2mat.cu:10-15
__device__ __forceinline__ float4 bary_lerp(float4 m1, float4 m2, float4 m3, float t1, float t2) {
  const float t3 = 1.f - t1 - t2;
  return make_float4(m1.x * t1 + m2.x * t2 + m3.x * t3, m1.y * t1 + m2.y * t2 + m3.y * t3,
                     m1.z * t1 + m2.z * t2 + m3.z * t3, m1.w * t1 + m2.w * t2 + m3.w * t3);
}

// dist2mat.cu:17-36
__device__ __forceinline__ void solve_quadratic(float A, float B, float C, float& r0, float& r1) {
  r0 = -1.f;
  r1 = -1.f;
  if (A == 0.f) {
    if (B != 0.f) {
      r0 = -C / B;
      r1 = r0;
    }
  } else {
    const float delta = B * B - 4.f * A * C;
    if (delta < 0.f) return;
    const float sd = sqrtf(delta);
    r0 = (-B - sd) / (2 * A);
    r1 = (-B + sd) / (2 * A);
  }
}

// dist2mat.cu:38-69
__device__ float d_cone(f3 pos, float4 m1, float4 m2) {
  f3 c1{m1.x, m1.y, m1.z}, c2{m2.x, m2.y, m2.z};
  float r1 = m1.w, r2 = m2.w;
  if (r1 > r2) {
    f3 tc = c1;
    c1 = c2;
    c2 = tc;
    float tr = r1;
    r1 = r2;
    r2 = tr;
  }
  const f3 c21 = sub3(c1, c2);
  const f3 cq2 = sub3(c2, pos);
  const float A = dot3(c21, c21);
  const float D = 2.f * dot3(c21, cq2);
  const float F = dot3(cq2, cq2);
  const float R1 = r1 - r2;
  // the literal 4.0 promotes the radicand to double (:62)
  const double rad = ((double)(D * D) - 4.0 * (double)A * (double)F) * (double)(R1 * R1 - A) *
                     (double)R1 * (double)R1;
  float t = -(A * D - R1 * R1 * D) - sqrtf((float)rad);
  t /= 2.f * (A * A - A * R1 * R1);
  t = clampf(t, 0.f, 1.f);
  const f3 sp{lerpf(c1.x, c2.x, t), lerpf(c1.y, c2.y, t), lerpf(c1.z, c2.z, t)};
  const float lr = lerpf(r1, r2, t);
  return len3(sub3(pos, sp)) - lr;
}

// dist2mat.cu:71-193, cut in two: the interior solve (sets `inside`) and the boundary-cone fallback, so
// that the queue kernel can run each part with converged lanes
__device__ float d_slab_main(f3 pos, float4 m1, float4 m2, float4 m3, bool& inside) {
  const f3 c31{m1.x - m3.x, m1.y - m3.y, m1.z - m3.z};
  const f3 c32{m2.x - m3.x, m2.y - m3.y, m2.z - m3.z};
  const f3 cm3{m3.x - pos.x, m3.y - pos.y, m3.z - pos.z};
  const float R1 = m1.w - m3.w;
  const float R2 = m2.w - m3.w;
  const float A = dot3(c31, c31);
  const float B = 2.f * dot3(c31, c32);
  const float C = dot3(c32, c32);
  const float D = 2.f * dot3(c31, cm3);
  const float E = 2.f * dot3(c32, cm3);
  const float F = dot3(cm3, cm3);
  float t1 = -1.f, t2 = -1.f;
  if (R1 == 0.f && R2 == 0.f) {
    const float denom = 4.f * A * C - B * B;
    t1 = (float)(((double)(B * E) - 2.0 * (double)C * (double)D) / (double)denom);
    t2 = (float)(((double)(B * D) - 2.0 * (double)A * (double)E) / (double)denom);
  } else if (R1 != 0.f && R2 == 0.f) {
    const float H2 = -B / (2.f * C);
    const float K2 = -E / (2.f * C);
    const float W1 = sq(2.f * A + B * H2) - 4.f * R1 * R1 * (A + B * H2 + C * H2 * H2);
    const float W2 = 2.f * (2.f * A + B * H2) * (B * K2 + D) -
                     4.f * R1 * R1 * (B * K2 + 2.f * C * H2 * K2 + D + E * H2);
    const float W3 = sq(B * K2 + D) - 4.f * R1 * R1 * (C * K2 * K2 + E * K2 + F);
    float r0, r1;
    solve_quadratic(W1, W2, W3, r0, r1);
    const float t21 = H2 * r0 + K2;
    const float t22 = H2 * r1 + K2;
    const float dis = d_sphere(pos, bary_lerp(m1, m2, m3, r0, t21));
    t1 = r0;
    t2 = t21;
    const float dis2 = d_sphere(pos, bary_lerp(m1, m2, m3, r1, t22));
    if (dis2 < dis) {
      t1 = r1;
      t2 = t22;
    }
  } else if (R1 == 0.f && R2 != 0.f) {
    const float H1 = -B / (2.f * A);
    const float K1 = -D / (2.f * A);
    const float W1 = sq(2.f * C + B * H1) - 4.f * R2 * R2 * (C + B * H1 + A * H1 * H1);
    const float W2 = 2.f * (2.f * C + B * H1) * (B * K1 + E) -
                     4.f * R2 * R2 * (B * K1 + 2.f * A * H1 * K1 + E + D * H1);
    const float W3 = sq(B * K1 + E) - 4.f * R2 * R2 * (A * K1 * K1 + D * K1 + F);
    float r0, r1;
    solve_quadratic(W1, W2, W3, r0, r1);
    const float t11 = H1 * r0 + K1;
    const float t12 = H1 * r1 + K1;
    const float dis = d_sphere(pos, bary_lerp(m1, m2, m3, t11, r0));
    t1 = t11;
    t2 = r0;
    const float dis2 = d_sphere(pos, bary_lerp(m1, m2, m3, t12, r1));
    if (dis2 < dis) {
      t1 = t12;
      t2 = r1;
    }
  } else {
    const float L1 = 2.f * A * R2 - B * R1;
    const float L2 = 2.f * C * R1 - B * R2;
    const float L3 = E * R1 - D * R2;
    if (L1 == 0.f && L2 != 0.f) {
      t2 = -L3 / L2;
      const float W1 = 4.f * A * A - 4.f * R1 * R1 * A;
      const float W2 = 4.f * A * (B * t2 + D) - 4.f * R1 * R1 * (B * t2 + D);
      const float W3 = sq(B * t2 + D) - (C * t2 * t2 + E * t2 + F);
      float r0, r1;
      solve_quadratic(W1, W2, W3, r0, r1);
      const float dis = d_sphere(pos, bary_lerp(m1, m2, m3, r0, t2));
      t1 = r0;
      if (d_sphere(pos, bary_lerp(m1, m2, m3, r1, t2)) < dis) t1 = r1;
    } else if (L1 != 0.f && L2 == 0.f) {
      t1 = L3 / L1;
      const float W1 = 4.f * C * C - 4.f * R2 * R2 * C;
      const float W2 = 4.f * C * (B * t1 + E) - 4.f * R2 * R2 * (B * t1 + E);
      const float W3 = sq(B * t1 + E) - (A * t1 * t1 + D * t1 + F);
      float r0, r1;
      solve_quadratic(W1, W2, W3, r0, r1);
      const float dis = d_sphere(pos, bary_lerp(m1, m2, m3, t1, r0));
      t2 = r0;
      if (d_sphere(pos, bary_lerp(m1, m2, m3, t1, r1)) < dis) t2 = r1;
    } else {
      const float H3 = L2 / L1;
      const float K3 = L3 / L1;
      const float W1 = sq(2.f * C + B * H3) - 4.f * R2 * R2 * (A * H3 * H3 + B * H3 + C);
      const float W2 = 2.f * (2.f * C + B * H3) * (B * K3 + E) -
                       4.f * R2 * R2 * (2.f * A * H3 * K3 + B * K3 + D * H3 + E);
      const float W3 = sq(B * K3 + E) - 4.f * R2 * R2 * (A * K3 * K3 + D * K3 + F);
      float r0, r1;
      solve_quadratic(W1, W2, W3, r0, r1);
      const float t11 = H3 * r0 + K3;
      const float t12 = H3 * r1 + K3;
      const float dis = d_sphere(pos, bary_lerp(m1, m2, m3, t11, r0));
      t1 = t11;
      t2 = r0;
      if (d_sphere(pos, bary_lerp(m1, m2, m3, t12, r1)) < dis) {
        t1 = t12;
        t2 = r1;
      }
    }
  }
  if ((t1 + t2) < 1.f && t1 >= 0.f && t1 <= 1.f && t2 >= 0.f && t2 <= 1.f) {
    inside = true;
    return d_sphere(pos, bary_lerp(m1, m2, m3, t1, t2));
  }
  inside = false;
  return 0.f;
}

// the slab's fallback when the foot point leaves the triangle: its three boundary cones (:186-192)
__device__ __forceinline__ float d_slab_boundary(f3 pos, float4 m1, float4 m2, float4 m3) {
  const float dis1 = d_cone(pos, m1, m3);
  const float dis2 = d_cone(pos, m2, m3);
  const float dis3 = d_cone(pos, m1, m2);
  return fminf(dis1, fminf(dis2, dis3));
}

__device__ __forceinline__ float d_slab(f3 pos, float4 m1, float4 m2, float4 m3) {
  bool inside;
  const float d = d_slab_main(pos, m1, m2, m3, inside);
  return inside ? d : d_slab_boundary(pos, m1, m2, m3);
}

// order-preserving float -> uint key (for REDUX min)
__device__ __forceinline__ unsigned fkey(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float funkey(unsigned k) {
  const unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

// The flagged-tie class: the two best distances of a sample differ by less than 1e-6 relative to the larger of the
// distances and the sample's own coordinates.  A signed distance |p - c| - r is a cancelling difference of terms of
// the coordinates' size, so two correct evaluations of it (FMA-contracted or not) differ by ~1 ulp OF THE COORDINATES,
// however small the distance itself is; two candidates closer than that cannot be ranked.
__device__ __forceinline__ unsigned char tie_flag(float best, float second, const float* __restrict__ p) {
  const float scale = fmaxf(fmaxf(fabsf(best), fabsf(second)), fmaxf(fabsf(p[0]), fmaxf(fabsf(p[1]), fabsf(p[2]))));
  return (second - best) <= 1e-6f * scale && second < 1e15f;
}

__global__ void __launch_bounds__(256) k_dist2mat(const float* __restrict__ samples,
                                                  const float4* __restrict__ spheres,
                                                  const int* __restrict__ prims,
                                                  const unsigned* __restrict__ offsets,
                                                  const unsigned* __restrict__ counts, int n_samples,
                                                  float* __restrict__ result, int* __restrict__ closest_id,
                                                  unsigned char* __restrict__ tie) {
  const int lane = threadIdx.x & 31;
  const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
  for (int smp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; smp < n_samples; smp += warps_per_grid) {
    const int num_prim = (int)counts[smp];
    const long long off = offsets[smp];
    const f3 pos{samples[3 * (size_t)smp], samples[3 * (size_t)smp + 1], samples[3 * (size_t)smp + 2]};
    float best = 1e16f, second = 1e16f;
    int best_id = -1;
    for (int i = lane; i < num_prim; i += 32) {
      const int* pr = prims + 3 * (off + i);
      const int px = pr[0], py = pr[1], pz = pr[2];
      float dist = 1e16f;
      if (px == -1 && py == -1)
        dist = d_sphere(pos, spheres[pz]);
      else if (px == -1)
        dist = d_cone(pos, spheres[py], spheres[pz]);
      else
        dist = d_slab(pos, spheres[px], spheres[py], spheres[pz]);
      if (dist < best) {  // first strict minimum over the lane's stride (:248-251)
        second = best;
        best = fminf(best, dist);
        best_id = i;
      } else if (dist < second)
        second = dist;
    }
    const float red = funkey(__reduce_min_sync(0xffffffffu, fkey(best)));
    // winner = HIGHEST lane whose minimum is within 1e-10 of the block minimum (:269-276)
    const unsigned eq = __ballot_sync(0xffffffffu, fabsf(best - red) < 1e-10f);
    const int win = 31 - __clz(eq);
    const int win_id = __shfl_sync(0xffffffffu, best_id, win);
    // flagged-tie class: best distance among all OTHER primitives within 1e-6 relative
    const float other = (lane == win) ? second : best;
    const float sec = funkey(__reduce_min_sync(0xffffffffu, fkey(other)));
    if (lane == 0) {
      result[smp] = red;
      closest_id[smp] = win_id;
      if (tie) tie[smp] = tie_flag(red, sec, samples + 3 * (size_t)smp);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Queue-compacted variant (the one launched).  The warp-per-sample kernel above runs ~13 of 32 lanes
// per instruction: a sample's list (~22 primitives) does not fill the warp and its spheres, cones and
// slabs take different code paths, as does the slab's boundary-cone fallback.  Here a warp takes a
// BATCH of up to 32 consecutive samples and
//   A  classifies their primitives (spheres are evaluated on the spot), pushing cones to the front and
//      slabs to the back of a shared-memory queue; entry = (sample-in-batch << 16 | index in its list)
//   B  evaluates the cone queue, 32 converged lanes at a time
//   C  evaluates the slab interior solve the same way; slabs whose foot point leaves the triangle are
//      re-queued in place
//   D  evaluates the boundary-cone fallbacks, converged
//   E  one LANE per sample reduces that sample's distances, reproducing the reference's block-level
//      tie rule (lane l keeps the first strict minimum over i = l, l+32, ...; the winner is the highest
//      lane within 1e-10 of the minimum, dist2mat.cu:248-276) from the stored per-primitive distances.
// Every distance is computed by the same device functions as above: results are bit-identical.
// ---------------------------------------------------------------------------------------------
template <int CAP>
__device__ __forceinline__ void d2m_sample_direct(int smp, int lane, const float* __restrict__ samples,
                                                  const float4* __restrict__ spheres, const int* __restrict__ prims,
                                                  const unsigned* __restrict__ offsets, const unsigned* __restrict__ counts,
                                                  float* __restrict__ result, int* __restrict__ closest_id,
                                                  unsigned char* __restrict__ tie) {
  const int num_prim = (int)counts[smp];
  const long long off = offsets[smp];
  const f3 pos{samples[3 * (size_t)smp], samples[3 * (size_t)smp + 1], samples[3 * (size_t)smp + 2]};
  float best = 1e16f, second = 1e16f;
  int best_id = -1;
  for (int i = lane; i < num_prim; i += 32) {
    const int* pr = prims + 3 * (off + i);
    const int px = pr[0], py = pr[1], pz = pr[2];
    float dist;
    if (px == -1 && py == -1)
      dist = d_sphere(pos, spheres[pz]);
    else if (px == -1)
      dist = d_cone(pos, spheres[py], spheres[pz]);
    else
      dist = d_slab(pos, spheres[px], spheres[py], spheres[pz]);
    if (dist < best) {
      second = best;
      best = fminf(best, dist);
      best_id = i;
    } else if (dist < second)
      second = dist;
  }
  const float red = funkey(__reduce_min_sync(0xffffffffu, fkey(best)));
  const unsigned eq = __ballot_sync(0xffffffffu, fabsf(best - red) < 1e-10f);
  const int win = 31 - __clz(eq);
  const int win_id = __shfl_sync(0xffffffffu, best_id, win);
  const float other = (lane == win) ? second : best;
  const float sec = funkey(__reduce_min_sync(0xffffffffu, fkey(other)));
  if (lane == 0) {
    result[smp] = red;
    closest_id[smp] = win_id;
    if (tie) tie[smp] = tie_flag(red, sec, samples + 3 * (size_t)smp);
  }
}

template <int CAP, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) k_dist2mat_q(const float* __restrict__ samples,
                                                          const float4* __restrict__ spheres,
                                                          const int* __restrict__ prims,
                                                          const unsigned* __restrict__ offsets,
                                                          const unsigned* __restrict__ counts, int n_samples,
                                                          float* __restrict__ result, int* __restrict__ closest_id,
                                                          unsigned char* __restrict__ tie) {
  extern __shared__ __align__(16) unsigned char d2m_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float* sd = reinterpret_cast<float*>(d2m_smem) + (size_t)wib * CAP;                       // distance per slot
  unsigned* sq = reinterpret_cast<unsigned*>(d2m_smem) + (size_t)WARPS * CAP + (size_t)wib * CAP;  // queue
  const unsigned lt = (1u << lane) - 1u;
  const int n_batches = (n_samples + 31) >> 5;
  for (int batch = blockIdx.x * WARPS + wib; batch < n_batches; batch += gridDim.x * WARPS) {
    const int smp = batch * 32 + lane;
    const bool have = smp < n_samples;
    const int cnt = have ? (int)counts[smp] : 0;
    const unsigned off = have ? offsets[smp] : 0u;
    f3 pos{0.f, 0.f, 0.f};
    if (have) pos = f3{samples[3 * (size_t)smp], samples[3 * (size_t)smp + 1], samples[3 * (size_t)smp + 2]};
    // inclusive prefix of the list lengths
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int excl = incl - cnt;
    const int n_have = min(32, n_samples - batch * 32);
    int s0 = 0;
    while (s0 < n_have) {
      // sub-batch [s0, s1): as many samples as fit the CAP slots
      const int ex0 = __shfl_sync(0xffffffffu, excl, s0);
      const unsigned fits = __ballot_sync(0xffffffffu, lane >= s0 && lane < n_have && incl - ex0 <= CAP);
      int s1 = s0 + __popc(fits);
      if (s1 == s0) {  // a single list longer than CAP: the whole warp walks it directly
        d2m_sample_direct<CAP>(batch * 32 + s0, lane, samples, spheres, prims, offsets, counts, result, closest_id, tie);
        s0++;
        continue;
      }
      const int base = excl - ex0;  // first slot of this lane's sample (valid for lanes in [s0, s1))
      // ---- A: classify; spheres on the spot -----------------------------------------------------
      int n_cone = 0, n_slab = 0;
      for (int s = s0; s < s1; s++) {
        const int c = __shfl_sync(0xffffffffu, cnt, s);
        const unsigned o = __shfl_sync(0xffffffffu, off, s);
        const int bs = __shfl_sync(0xffffffffu, base, s);
        const f3 p{__shfl_sync(0xffffffffu, pos.x, s), __shfl_sync(0xffffffffu, pos.y, s),
                   __shfl_sync(0xffffffffu, pos.z, s)};
        for (int i0 = 0; i0 < c; i0 += 32) {
          const int i = i0 + lane;
          int kind = 0;  // 0 none / sphere (done), 1 cone, 2 slab
          if (i < c) {
            const int* pr = prims + 3 * ((size_t)o + i);
            const int px = pr[0], py = pr[1], pz = pr[2];
            if (px == -1 && py == -1)
              sd[bs + i] = d_sphere(p, spheres[pz]);
            else
              kind = (px == -1) ? 1 : 2;
          }
          const unsigned mc = __ballot_sync(0xffffffffu, kind == 1), ms = __ballot_sync(0xffffffffu, kind == 2);
          const unsigned e = ((unsigned)s << 16) | (unsigned)i;
          if (kind == 1) sq[n_cone + __popc(mc & lt)] = e;
          if (kind == 2) sq[CAP - 1 - (n_slab + __popc(ms & lt))] = e;
          n_cone += __popc(mc);
          n_slab += __popc(ms);
        }
      }
      __syncwarp();
      // ---- B: cones -----------------------------------------------------------------------------
      for (int n0 = 0; n0 < n_cone; n0 += 32) {
        const int n = n0 + lane;
        const bool act = n < n_cone;
        const unsigned e = sq[act ? n : 0];
        const int sm = (int)(e >> 16), i = (int)(e & 0xffffu);
        const unsigned o = __shfl_sync(0xffffffffu, off, sm);
        const int bs = __shfl_sync(0xffffffffu, base, sm);
        const f3 p{__shfl_sync(0xffffffffu, pos.x, sm), __shfl_sync(0xffffffffu, pos.y, sm),
                   __shfl_sync(0xffffffffu, pos.z, sm)};
        if (act) {
          const int* pr = prims + 3 * ((size_t)o + i);
          sd[bs + i] = d_cone(p, spheres[pr[1]], spheres[pr[2]]);
        }
      }
      // ---- C: slab interior solve; boundary cases re-queued in place ----------------------------
      int n_fb = 0;
      for (int n0 = 0; n0 < n_slab; n0 += 32) {
        const int n = n0 + lane;
        const bool act = n < n_slab;
        const unsigned e = sq[CAP - 1 - (act ? n : 0)];
        const int sm = (int)(e >> 16), i = (int)(e & 0xffffu);
        const unsigned o = __shfl_sync(0xffffffffu, off, sm);
        const int bs = __shfl_sync(0xffffffffu, base, sm);
        const f3 p{__shfl_sync(0xffffffffu, pos.x, sm), __shfl_sync(0xffffffffu, pos.y, sm),
                   __shfl_sync(0xffffffffu, pos.z, sm)};
        bool inside = true;
        if (act) {
          const int* pr = prims + 3 * ((size_t)o + i);
          const float d = d_slab_main(p, spheres[pr[0]], spheres[pr[1]], spheres[pr[2]], inside);
          if (inside) sd[bs + i] = d;
        }
        const unsigned mf = __ballot_sync(0xffffffffu, !inside);  // also orders this round's queue reads
        if (!inside) sq[CAP - 1 - (n_fb + __popc(mf & lt))] = e;  // n_fb + rank <= n: never ahead of a reader
        n_fb += __popc(mf);
      }
      __syncwarp();
      // ---- D: boundary cones of the re-queued slabs ---------------------------------------------
      for (int n0 = 0; n0 < n_fb; n0 += 32) {
        const int n = n0 + lane;
        const bool act = n < n_fb;
        const unsigned e = sq[CAP - 1 - (act ? n : 0)];
        const int sm = (int)(e >> 16), i = (int)(e & 0xffffu);
        const unsigned o = __shfl_sync(0xffffffffu, off, sm);
        const int bs = __shfl_sync(0xffffffffu, base, sm);
        const f3 p{__shfl_sync(0xffffffffu, pos.x, sm), __shfl_sync(0xffffffffu, pos.y, sm),
                   __shfl_sync(0xffffffffu, pos.z, sm)};
        if (act) {
          const int* pr = prims + 3 * ((size_t)o + i);
          sd[bs + i] = d_slab_boundary(p, spheres[pr[0]], spheres[pr[1]], spheres[pr[2]]);
        }
      }
      __syncwarp();
      // ---- E: one lane per sample, the reference's tie rule from the stored distances -----------
      if (lane >= s0 && lane < s1) {
        const float* d = sd + base;
        float m = 1e16f;
        for (int i = 0; i < cnt; i++) {
          const float v = d[i];
          if (v < m) m = v;
        }
        int id = -1;
        float sec = 1e16f;
        if (m < 1e16f) {
          // winning lane class L = highest (i mod 32) holding a value within 1e-10 of the minimum
          int L = -1;
          for (int i = 0; i < cnt; i++)
            if (fabsf(d[i] - m) < 1e-10f) L = max(L, i & 31);
          // that lane's first strict minimum over i = L, L+32, ...
          float b = 1e16f;
          for (int i = L; i < cnt; i += 32) {
            const float v = d[i];
            if (v < b) {
              b = v;
              id = i;
            }
          }
          for (int i = 0; i < cnt; i++) {
            const float v = d[i];
            if (i != id && v < sec) sec = v;
          }
        }
        result[smp] = m;
        closest_id[smp] = id;
        if (tie) tie[smp] = tie_flag(m, sec, samples + 3 * (size_t)smp);
      }
      __syncwarp();
      s0 = s1;
    }
  }
}

}  // namespace

void d2m_upload(mb_ctx* ctx, const float* spheres, int n_sph, const float* samples, int n_samples,
                const unsigned* offset, const unsigned* count, const int* prims, long n_prims) {
  D2MDev& D = ctx->d2m;
  cudaStream_t s = ctx->stream;
  D.spheres.reserve(n_sph);
  D.samples.reserve(3 * (size_t)n_samples);
  D.offset.reserve(n_samples);
  D.count.reserve(n_samples);
  D.prims.reserve(3 * (size_t)n_prims);
  D.result.reserve(n_samples);
  D.closest.reserve(n_samples);
  D.tie.reserve(n_samples);
  MB_CUDA(cudaMemcpyAsync(D.spheres.p, spheres, sizeof(float4) * (size_t)n_sph, cudaMemcpyHostToDevice, s));
  MB_CUDA(cudaMemcpyAsync(D.samples.p, samples, sizeof(float) * 3 * (size_t)n_samples, cudaMemcpyHostToDevice, s));
  MB_CUDA(cudaMemcpyAsync(D.offset.p, offset, sizeof(unsigned) * (size_t)n_samples, cudaMemcpyHostToDevice, s));
  MB_CUDA(cudaMemcpyAsync(D.count.p, count, sizeof(unsigned) * (size_t)n_samples, cudaMemcpyHostToDevice, s));
  MB_CUDA(cudaMemcpyAsync(D.prims.p, prims, sizeof(int) * 3 * (size_t)n_prims, cudaMemcpyHostToDevice, s));
  D.n_sph = n_sph;
  D.n_samples = n_samples;
  D.n_prims = n_prims;
  // header contract: the caller's arrays may be freed on return (copies from pinned memory are truly asynchronous)
  MB_CUDA(cudaStreamSynchronize(s));
}

void d2m_run(mb_ctx* ctx, float* kernel_ms) {
  D2MDev& D = ctx->d2m;
  MB_REQUIRE(D.n_samples >= 0 && D.spheres.p, MB_ERR_STATE, "mb_dist2mat_upload must be called first");
  cudaStream_t s = ctx->stream;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (kernel_ms) {
    MB_CUDA(cudaEventCreate(&e0));
    MB_CUDA(cudaEventCreate(&e1));
    MB_CUDA(cudaEventRecord(e0, s));
  }
  if (D.n_samples > 0) {
    ctx->n_launches++;
    if (ctx->d2m_variant == 1) {  // warp-per-sample kernel (kept for A/B measurements)
      long long blocks = ((long long)D.n_samples + 7) / 8;
      blocks = std::min<long long>(blocks, (long long)ctx->sm_count * 32);
      k_dist2mat<<<(unsigned)blocks, 256, 0, s>>>(D.samples.p, D.spheres.p, D.prims.p, D.offset.p, D.count.p,
                                                 D.n_samples, D.result.p, D.closest.p, D.tie.p);
    } else {
      constexpr int CAP = 768, WARPS = 8;
      const size_t smem = (size_t)WARPS * CAP * 8;
      static bool attr_set_dev[64] = {false};
      bool& attr_set = attr_set_dev[ctx->device & 63];  // function attributes are per device
      if (!attr_set) {
        MB_CUDA(cudaFuncSetAttribute(k_dist2mat_q<CAP, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
      }
      static int per_sm_dev[64] = {0};
      int& per_sm = per_sm_dev[ctx->device & 63];
      if (per_sm < 1) {
        MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dist2mat_q<CAP, WARPS>, 32 * WARPS, smem));
        if (per_sm < 1) per_sm = 1;
      }
      const long long batches = ((long long)D.n_samples + 31) / 32;
      long long blocks = (batches + WARPS - 1) / WARPS;
      blocks = std::min<long long>(blocks, (long long)ctx->sm_count * per_sm);  // persistent: every SM full
      k_dist2mat_q<CAP, WARPS><<<(unsigned)blocks, 32 * WARPS, smem, s>>>(D.samples.p, D.spheres.p, D.prims.p, D.offset.p,
                                                                         D.count.p, D.n_samples, D.result.p,
                                                                         D.closest.p, D.tie.p);
    }
    MB_CUDA(cudaGetLastError());
  }
  if (kernel_ms) {
    MB_CUDA(cudaEventRecord(e1, s));
    MB_CUDA(cudaEventSynchronize(e1));
    MB_CUDA(cudaEventElapsedTime(kernel_ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
}

void d2m_fetch(mb_ctx* ctx, float* result, int* closest_id, unsigned char* tie_flag) {
  D2MDev& D = ctx->d2m;
  cudaStream_t s = ctx->stream;
  if (result) MB_CUDA(cudaMemcpyAsync(result, D.result.p, sizeof(float) * (size_t)D.n_samples, cudaMemcpyDeviceToHost, s));
  if (closest_id) MB_CUDA(cudaMemcpyAsync(closest_id, D.closest.p, sizeof(int) * (size_t)D.n_samples, cudaMemcpyDeviceToHost, s));
  if (tie_flag) MB_CUDA(cudaMemcpyAsync(tie_flag, D.tie.p, (size_t)D.n_samples, cudaMemcpyDeviceToHost, s));
  MB_CUDA(cudaStreamSynchronize(s));
}
