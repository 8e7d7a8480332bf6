// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// The reference's src/dist2mat/dist2mat.cu compiled IN PLACE (found via -I, never copied).
// Its distance functions are __host__ __device__ (dist2mat.cu:5-193), so this TU can call them
// on the host (IEEE, -ffp-contract=off) as the CPU oracle / CPU baseline, and on a GPU box it
// can also run the reference's own kernel through compute_closest_dist2mat (dist2mat.cu:280-315).
//
// ONE deliberate difference between the host compile of this file and the reference's device build, found by running
// both on the B200 (tests/test_gpu_reference_build.py): clamp(t, 0.f, 1.f) = fmaxf(0, fminf(t, 1))
// (cuda_helper_math.h:932-934) is compiled by nvcc for the device into a SATURATE modifier, and .sat maps NaN to +0,
// whereas IEEE fminf/fmaxf on the host map it to 1.  t IS NaN whenever the two spheres of a cone are nested
// ((D*D - 4AF)(R1*R1 - A) R1*R1 < 0, dist2mat.cu:61): the device build then returns the distance to the LARGER
// sphere (the true envelope), the host build the distance to the smaller one.  What LibMAT runs is the device
// build, so the host functions below are made to follow it: inside the reference's source, clamp is routed to
// ref_clamp_sat (device semantics).  Nothing else differs beyond last-bit rounding (FMA contraction, libm powf).
#include <cub/cub.cuh>  // everything dist2mat.cu includes is pulled in first (include-guarded) ...
#include "dist2mat.h"   // ... so that the macro below only touches the body of dist2mat.cu
static inline __host__ __device__ float ref_clamp_sat(float f, float a, float b) {
#ifdef __CUDA_ARCH__
  return fmaxf(a, fminf(f, b));  // the device compiler's own lowering (.sat)
#else
  return (f != f) ? a : fmaxf(a, fminf(f, b));
#endif
}
#define clamp(f, a, b) ref_clamp_sat(f, a, b)
#include "dist2mat.cu"
#undef clamp

#include <omp.h>
#include <time.h>

// per-(sample, primitive) evaluation of the reference's distance functions ON THE DEVICE (its own device build):
// lets a test compare device and host builds of the reference primitive by primitive
__global__ void ref_eval_prims_kernel(const float3* pos, const float4* spheres, const int3* prims, int n, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int3 prim = prims[i];
  float dist = 1e16f;
  if (prim.x == -1 && prim.y == -1)
    dist = distance_to_sphere(pos[i], spheres[prim.z]);
  else if (prim.x == -1 && prim.y != -1)
    dist = compute_distance_to_cone(pos[i], spheres[prim.y], spheres[prim.z]);
  else if (prim.x != -1)
    dist = compute_distance_to_slab(pos[i], spheres[prim.x], spheres[prim.y], spheres[prim.z]);
  out[i] = dist;
}

extern "C" {

// n (position, primitive) pairs evaluated by the reference's device code; returns 0, <0 without a device
int ref_d2m_eval_prims_gpu(const float* spheres, int n_sph, const float* pos, const int* prims, int n, float* out) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return -1;
  float4* d_s; float3* d_p; int3* d_q; float* d_o;
  cudaMalloc(&d_s, sizeof(float4) * (size_t)n_sph);
  cudaMalloc(&d_p, sizeof(float3) * (size_t)n);
  cudaMalloc(&d_q, sizeof(int3) * (size_t)n);
  cudaMalloc(&d_o, sizeof(float) * (size_t)n);
  cudaMemcpy(d_s, spheres, sizeof(float4) * (size_t)n_sph, cudaMemcpyHostToDevice);
  cudaMemcpy(d_p, pos, sizeof(float3) * (size_t)n, cudaMemcpyHostToDevice);
  cudaMemcpy(d_q, prims, sizeof(int3) * (size_t)n, cudaMemcpyHostToDevice);
  ref_eval_prims_kernel<<<(n + 127) / 128, 128>>>(d_p, d_s, d_q, n, d_o);
  const cudaError_t e = cudaMemcpy(out, d_o, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost);
  cudaFree(d_s); cudaFree(d_p); cudaFree(d_q); cudaFree(d_o);
  return e == cudaSuccess ? 0 : -2;
}


float ref_d2m_sphere(const float* p, const float* s) {
  return distance_to_sphere(make_float3(p[0], p[1], p[2]), make_float4(s[0], s[1], s[2], s[3]));
}
float ref_d2m_cone(const float* p, const float* a, const float* b) {
  return compute_distance_to_cone(make_float3(p[0], p[1], p[2]), make_float4(a[0], a[1], a[2], a[3]),
                                  make_float4(b[0], b[1], b[2], b[3]));
}
float ref_d2m_slab(const float* p, const float* a, const float* b, const float* c) {
  return compute_distance_to_slab(make_float3(p[0], p[1], p[2]), make_float4(a[0], a[1], a[2], a[3]),
                                  make_float4(b[0], b[1], b[2], b[3]),
                                  make_float4(c[0], c[1], c[2], c[3]));
}

// Host loop over samples with the reference's distance functions and the kernel's lane/tie rule
// (dist2mat.cu:224-276).  Returns seconds in the loop.
double ref_d2m_run_host(const float* spheres, const float* samples, long n_samples,
                        const unsigned* offset, const unsigned* count, const int* prims,
                        float* result, int* closest_id, int n_threads) {
  if (n_threads > 0) omp_set_num_threads(n_threads);
  const float4* sph = reinterpret_cast<const float4*>(spheres);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
#pragma omp parallel for schedule(dynamic, 1024)
  for (long s = 0; s < n_samples; s++) {
    const int num_prim = (int)count[s];
    const long off = offset[s];
    const float3 pos = make_float3(samples[3 * s], samples[3 * s + 1], samples[3 * s + 2]);
    float min_distance[kWarpSize];
    int min_id[kWarpSize];
    for (int tid = 0; tid < kWarpSize; tid++) {
      float local_closest_dist = 1e16f;
      int local_closest_id = -1;
      for (int i = tid; i < num_prim && tid < num_prim; i += kWarpSize) {
        const int* pr = prims + 3 * (off + i);
        int3 prim = make_int3(pr[0], pr[1], pr[2]);
        float dist = 1e16f;
        if (prim.x == -1 && prim.y == -1)
          dist = distance_to_sphere(pos, sph[prim.z]);
        else if (prim.x == -1 && prim.y != -1)
          dist = compute_distance_to_cone(pos, sph[prim.y], sph[prim.z]);
        else if (prim.x != -1)
          dist = compute_distance_to_slab(pos, sph[prim.x], sph[prim.y], sph[prim.z]);
        if (dist < local_closest_dist) {
          local_closest_dist = fminf(local_closest_dist, dist);
          local_closest_id = i;
        }
      }
      min_distance[tid] = local_closest_dist;
      min_id[tid] = local_closest_id;
    }
    float reduced = min_distance[0];
    for (int i = 1; i < kWarpSize; i++)
      if (min_distance[i] < reduced) reduced = min_distance[i];
    result[s] = reduced;
    for (int i = 0; i < kWarpSize; ++i)
      if (fabsf(min_distance[i] - reduced) < 1e-10f) closest_id[s] = min_id[i];
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

// The reference's own GPU entry point, unmodified (needs a GPU).  Returns milliseconds of the
// whole call (H2D + kernel + D2H, like the reference times it) or <0 without a device.
double ref_d2m_run_gpu(const float* spheres, int n_sph, const float* samples, int n_samples,
                       const unsigned* offset, const unsigned* count, const int* prims, long n_prims,
                       float* result, int* closest_id) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return -1.0;
  GpuBuffer<float4> b_sph(n_sph);
  GpuBuffer<float3> b_smp(n_samples);
  GpuBuffer<uint> b_off(n_samples), b_cnt(n_samples);
  GpuBuffer<int3> b_pr(n_prims);
  GpuBuffer<float> b_res(n_samples);
  GpuBuffer<int> b_id(n_samples);
  memcpy(b_sph.HPtr(), spheres, sizeof(float4) * (size_t)n_sph);
  memcpy(b_smp.HPtr(), samples, sizeof(float3) * (size_t)n_samples);
  memcpy(b_off.HPtr(), offset, sizeof(uint) * (size_t)n_samples);
  memcpy(b_cnt.HPtr(), count, sizeof(uint) * (size_t)n_samples);
  memcpy(b_pr.HPtr(), prims, sizeof(int3) * (size_t)n_prims);
  for (int i = 0; i < n_samples; i++) {
    b_res.HPtr()[i] = 1e28f;  // fix_geo_error.cxx:300-366 initial fill
    b_id.HPtr()[i] = -1;
  }
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  compute_closest_dist2mat(b_sph, n_samples, b_smp, b_off, b_cnt, b_pr, b_res, b_id);
  cudaDeviceSynchronize();
  clock_gettime(CLOCK_MONOTONIC, &t1);
  memcpy(result, b_res.HPtr(), sizeof(float) * (size_t)n_samples);
  memcpy(closest_id, b_id.HPtr(), sizeof(int) * (size_t)n_samples);
  return 1e3 * ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec));
}

// The reference's kernel alone (ClosestDistanceToLocalMat, dist2mat.cu:195-278, in its own launch shape: one
// 32-thread block per sample, :301-304) on resident buffers: `warmup` untimed + `reps` timed launches between CUDA
// events.  Returns the mean milliseconds per launch (<0 without a device); results of the last launch are copied out.
double ref_d2m_kernel_ms(const float* spheres, int n_sph, const float* samples, int n_samples,
                         const unsigned* offset, const unsigned* count, const int* prims, long n_prims,
                         float* result, int* closest_id, int warmup, int reps) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return -1.0;
  GpuBuffer<float4> b_sph(n_sph);
  GpuBuffer<float3> b_smp(n_samples);
  GpuBuffer<uint> b_off(n_samples), b_cnt(n_samples);
  GpuBuffer<int3> b_pr(n_prims);
  GpuBuffer<float> b_res(n_samples);
  GpuBuffer<int> b_id(n_samples);
  memcpy(b_sph.HPtr(), spheres, sizeof(float4) * (size_t)n_sph);
  memcpy(b_smp.HPtr(), samples, sizeof(float3) * (size_t)n_samples);
  memcpy(b_off.HPtr(), offset, sizeof(uint) * (size_t)n_samples);
  memcpy(b_cnt.HPtr(), count, sizeof(uint) * (size_t)n_samples);
  memcpy(b_pr.HPtr(), prims, sizeof(int3) * (size_t)n_prims);
  for (int i = 0; i < n_samples; i++) {
    b_res.HPtr()[i] = 1e28f;
    b_id.HPtr()[i] = -1;
  }
  b_sph.H2D(); b_smp.H2D(); b_off.H2D(); b_cnt.H2D(); b_pr.H2D(); b_res.H2D(); b_id.H2D();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float total = 0.f;
  for (int it = 0; it < warmup + reps; it++) {
    cudaEventRecord(e0);
    ClosestDistanceToLocalMat<<<n_samples, kWarpSize>>>(b_smp.DPtr(), b_sph.DPtr(), b_pr.DPtr(), b_off.DPtr(),
                                                        b_cnt.DPtr(), b_res.DPtr(), b_id.DPtr());
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (it >= warmup) total += ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (cudaGetLastError() != cudaSuccess) return -2.0;
  b_res.D2H();
  b_id.D2H();
  if (result) memcpy(result, b_res.HPtr(), sizeof(float) * (size_t)n_samples);
  if (closest_id) memcpy(closest_id, b_id.HPtr(), sizeof(int) * (size_t)n_samples);
  return reps > 0 ? (double)total / reps : 0.0;
}

}  // extern "C"
