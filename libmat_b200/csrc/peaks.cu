// Measured arithmetic peaks of the device: dependent-chain-free FFMA and DFMA loops on every SM.  The HBM roofline
// says nothing about kernels that move 1-5 % of the memory bandwidth (K3, K5 are issue-bound); bench.py reports their
// algorithmic flops against these measured peaks as the secondary, compute roofline.
#include "mb_internal.h"

namespace {

template <typename T>
__global__ void __launch_bounds__(256) k_peak_fma(T* out, int iters) {
  // 8 independent accumulator chains per thread keep the FMA pipe full
  T a0 = (T)threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const T m = (T)0.999999, c = (T)1e-6;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
      a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <typename T>
double measure(mb_ctx* ctx, int iters) {
  cudaStream_t s = ctx->stream;
  const int blocks = ctx->sm_count * 8, threads = 256;
  DevBuf<T> out;
  out.reserve((size_t)blocks * threads);
  cudaEvent_t e0, e1;
  MB_CUDA(cudaEventCreate(&e0));
  MB_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; rep++) {
    MB_CUDA(cudaEventRecord(e0, s));
    ctx->n_launches++;
    k_peak_fma<T><<<blocks, threads, 0, s>>>(out.p, iters);
    MB_CUDA(cudaEventRecord(e1, s));
    MB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    MB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * threads;  // 64 FMAs per iteration and thread
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  MB_CUDA(cudaGetLastError());
  return best;
}

}  // namespace

void peaks_measure(mb_ctx* ctx, double* fp32_tflops, double* fp64_tflops) {
  // this TU is compiled with -fmad=true (build.py: FMAD_SOURCES): a * m + c is one FFMA / DFMA
  if (fp32_tflops) *fp32_tflops = measure<float>(ctx, 4096);
  if (fp64_tflops) *fp64_tflops = measure<double>(ctx, 512);
}
