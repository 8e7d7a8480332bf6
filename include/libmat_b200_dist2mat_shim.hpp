// libmat_b200 -- C++ shim with the reference's exact host signature for dist2mat.
//
// Header-only; compile it inside LibMAT (it uses LibMAT's own GpuBuffer<T>, include/cuda_utils.h:28-154)
// and link libmat_b200.so.  It replaces the body of
//     compute_closest_dist2mat      reference src/dist2mat/dist2mat.cu:280-315
//                                   (declaration src/dist2mat/dist2mat.h:19-24)
// which load_and_compute_sample_dist2mat_gpubuffer calls (src/matfun_fix/fix_geo_error.cxx:365-366).
// Like the reference it reads the HOST side of the seven buffers and leaves the answers in
// results.HPtr() / closest_mat_id.HPtr() (the reference's trailing D2H, dist2mat.cu:306-307).
#pragma once

#include <cstdio>

#include "cuda_utils.h"  // LibMAT's GpuBuffer
#include "libmat_b200.h"
#include "libmat_b200_shim_ctx.hpp"

#ifndef LIBMAT_B200_NO_REFERENCE_NAMES
inline void compute_closest_dist2mat(GpuBuffer<float4>& spheres, const int num_samples, GpuBuffer<float3>& samples,
                                     GpuBuffer<uint>& offset, GpuBuffer<uint>& num_per_sample,
                                     GpuBuffer<int3>& prims, GpuBuffer<float>& results,
                                     GpuBuffer<int>& closest_mat_id) {
  mb_ctx* ctx = libmat_b200::thread_ctx();
  if (!ctx) return;
  results.HResize((size_t)num_samples);
  closest_mat_id.HResize((size_t)num_samples);
  const int rc = mb_dist2mat(ctx, reinterpret_cast<const float*>(spheres.HPtr()), (int)spheres.HSize(),
                             reinterpret_cast<const float*>(samples.HPtr()), num_samples, offset.HPtr(),
                             num_per_sample.HPtr(), reinterpret_cast<const int*>(prims.HPtr()),
                             (long)prims.HSize(), results.HPtr(), closest_mat_id.HPtr(), nullptr);
  if (rc) std::fprintf(stderr, "[libmat_b200] mb_dist2mat: %s\n", mb_last_error(ctx));
}
#endif
