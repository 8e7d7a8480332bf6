// libmat_b200 -- the per-host-thread context the C++ shims share (the reference is single-threaded
// on device 0: src/rpd3d/voronoi.cu:465).
#pragma once

#include <cstdio>
#include <cstdlib>

#include "libmat_b200.h"

namespace libmat_b200 {

// one context per host thread, created on first use (the reference is single-threaded, device 0)
inline mb_ctx* thread_ctx() {
  struct Holder {
    mb_ctx* ctx = nullptr;
    ~Holder() {
      if (ctx) mb_destroy(ctx);
    }
  };
  static thread_local Holder h;
  if (!h.ctx) {
    int err = 0;
    const char* dev = std::getenv("MB_DEVICE");
    h.ctx = mb_create(dev ? std::atoi(dev) : -1, &err);
    if (!h.ctx) std::fprintf(stderr, "[libmat_b200] mb_create failed (%d): no CUDA device, no CPU fallback\n", err);
  }
  return h.ctx;
}

}  // namespace libmat_b200
