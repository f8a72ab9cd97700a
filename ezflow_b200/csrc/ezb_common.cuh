// Shared host-side helpers for the ezflow_b200 C-ABI library (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/ezflow_b200.h"

namespace ezb {

// thread-local error string, read back through ezb_last_error()
inline char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

#define EZB_CUDA_OK(expr)                                                                      \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return ::ezb::fail(EZB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                         __FILE__, __LINE__);                                                  \
  } while (0)

#define EZB_REQUIRE(cond, ...)                                \
  do {                                                        \
    if (!(cond)) return ::ezb::fail(EZB_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define EZB_LAUNCH_OK()                                                                             \
  do {                                                                                              \
    cudaError_t _e = cudaGetLastError();                                                            \
    if (_e != cudaSuccess)                                                                          \
      return ::ezb::fail(EZB_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                         __FILE__, __LINE__);                                                       \
  } while (0)

inline int cdiv(int a, int b) { return (a + b - 1) / b; }
inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return cdiv(a, b) * b; }

inline int num_sms() {
  static thread_local int dev_cached = -1;
  static thread_local int sms = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev != dev_cached) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    dev_cached = dev;
  }
  return sms > 0 ? sms : 148;
}

// Pyramid level geometry shared by K1/K2/K3 (see ezb_corr_pyramid_layout in the header).
struct LevelGeom {
  int h, w, pitch;
  int64_t qstride;
};

inline bool pyramid_geom(int H, int W, int L, int dtype, LevelGeom* g) {
  const int es = dtype == EZB_F16 ? 2 : 4;
  const int align = 16 / es;
  for (int l = 0; l < L; ++l) {
    g[l].h = H >> l;
    g[l].w = W >> l;
    if (g[l].h < 1 || g[l].w < 1) return false;
    g[l].pitch = round_up(g[l].w, align);
    g[l].qstride = (int64_t)g[l].h * g[l].pitch;
  }
  return true;
}

}  // namespace ezb
