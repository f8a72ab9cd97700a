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

// --------------------------------------------------------------------------------------------
// Correlation-pyramid storage ("warp-strip-major").  The target image of every level is
// cut into the 8x32 level-0 patches the GEMM tiles over (level l: (8>>l) x (32>>l)); queries are taken
// 32 at a time (one epilogue warp).  For one (pair, patch, query group) everything is one contiguous
// *group block*, split into sections of 32 lines (one line per query of the group):
//   fp16:  sec 0-3  level-0 rows (2s, 2s+1)   128 B lines      fp32:  sec 0-7  level-0 row s      128 B lines
//          sec 4    level-1 (4 x 16)           128 B lines             sec 8-9  level-1 rows 2k,2k+1 128 B lines
//          sec 5    level-2 (2 x 8) + level-3  64 B lines              sec 10   level-2 + level-3    128 B lines
// Inside a line the 16-byte chunks are XOR-swizzled with the lane index, so the epilogue's shared-memory
// staging (which is byte-identical to the global block) is bank-conflict free; sectors and 64-byte
// DRAM atoms are preserved by the XOR.  Blocks are ordered [pair][patch][group]: a GEMM tile (128
// queries x one patch) writes 4 consecutive blocks = 88 KB (fp16) of contiguous memory.
// --------------------------------------------------------------------------------------------
constexpr int PYR_PH = 8, PYR_PW = 32, PYR_QG = 32, PYR_MAX_LEVELS = 4;

struct PyrLayout {
  int H, W, L, es;
  int n_ph, n_pw, n_patches, n_groups;
  long long group_bytes, pair_bytes;
};

__host__ __device__ inline long long pyr_group_bytes(int es) { return es == 2 ? 22528 : 45056; }

inline bool pyr_layout(int H, int W, int L, int dtype, PyrLayout* p) {
  if (H < 1 || W < 1 || L < 1 || L > PYR_MAX_LEVELS) return false;
  if ((H >> (L - 1)) < 1 || (W >> (L - 1)) < 1) return false;
  p->H = H; p->W = W; p->L = L;
  p->es = dtype == EZB_F16 ? 2 : 4;
  p->n_ph = cdiv(H, PYR_PH);
  p->n_pw = cdiv(W, PYR_PW);
  p->n_patches = p->n_ph * p->n_pw;
  p->n_groups = cdiv(H * W, PYR_QG);
  p->group_bytes = pyr_group_bytes(p->es);
  p->pair_bytes = (long long)p->n_patches * p->n_groups * p->group_bytes;
  return true;
}

// byte offset, inside a group block, of element (yy, xx) of level l's tile for query `lane` of the group
__host__ __device__ inline int pyr_elem_offset(int es, int l, int lane, int yy, int xx) {
  int sec, inl, line = 128;
  if (es == 2) {
    if (l == 0) { sec = yy >> 1; inl = (yy & 1) * 64 + xx * 2; }
    else if (l == 1) { sec = 4; inl = yy * 32 + xx * 2; }
    else { sec = 5; line = 64; inl = (l == 2) ? yy * 16 + xx * 2 : 32 + xx * 2; }
  } else {
    if (l == 0) { sec = yy; inl = xx * 4; }
    else if (l == 1) { sec = 8 + (yy >> 1); inl = (yy & 1) * 64 + xx * 4; }
    else { sec = 10; inl = (l == 2) ? yy * 32 + xx * 4 : 64 + xx * 4; }
  }
  const int chunk = (inl >> 4) ^ (lane & (line / 16 - 1));
  return sec * 4096 + lane * line + chunk * 16 + (inl & 15);
}

}  // namespace ezb
