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

// fp32 -> fp16 operand preparation shared by the tcgen05 kernels (defined in corr_pyramid.cu)
namespace prep {
// amax[which*B + b] = max |t_which[b, :]| as float bits (zeroes amax first); per_b elements per batch entry
int launch_absmax(const float* t0, const float* t1, long long per_b, unsigned int* amax, int B, cudaStream_t st);
// fp32 (B,C,N) -> fp16 (B,N,Cp) scaled by 2^-floor(log2(amax[b])) (into (-2,2)); channels >= C zero-filled
int launch_convert_kmajor(const float* src, __half* dst, const unsigned int* amax, int B, int C, int Cp, int N, cudaStream_t st);
// both tensors in one launch: src0 with amax[b], src1 with amax[B + b]
int launch_convert_kmajor2(const float* src0, __half* dst0, const float* src1, __half* dst1, const unsigned int* amax, int B, int C,
                           int Cp, int N, cudaStream_t st);
// floor 2x2 average pool of `planes` images (hs x ws, plane stride src_plane) into (hd x wd, plane stride dst_plane)
int launch_pool2x2(const float* src, float* dst, long long planes, int hs, int ws, long long src_plane, int hd, int wd,
                   long long dst_plane, cudaStream_t st);
}  // namespace prep

// Pyramid level geometry shared by K1/K2/K3 (see ezb_corr_pyramid_layout in the header).
struct LevelGeom {
  int h, w, pitch;
  int64_t qstride;
};

inline bool pyramid_geom(int H, int W, int L, int dtype, LevelGeom* g) {
  // gradient levels: rows packed (pitch = w), per-query stride rounded up to 16 bytes so that a level can be
  // described to TMA as a 2-D tensor (rows = queries, columns = y*w + x)
  const int es = dtype == EZB_F16 ? 2 : 4;
  const int align = 16 / es;
  for (int l = 0; l < L; ++l) {
    g[l].h = H >> l;
    g[l].w = W >> l;
    if (g[l].h < 1 || g[l].w < 1) return false;
    g[l].pitch = g[l].w;
    g[l].qstride = round_up(g[l].h * g[l].w, align);
  }
  return true;
}

// --------------------------------------------------------------------------------------------
// Correlation-pyramid storage ("subtile-sequence").
// For one query, level l of the pyramid is an (H>>l) x (W>>l) image.  Every level is cut into 8x8-pixel
// *subtiles* (128 B in fp16 = one DRAM fetch unit on B200, measured with tools/micro/dram_gran.cu: any
// access to a 128-byte line costs 128 bytes of DRAM traffic; a (2r+2)^2 = 10x10 window touches on average
// 4.5 such lines); the subtiles of all levels form one sequence
//     s = sub_base[l] + (y/8) * sw_l + (x/8),  sw_l = ceil(w_l/8),
// and 4 consecutive subtiles are one GEMM *tile*: 256 accumulator columns = one 256-element *record* per
// query,  column n = (s % 4) * 64 + (y % 8) * 8 + (x % 8).  The pooled levels are ordinary GEMM columns:
// avg-pooling is linear, so level l is the correlation with the l-times pooled fmap2.
// Queries are taken 32 at a time (one epilogue warp).  For one (pair, tile, query group) the 32 records
// form a contiguous *group block* of sections; section j holds, for each of the 32 queries, a 128-byte
// line with record elements [j*EPL, (j+1)*EPL), EPL = 64 (fp16) or 32 (fp32); a subtile is one 128-byte line in
// fp16 (two in fp32).  The GEMM epilogue writes lines straight from registers: four threads hold the four 32-byte
// quarters of a query's line, so one 256-bit store instruction of a warp writes 8 complete 128-byte lines.
// Blocks are ordered [pair][tile][group]: one GEMM tile (128 queries x one record) writes 4 consecutive blocks =
// 64 KB (fp16) of contiguous memory.  Pixels outside a level's valid area (and
// levels >= L) are exact zeros.
// --------------------------------------------------------------------------------------------
constexpr int PYR_SY = 8, PYR_SX = 8, PYR_SUB = PYR_SY * PYR_SX, PYR_QG = 32, PYR_MAX_LEVELS = 4, PYR_REC = 256;
constexpr int PYR_SUBS_PER_TILE = PYR_REC / PYR_SUB;

struct PyrLayout {
  int H, W, L, es;
  int sw[PYR_MAX_LEVELS], sub_base[PYR_MAX_LEVELS + 1];  // subtiles per row of level l; first subtile of level l
  int n_tiles, n_groups;
  long long group_bytes, pair_bytes;
};

__host__ __device__ inline long long pyr_group_bytes(int es) { return (long long)PYR_QG * PYR_REC * es; }

inline bool pyr_layout(int H, int W, int L, int dtype, PyrLayout* p) {
  if (H < 1 || W < 1 || L < 1 || L > PYR_MAX_LEVELS) return false;
  if ((H >> (L - 1)) < 1 || (W >> (L - 1)) < 1) return false;
  p->H = H; p->W = W; p->L = L;
  p->es = dtype == EZB_F16 ? 2 : 4;
  int base = 0;
  for (int l = 0; l < PYR_MAX_LEVELS; ++l) {
    p->sub_base[l] = base;
    p->sw[l] = l < L ? cdiv(W >> l, PYR_SX) : 0;
    if (l < L) base += cdiv(H >> l, PYR_SY) * p->sw[l];
  }
  p->sub_base[PYR_MAX_LEVELS] = base;
  p->n_tiles = cdiv(base, PYR_SUBS_PER_TILE);
  p->n_groups = cdiv(H * W, PYR_QG);
  p->group_bytes = pyr_group_bytes(p->es);
  p->pair_bytes = (long long)p->n_tiles * p->n_groups * p->group_bytes;
  return true;
}

// byte offset, inside a group block, of record element n (0..255) of query `lane` of the group
__host__ __device__ inline int pyr_record_offset(int es, int lane, int n) {
  const int epl = 128 / es;  // record elements per 128-byte line
  const int sec = n / epl;
  return sec * 4096 + lane * 128 + (n - sec * epl) * es;
}
// The GEMM epilogue reads the accumulator with tcgen05.ld.16x256b, which hands thread t the columns 8i + 2(t%4) + {0,1};
// for a thread's values to be 32 contiguous bytes of its query's line, accumulator column c of a line holds line element
//   m(c) = ((c%8)/2 * (epl/8) + c/8) * 2 + c%2,
// i.e. row c of the packed B operand is the f2 pixel of record element m(c) (pack_record_operand_kernel).
__host__ __device__ inline int pyr_line_element_of_column(int epl, int c) { return (((c & 7) >> 1) * (epl >> 3) + (c >> 3)) * 2 + (c & 1); }

}  // namespace ezb
