// K1 + K2: RAFT all-pairs correlation pyramid, all levels produced by ONE tcgen05 GEMM.
//
// Replaces MutliScalePairwise4DCorr.__init__/.corr (reference ezflow/similarity/correlation/
// pairwise.py:26-41, :78-88):  corr[b,q,t] = sum_c f1[b,c,q] f2[b,c,t] / sqrt(C), then L-1 floor
// 2x2 average pools over the target axes (t = h2*W + w2).
//
// Average pooling is linear, so level l of the pyramid is the correlation of f1 with the l-times
// average-pooled f2:  pool_l(corr)[q, t_l] = < f1[:,q], pool_l(f2)[:, t_l] > / sqrt(C).  The pooled
// feature maps are tiny (1/4, 1/16, 1/64 of f2), so the pooled levels are simply more GEMM columns:
// the tensor core does the pooling, the epilogue is a pure convert-and-store, and the pyramid is
// written exactly once (never re-read to pool it).
//
// Design:
//   prep   : per-pair absmax -> power-of-two pre-scale;
//            f1: fp32 (B,C,N) -> fp16 K-major (B,N,Cp)                                   [A operand]
//            f2: fp32 2x2 pools -> levels 1..L-1; then for every record tile a 256 x Cp fp16
//                operand tile whose row n is the (pooled) f2 pixel of record column n     [B operand]
//                (ezb_common.cuh: 4 subtiles of 8x8 pixels, all levels in one sequence;
//                rows permuted inside each 128-byte line for the epilogue, pyr_line_element_of_column)
//   gemm   : persistent, warp-specialised tcgen05 kernel.  One CTA tile = 128 queries (TMEM lanes) x
//            one record (256 TMEM columns).  The record operand (128 KB for C=256) stays resident in
//            shared memory while the CTA sweeps query blocks (A streamed through a 6-stage TMA ring);
//            two 256-column accumulators double-buffer MMA against the epilogue.
//            CTAs run as clusters of 2 that sweep the SAME query blocks against two different records:
//            each CTA fetches half of every A stage and TMA-multicasts it to both, which halves the L2
//            read traffic of the re-streamed A operand (the kernel is energy-bound at the board power cap; every
//            byte that does not move helps: profiles/r01_gemm_experiments.md).
//   epilogue: TMEM -> registers (tcgen05.ld.16x256b) -> (scale) -> fp16/fp32 -> 256-bit streaming stores straight
//            from registers: the operand rows are ordered so that a quad of threads owns one query's 128-byte
//            line, i.e. every store instruction writes 8 complete lines; no shared-memory staging.
#include "ezb_common.cuh"
#include "ezb_ptx.cuh"
#include "ezb_tma.cuh"

#include <cstdlib>

namespace ezb {
namespace k1 {

constexpr int BM = 128;           // queries per tile  (UMMA M, TMEM lanes)
constexpr int BN = PYR_REC;       // record elements   (UMMA N = 256, TMEM columns)
constexpr int BK = 64;            // fp16 elements per 128-byte swizzle row
constexpr int KB_MAX = 4;         // resident B k-blocks  (C <= 256)
#ifndef EZB_GEMM_STAGES
#define EZB_GEMM_STAGES 6
#endif
constexpr int STAGES = EZB_GEMM_STAGES;  // A ring
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB
#ifndef EZB_GEMM_CLUSTER
#define EZB_GEMM_CLUSTER 2
#endif
constexpr int CLUSTER = EZB_GEMM_CLUSTER;    // CTAs sharing one A stream
constexpr int A_SLICE_ROWS = BM / CLUSTER;   // rows of an A stage each CTA of the cluster fetches (and multicasts)
constexpr int B_KB_BYTES = BN * BK * 2;      // 32 KB
#ifndef EZB_GEMM_EPI_SPLIT
#define EZB_GEMM_EPI_SPLIT 1
#endif
constexpr int EPI_SPLIT = EZB_GEMM_EPI_SPLIT;  // epilogue warps per TMEM lane quadrant (each takes 1/EPI_SPLIT of the record)
constexpr int EPI_WARPS = 4 * EPI_SPLIT;
constexpr int NUM_THREADS = 128 + EPI_WARPS * 32;
constexpr int SMEM_B = 0;
constexpr int SMEM_A = SMEM_B + KB_MAX * B_KB_BYTES;
constexpr int SMEM_BAR = SMEM_A + STAGES * A_STAGE_BYTES;
constexpr int SMEM_BYTES = SMEM_BAR + 256 + 1024;  // + barrier block + alignment slack
constexpr int TMEM_COLS = 512;                     // two 256-column accumulators

struct GemmParams {
  int B, N, KB, L;
  int n_recs, n_qblocks, n_groups;
  long long group_bytes;
  const float* epi_scale;  // [B]
  char* pyr;               // pyramid buffer, group blocks ordered [pair][record tile][group]
  int dbg;                 // profiling experiments only (EZB_GEMM_DEBUG env): 1 no global stores, 4 no MMA issue,
                           // 8 no A loads, 16 no TMEM reads, 64 no MMA for the last 15 % of the records
};

// ------------------------------------------------------------------------------ prep kernels
__global__ void absmax_kernel(const float* __restrict__ f1, const float* __restrict__ f2, long long per_b,
                              unsigned int* __restrict__ amax /*[2][B]*/, int B) {
  const int b = blockIdx.y, which = blockIdx.z;
  const float* x = (which ? f2 : f1) + (long long)b * per_b;
  float m = 0.f;
  const long long n4 = per_b >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (vec_ok) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n4; i += 8 * stride) {  // eight independent 16-byte loads in flight per thread
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(x4 + i + u * stride);
#pragma unroll
      for (int u = 0; u < 8; ++u) m = fmaxf(m, fmaxf(fmaxf(fabsf(v[u].x), fabsf(v[u].y)), fmaxf(fabsf(v[u].z), fabsf(v[u].w))));
    }
    for (; i < n4; i += stride) {
      float4 v = __ldg(x4 + i);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    for (long long k = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; k < per_b; k += stride) m = fmaxf(m, fabsf(x[k]));
  } else {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_b; i += (long long)gridDim.x * blockDim.x)
      m = fmaxf(m, fabsf(x[i]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax + which * B + b, __float_as_uint(m));
}

__device__ __forceinline__ int pow2_exponent(unsigned int amax_bits) {
  // floor(log2(amax)) for normal floats; 0 for zero/denormal input
  const int e = (int)((amax_bits >> 23) & 0xff);
  return e == 0 ? 0 : e - 127;
}
__device__ __forceinline__ float pow2f(int e) {  // 2^e, |e| <= 126
  e = max(-126, min(126, e));
  return __uint_as_float((unsigned int)(e + 127) << 23);
}
// fp16 storage: operands are pre-scaled into (-2,2), so |acc| < 4C.  Stored = acc * F with F the largest
// power of two keeping |stored| < 2^15: fp16 keeps full relative precision over ~19 binades below that.
// F is folded into the B operand (|B| < 2F: 64 for C = 256, at most 2^14 for C = 1 — still fp16-safe).
__device__ __forceinline__ float fp16_store_factor(int C) {
  int lg = 0;
  while ((1 << lg) < 4 * C) ++lg;
  return pow2f(15 - lg);
}

// bscale[b] multiplies f2 when the B operand is packed, epi[b] multiplies the fp32 accumulator in the
// epilogue, deq[b] multiplies stored values at lookup time:  2^-e1 * bscale * epi * deq == 1/sqrt(C).
__global__ void scales_kernel(const unsigned int* __restrict__ amax, int B, int C, int dtype, float* __restrict__ bscale,
                              float* __restrict__ epi, float* __restrict__ deq) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int e1 = pow2_exponent(amax[b]), e2 = pow2_exponent(amax[B + b]);
  const float true_scale = pow2f(e1 + e2) * rsqrtf((float)C);  // exact for power-of-four C, <= 1 ulp otherwise
  if (dtype == EZB_F16) {
    const float F = fp16_store_factor(C);
    bscale[b] = pow2f(-e2) * F;
    epi[b] = 1.0f;  // nothing left to do in the epilogue
    deq[b] = true_scale / F;
  } else {
    bscale[b] = pow2f(-e2);
    epi[b] = true_scale;
    deq[b] = 1.0f;
  }
}

// A operand: fp32 (B,C,N) -> fp16 (B,N,Cp) scaled by 2^-e1 (per pair); channels >= C zero-filled.
// With a second tensor (src1/dst1), blockIdx.z in [B, 2B) converts it with the scales amax[B + b].
__global__ void __launch_bounds__(256) convert_kmajor_kernel(const float* __restrict__ f1, __half* __restrict__ o1,
                                                             const float* __restrict__ src1, __half* __restrict__ dst1,
                                                             const unsigned int* __restrict__ amax, int B, int C, int Cp, int N) {
  __shared__ float tile[64][33];
  const int which = blockIdx.z >= B, b = blockIdx.z - which * B;
  const float* src = (which ? src1 : f1) + (long long)b * C * N;
  __half* dst = (which ? dst1 : o1) + (long long)b * N * Cp;
  const float s = pow2f(-pow2_exponent(amax[blockIdx.z]));
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + warp * 8 + i, n = n0 + lane;
    tile[warp * 8 + i][lane] = (c < C && n < N) ? __ldg(src + (long long)c * N + n) * s : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int nl = warp * 4 + i, n = n0 + nl;
    if (n < N) {
      const __half2 v = __floats2half2_rn(tile[2 * lane][nl], tile[2 * lane + 1][nl]);
      *reinterpret_cast<__half2*>(dst + (long long)n * Cp + c0 + 2 * lane) = v;
    }
  }
}

// fmap2 levels 1..L-1: successive floor 2x2 average pools (F.avg_pool2d(.., 2, stride=2), pairwise.py:40)
// applied to the feature map instead of the 4-D volume.  planes = B*C.
__global__ void __launch_bounds__(256) pool2x2_kernel(const float* __restrict__ src, float* __restrict__ dst, long long planes,
                                                      int hs, int ws, long long src_plane, int hd, int wd, long long dst_plane) {
  const long long total = planes * hd * wd;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % wd);
    const long long r = i / wd;
    const int y = (int)(r % hd);
    const long long pl = r / hd;
    const float* s = src + pl * src_plane + (long long)(2 * y) * ws + 2 * x;
    dst[pl * dst_plane + (long long)y * wd + x] = 0.25f * ((__ldg(s) + __ldg(s + 1)) + (__ldg(s + ws) + __ldg(s + ws + 1)));
  }
}

// B operand: for every (pair, record tile) a [256 rows][Cp] fp16 tile; row n holds the level-l pixel of
// (pooled) f2 that record column n stands for (zero outside the level).
// Block = (record tile, 32-channel chunk, pair); every thread gathers its pixel over the chunk's channels, then
// rows are written K-major.
constexpr int PACK_CH = 32;
struct PackParams {
  const float* lvl[PYR_MAX_LEVELS];  // level l of f2: (B, C, h[l], w[l]) fp32
  int h[PYR_MAX_LEVELS], w[PYR_MAX_LEVELS], sw[PYR_MAX_LEVELS], sub_base[PYR_MAX_LEVELS + 1];
  int L, C, Cp, n_recs, epl;  // epl: record elements per 128-byte line of the pyramid (64 fp16 / 32 fp32)
  const float* bscale;  // [B]
  __half* bp;           // (B, n_recs, 256, Cp)
};
__global__ void __launch_bounds__(256) pack_record_operand_kernel(const PackParams P) {
  __shared__ float tile[PACK_CH][PYR_REC + 1];
  const int rec = blockIdx.x, c0 = blockIdx.y * PACK_CH, b = blockIdx.z;
  // operand row (= accumulator column) threadIdx.x holds record element n (ezb_common.cuh: the epilogue's
  // tcgen05.ld.16x256b hands every thread 32 contiguous bytes of a line)
  const int col = threadIdx.x;
  const int n = col / P.epl * P.epl + pyr_line_element_of_column(P.epl, col % P.epl);
  const int s = rec * PYR_SUBS_PER_TILE + n / PYR_SUB;
  int l = 0;
  while (l + 1 < P.L && s >= P.sub_base[l + 1]) ++l;
  const float* src = nullptr;
  long long cstride = 0;
  if (s < P.sub_base[PYR_MAX_LEVELS]) {
    const int ls = s - P.sub_base[l];
    const int sy = ls / P.sw[l], sx = ls - sy * P.sw[l];
    const int y = sy * PYR_SY + (n % PYR_SUB) / PYR_SX, x = sx * PYR_SX + n % PYR_SX;
    if (y < P.h[l] && x < P.w[l]) {
      cstride = (long long)P.h[l] * P.w[l];
      src = P.lvl[l] + ((long long)b * P.C + c0) * cstride + (long long)y * P.w[l] + x;
    }
  }
  const float sc = __ldg(P.bscale + b);
  const int cmax = min(PACK_CH, P.C - c0);  // channels >= C are zero-filled
#pragma unroll 8
  for (int c = 0; c < PACK_CH; ++c) tile[c][col] = (src != nullptr && c < cmax) ? __ldg(src + c * cstride) * sc : 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __half* dst = P.bp + ((long long)b * P.n_recs + rec) * PYR_REC * P.Cp + c0 + lane;
  for (int row = warp; row < PYR_REC; row += 8) dst[(long long)row * P.Cp] = __float2half_rn(tile[lane][row]);
}

// Pool + pack in ONE pass over f2 (round 2; replaces L-1 pool2x2 launches and pack_record_operand for the forward build):
// block = (16 x 16 level-0 pixels of one pair, 64 channels), 512 threads.  The region is loaded once (16-byte loads, 8 in flight per
// thread) and transposed into shared memory [pixel][channel]; level 1 is pooled in registers on the way, levels 2..3 from
// shared memory, with the exact expression of pool2x2_kernel (floor pooling: a valid level-l pixel always has its four sources, invalid ones are
// zero), so the region yields 256 + 64 + 16 + 4 operand rows; each row leaves as 128 contiguous bytes (64 channels fp16)
// of the (B, n_recs, 256, Cp) operand.  f2 is read once (pool + pack read it 1.3 + 1.3 times and wrote the pooled levels),
// and the row pieces are full 128-byte lines instead of 64-byte ones.  The region grid is extended so that every pixel
// of every existing subtile of every level is written (zeros outside a level); the blocks of region 0 also zero the
// padding subtiles at the end of the sequence.
constexpr int PP_R = 16, PP_CH = 64, PP_PITCH = PP_CH + 4, PP_ROWS = 256 + 64 + 16 + 4, PP_NT = 512;
constexpr size_t PP_SMEM = (size_t)PP_ROWS * PP_PITCH * sizeof(float);
struct PoolPackParams {
  const float* f2;  // (B, C, H, W)
  int H, W, C, Cp, L, n_recs, epl, rx, ry;
  int h[PYR_MAX_LEVELS], w[PYR_MAX_LEVELS], sw[PYR_MAX_LEVELS], nsy[PYR_MAX_LEVELS], sub_base[PYR_MAX_LEVELS + 1];
  const float* bscale;
  __half* bp;
};
template <bool VEC>
__global__ void __launch_bounds__(PP_NT, 2) pool_pack_kernel(const PoolPackParams P) {
  extern __shared__ __align__(16) float pp[];  // [PP_ROWS][PP_PITCH]
  __shared__ int rowoff[PP_ROWS];
  const int tid = threadIdx.x;
  const int ry = blockIdx.x / P.rx, rx = blockIdx.x - ry * P.rx, c0 = blockIdx.y * PP_CH, b = blockIdx.z;
  const int y0 = ry * PP_R, x0 = rx * PP_R;
  const long long HW = (long long)P.H * P.W;
  const float* src = P.f2 + ((long long)b * P.C + c0) * HW;

  // ---- level 0 + level 1: thread = (4 consecutive channels, 2 rows, pixel quad): 8 loads in flight; the 4 channels of a pixel
  //      leave as ONE 16-byte shared-memory store (a quarter-warp = 8 channel groups of one pixel: conflict-free), and the
  //      thread pools its 2 x 4 pixels into two level-1 pixels in registers (same expression as pool2x2_kernel)
  float4 v[2][4];
  {
    const int lane = tid & 31, warp = tid >> 5;
    const int cg = (lane & 7) + 8 * (warp & 1), xq = lane >> 3, rp = warp >> 1;
    const int x = x0 + 4 * xq;
    const float* g0 = src + (long long)(4 * cg) * HW + (long long)(y0 + 2 * rp) * P.W + x;
    const int nch = min(4, P.C - c0 - 4 * cg);  // channels of this group inside C (the rest are zero)
#pragma unroll
    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[rr][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < nch && y0 + 2 * rp + rr < P.H) {
          const float* g = g0 + (long long)k * HW + rr * P.W;
          if constexpr (VEC) {
            if (x < P.W) v[rr][k] = __ldg(reinterpret_cast<const float4*>(g));
          } else {
            if (x < P.W) v[rr][k].x = __ldg(g);
            if (x + 1 < P.W) v[rr][k].y = __ldg(g + 1);
            if (x + 2 < P.W) v[rr][k].z = __ldg(g + 2);
            if (x + 3 < P.W) v[rr][k].w = __ldg(g + 3);
          }
        }
      }
  }
  // ---- while the loads are in flight: the operand row every one of the region's pixels goes to (element offset inside the
  //      pair's operand, -1: no such subtile on the extended region grid), computed once per row instead of once per lane
  if (tid < PP_ROWS) {
    const int r = tid;
    int l = 0, base = 0, nsh = 4;  // the region holds (1 << nsh)^2 pixels of level l
    if (r >= 336) { l = 3; base = 336; nsh = 1; } else if (r >= 320) { l = 2; base = 320; nsh = 2; } else if (r >= 256) { l = 1; base = 256; nsh = 3; }
    const int p = r - base, py = p >> nsh, px = p & ((1 << nsh) - 1);
    const int y = (y0 >> l) + py, x = (x0 >> l) + px;
    const int sy = y / PYR_SY, sx = x / PYR_SX;
    int off = -1;
    if (l < P.L && sy < P.nsy[l] && sx < P.sw[l]) {
      const int epl_mask = P.epl - 1, e8_shift = P.epl == 64 ? 3 : 2;  // epl = 64 (fp16 pyramid) or 32 (fp32): shifts, no divisions
      const int sidx = P.sub_base[l] + sy * P.sw[l] + sx;
      const int n = (sidx % PYR_SUBS_PER_TILE) * PYR_SUB + (y % PYR_SY) * PYR_SX + x % PYR_SX, m = n & epl_mask, t = m >> 1;
      const int col = n - m + (t & ((1 << e8_shift) - 1)) * 8 + (t >> e8_shift) * 2 + (m & 1);  // inverse of pyr_line_element_of_column
      off = ((sidx / PYR_SUBS_PER_TILE) * PYR_REC + col) * P.Cp;
    }
    rowoff[r] = off;
  }
  {
    const int lane = tid & 31, warp = tid >> 5;
    const int cg = (lane & 7) + 8 * (warp & 1), xq = lane >> 3, rp = warp >> 1;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      float* d = pp + ((2 * rp + rr) * PP_R + 4 * xq) * PP_PITCH + 4 * cg;
      *reinterpret_cast<float4*>(d) = make_float4(v[rr][0].x, v[rr][1].x, v[rr][2].x, v[rr][3].x);
      *reinterpret_cast<float4*>(d + PP_PITCH) = make_float4(v[rr][0].y, v[rr][1].y, v[rr][2].y, v[rr][3].y);
      *reinterpret_cast<float4*>(d + 2 * PP_PITCH) = make_float4(v[rr][0].z, v[rr][1].z, v[rr][2].z, v[rr][3].z);
      *reinterpret_cast<float4*>(d + 3 * PP_PITCH) = make_float4(v[rr][0].w, v[rr][1].w, v[rr][2].w, v[rr][3].w);
    }
    if (P.L > 1) {
      const bool oky = (y0 >> 1) + rp < P.h[1];
      const bool ok0 = oky && (x0 >> 1) + 2 * xq < P.w[1], ok1 = oky && (x0 >> 1) + 2 * xq + 1 < P.w[1];
      float a[4], b2[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        a[k] = ok0 ? 0.25f * ((v[0][k].x + v[0][k].y) + (v[1][k].x + v[1][k].y)) : 0.f;
        b2[k] = ok1 ? 0.25f * ((v[0][k].z + v[0][k].w) + (v[1][k].z + v[1][k].w)) : 0.f;
      }
      float* d = pp + (256 + rp * 8 + 2 * xq) * PP_PITCH + 4 * cg;
      *reinterpret_cast<float4*>(d) = make_float4(a[0], a[1], a[2], a[3]);
      *reinterpret_cast<float4*>(d + PP_PITCH) = make_float4(b2[0], b2[1], b2[2], b2[3]);
    }
  }
  __syncthreads();
  // ---- levels 2, 3 from the level below (rows 320.., 336..)
  {
    const int c = tid & 63, q = tid >> 6;  // 8 pixels per pass
    int src_row = 256, dst_row = 320, n = 4;  // n x n pixels at the level being produced
#pragma unroll
    for (int l = 2; l < PYR_MAX_LEVELS; ++l) {
      if (l < P.L) {
        for (int p = q; p < n * n; p += PP_NT / 64) {
          const int py = p / n, px = p - py * n;
          const float* s0 = pp + (src_row + (2 * py) * (2 * n) + 2 * px) * PP_PITCH + c;
          const float* s1 = s0 + 2 * n * PP_PITCH;
          const bool ok = (y0 >> l) + py < P.h[l] && (x0 >> l) + px < P.w[l];
          pp[(dst_row + p) * PP_PITCH + c] = ok ? 0.25f * ((s0[0] + s0[PP_PITCH]) + (s1[0] + s1[PP_PITCH])) : 0.f;
        }
      }
      __syncthreads();
      src_row = dst_row;
      dst_row += n * n;
      n >>= 1;
    }
  }
  // ---- rows out: a quarter-warp per row, 16 bytes (8 channels) per lane
  const float sc = __ldg(P.bscale + b);
  const int piece = tid & 7;
  __half* bp = P.bp + (long long)b * P.n_recs * PYR_REC * P.Cp + c0 + 8 * piece;
  for (int r = tid >> 3; r < PP_ROWS; r += PP_NT / 8) {
    const int off = rowoff[r];
    if (off < 0) continue;
    const float* rs = pp + r * PP_PITCH + 8 * piece;
    const float4 a = *reinterpret_cast<const float4*>(rs), c = *reinterpret_cast<const float4*>(rs + 4);
    const __half2 h0 = __floats2half2_rn(a.x * sc, a.y * sc), h1 = __floats2half2_rn(a.z * sc, a.w * sc);
    const __half2 h2 = __floats2half2_rn(c.x * sc, c.y * sc), h3 = __floats2half2_rn(c.z * sc, c.w * sc);
    uint4 u;
    u.x = *reinterpret_cast<const uint32_t*>(&h0); u.y = *reinterpret_cast<const uint32_t*>(&h1);
    u.z = *reinterpret_cast<const uint32_t*>(&h2); u.w = *reinterpret_cast<const uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(bp + off) = u;
  }
  if (blockIdx.x == 0) {  // padding subtiles behind the last level: zero rows
    const int s_end = P.n_recs * PYR_SUBS_PER_TILE;
    for (int i = (tid >> 3); i < (s_end - P.sub_base[PYR_MAX_LEVELS]) * PYR_SUB; i += PP_NT / 8) {  // any column order: all zero
      const int sidx = P.sub_base[PYR_MAX_LEVELS] + i / PYR_SUB;
      const long long row = (long long)(sidx / PYR_SUBS_PER_TILE) * PYR_REC + (sidx % PYR_SUBS_PER_TILE) * PYR_SUB + i % PYR_SUB;
      *reinterpret_cast<uint4*>(bp + row * P.Cp) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
}

// ----------------------------------------------------------------------------- epilogue
__device__ __forceinline__ uint32_t pack_half2(uint32_t a_bits, uint32_t b_bits) {
  __half2 h = __floats2half2_rn(__uint_as_float(a_bits), __uint_as_float(b_bits));
  return *reinterpret_cast<uint32_t*>(&h);
}

// Epilogue of one tile for one warp: 32 queries x one 256-element record -> one group block, no shared memory.
// tcgen05.ld.16x256b gives thread t, for the two queries (t/4) and (t/4)+8 of a 16-lane half, the accumulator
// columns 8i + 2(t%4) + {0,1}; the B operand is packed so that these are 32 CONTIGUOUS bytes of the query's
// 128-byte line (ezb_common.cuh), so the four threads of a quad write one complete line and a warp-wide 256-bit
// store writes 8 complete lines.  Loads are software-pipelined one step ahead of the convert + store.
template <typename OutT>
__device__ __forceinline__ void epilogue_tile(uint32_t t_base /* lane = first lane of the quadrant */, int part, int lane,
                                              float scale, uint32_t tmem_empty_bar, char* gblock, int dbg) {
  constexpr bool kHalf = sizeof(OutT) == 2;
  constexpr int kSections = kHalf ? 4 : 8;  // 128-byte lines per record: 64 / 32 record elements each
  constexpr int kCols = kHalf ? 64 : 32;    // accumulator columns per section
  constexpr int kRegs = kCols / 2;          // registers per load: 2 queries x kCols/4 columns per thread
  constexpr int kSteps = kSections * 2 / EPI_SPLIT;  // (section, 16-lane half) steps of this warp
  const int step0 = part * kSteps;
  if ((dbg & 16) || gblock == nullptr) {    // nothing to write (profiling switch / query group beyond N)
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(tmem_empty_bar);
    return;
  }
  const int qrow = lane >> 2, j = lane & 3;
  auto load = [&](int lstep, uint32_t (&r)[kRegs]) {
    const int step = step0 + lstep;
    const uint32_t taddr = t_base + ((uint32_t)((step & 1) * 16) << 16) + (step >> 1) * kCols;
    if constexpr (kHalf) ptx::tmem_ld_16x256b_x8(taddr, r); else ptx::tmem_ld_16x256b_x4(taddr, r);
  };
  uint32_t r[2][kRegs];
  load(0, r[0]);
#pragma unroll
  for (int step = 0; step < kSteps; ++step) {
    ptx::tmem_ld_wait();  // `step` has landed
    if (step + 1 < kSteps) {
      load(step + 1, r[(step + 1) & 1]);
    } else {  // accumulator fully drained into registers: hand the TMEM buffer back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tmem_empty_bar);
    }
    const uint32_t(&v)[kRegs] = r[step & 1];
    const int gs = step0 + step;
    char* line0 = gblock + (gs >> 1) * 4096 + ((gs & 1) * 16 + qrow) * 128 + j * 32;  // query (t/4); +8 queries = +1024 B
    if (dbg & 1) continue;
    if constexpr (kHalf) {
      // fp16 storage: the store factor is folded into the B operand, the accumulator is the stored value
      uint32_t w0[8], w1[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        w0[i] = pack_half2(v[4 * i], v[4 * i + 1]);
        w1[i] = pack_half2(v[4 * i + 2], v[4 * i + 3]);
      }
      ptx::st_global_v8_hint(line0, w0[0], w0[1], w0[2], w0[3], w0[4], w0[5], w0[6], w0[7], ptx::kEvictFirst);
      ptx::st_global_v8_hint(line0 + 1024, w1[0], w1[1], w1[2], w1[3], w1[4], w1[5], w1[6], w1[7], ptx::kEvictFirst);
    } else {
      auto sc = [&](uint32_t bits) { return __float_as_uint(__uint_as_float(bits) * scale); };
      ptx::st_global_v8_hint(line0, sc(v[0]), sc(v[1]), sc(v[4]), sc(v[5]), sc(v[8]), sc(v[9]), sc(v[12]), sc(v[13]), ptx::kEvictFirst);
      ptx::st_global_v8_hint(line0 + 1024, sc(v[2]), sc(v[3]), sc(v[6]), sc(v[7]), sc(v[10]), sc(v[11]), sc(v[14]), sc(v[15]),
                             ptx::kEvictFirst);
    }
  }
}

// ------------------------------------------------------------------------------- the GEMM kernel
template <typename OutT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
    corr_pyramid_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                             const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  const uint32_t sB = base + SMEM_B, sA = base + SMEM_A, sBar = base + SMEM_BAR;
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (STAGES + s); };
  const uint32_t b_full = sBar + 8u * (2 * STAGES), b_empty = b_full + 8u;
  auto tmem_full = [&](int a) { return sBar + 8u * (2 * STAGES + 2 + a); };
  auto tmem_empty = [&](int a) { return sBar + 8u * (2 * STAGES + 4 + a); };
  volatile uint32_t* tmem_ptr_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + SMEM_BAR + 8 * (2 * STAGES + 6));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), CLUSTER);  // the MMA warps of every CTA of the cluster release a stage
    }
    ptx::mbar_init(b_full, 1);
    ptx::mbar_init(b_empty, 1);
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tmem_full(a), 1);
      ptx::mbar_init(tmem_empty(a), EPI_WARPS);  // one arrive per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_ptr_slot)), TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync();  // also a CTA barrier; the peer's mbarriers are initialised before anything lands on them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_slot;

  // Pairs are processed one after the other by the whole grid (so the L2 working set is ONE pair's
  // operands); within a pair every cluster owns a contiguous, balanced range of cluster tiles.
  // cluster tile t of pair b -> (record group rg = t / n_qblocks, qb = t % n_qblocks); CTA rank r of the
  // cluster computes record rg * CLUSTER + r (idle columns if that is past the last record).
  const uint32_t crank = ptx::cluster_ctarank();
  const int n_clusters = gridDim.x / CLUSTER, cid = blockIdx.x / CLUSTER;
  const long long tiles_per_pair = (long long)((p.n_recs + CLUSTER - 1) / CLUSTER) * p.n_qblocks;
  const long long t_begin = tiles_per_pair * cid / n_clusters;
  const long long t_end = tiles_per_pair * (cid + 1) / n_clusters;
  constexpr uint16_t kClusterMask = (1u << CLUSTER) - 1;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    uint32_t stage = 0, phase = 0, uphase = 0;
    long long cur_bp = -1;
    for (int b = 0; b < p.B; ++b)
      for (long long t = t_begin; t < t_end; ++t) {
        const int rg = (int)(t / p.n_qblocks);
        const int qb = (int)(t - (long long)rg * p.n_qblocks);
        const long long bp = (long long)b * p.n_recs + min(rg * CLUSTER + (int)crank, p.n_recs - 1);
        if (bp != cur_bp) {
          if (cur_bp >= 0) {  // all MMAs reading the previous record have completed
            ptx::mbar_wait(b_empty, uphase);
            uphase ^= 1;
          }
          ptx::mbar_expect_tx(b_full, (uint32_t)(p.KB * B_KB_BYTES));
          for (int kb = 0; kb < p.KB; ++kb)
            ptx::tma_load_2d_hint(sB + kb * B_KB_BYTES, &tmB, b_full, kb * BK, (int)(bp * BN), ptx::kEvictFirst);
          cur_bp = bp;
        }
        for (int kb = 0; kb < p.KB; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1);
          if (p.dbg & 8) {
            ptx::mbar_arrive(full_bar(stage));
          } else {
            // the whole stage (own slice + the peers' slices) lands on this CTA's barrier
            ptx::mbar_expect_tx(full_bar(stage), A_STAGE_BYTES);
            // fmap1 is re-read by every record sweep (87x at 1080p): pin it in L2
            ptx::tma_load_3d_multicast_hint(sA + stage * A_STAGE_BYTES + crank * (A_SLICE_ROWS * BK * 2), &tmA, full_bar(stage),
                                            kb * BK, qb * BM + (int)crank * A_SLICE_ROWS, b, kClusterMask, ptx::kEvictLast);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ptx::umma_idesc(/*f16*/ 0, BM, BN);
    uint32_t stage = 0, phase = 0, uphase = 0, tile = 0;
    long long cur_bp = -1;
    for (int b = 0; b < p.B; ++b)
      for (long long t = t_begin; t < t_end; ++t, ++tile) {
        const long long bp = (long long)b * p.n_recs + min((int)(t / p.n_qblocks) * CLUSTER + (int)crank, p.n_recs - 1);
        if (bp != cur_bp) {
          if (cur_bp >= 0) ptx::umma_commit(b_empty);
          ptx::mbar_wait(b_full, uphase);
          uphase ^= 1;
          cur_bp = bp;
        }
        const uint32_t acc = tile & 1, aphase = (tile >> 1) & 1;
        ptx::mbar_wait(tmem_empty(acc), aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.KB; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint64_t a_desc = ptx::umma_desc_kmajor_sw128(sA + stage * A_STAGE_BYTES);
          const uint64_t b_desc = ptx::umma_desc_kmajor_sw128(sB + kb * B_KB_BYTES);
          // dbg 64: no MMA for the last 15.5 % of the records (what pooling level 1 in the epilogue would save)
          const bool skip_mma = (p.dbg & 4) || ((p.dbg & 64) && (int)(t / p.n_qblocks) * CLUSTER >= p.n_recs * 147 / 174);
          if (!skip_mma) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)  // UMMA_K = 16 fp16 = 32 B = +2 in the descriptor's 16-byte units
              ptx::umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          }
          ptx::umma_commit_multicast(empty_bar(stage), kClusterMask);  // every producer of the cluster writes this stage
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(tmem_full(acc));
      }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int quad = warp & 3;         // TMEM lane quadrant this warp may access
    const int part = (warp - 4) >> 2;  // which share of the record's (section, half) steps
    uint32_t tile = 0;
    for (int b = 0; b < p.B; ++b) {
      const float scale = __ldg(p.epi_scale + b);
      for (long long t = t_begin; t < t_end; ++t, ++tile) {
        const int rg = (int)(t / p.n_qblocks);
        const int qb = (int)(t - (long long)rg * p.n_qblocks);
        const int rec = rg * CLUSTER + (int)crank;
        const long long bp = (long long)b * p.n_recs + rec;
        const uint32_t acc = tile & 1, aphase = (tile >> 1) & 1;
        ptx::mbar_wait(tmem_full(acc), aphase);
        ptx::tc_fence_after();
        const uint32_t t_base = tmem_base + acc * BN + ((uint32_t)(quad * 32) << 16);
        const int grp = qb * (BM / 32) + quad;  // query group of this warp; blocks are ordered [pair][patch][group]
        char* gblock = (grp < p.n_groups && rec < p.n_recs) ? p.pyr + (bp * p.n_groups + grp) * p.group_bytes : nullptr;
        epilogue_tile<OutT>(t_base, part, lane, scale, tmem_empty(acc), gblock, p.dbg);
      }
    }
  }

  ptx::tc_fence_before();
  __syncwarp();         // the single-lane roles above rejoin their warps (the cluster barrier is .aligned)
  ptx::cluster_sync();  // no CTA leaves while a peer may still multicast into it or arrive on its barriers
  if (warp == 2) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

// --------------------------------------------------------------------------------- host side
struct Workspace {
  unsigned int* amax;  // [2][B]
  float* bscale;       // [B]
  float* epi;          // [B]
  __half* f1h;         // (B,N,Cp)            A operand
  __half* f2p;         // (B,n_recs,256,Cp)   B operand record tiles
  float* f2l[PYR_MAX_LEVELS];  // pooled f2 levels 1..L-1, (B,C,h_l,w_l) fp32
};

static size_t ws_layout(int B, int C, const PyrLayout& lay, void* base, Workspace* ws) {
  const int Cp = round_up(C, BK), N = lay.H * lay.W;
  uintptr_t p = reinterpret_cast<uintptr_t>(base);
  auto take = [&](size_t bytes, size_t align) {
    p = (p + align - 1) / align * align;
    uintptr_t r = p;
    p += bytes;
    return r;
  };
  const uintptr_t amax = take((size_t)2 * B * 4, 256);
  const uintptr_t bscale = take((size_t)B * 4, 256);
  const uintptr_t epi = take((size_t)B * 4, 256);
  const uintptr_t f1h = take((size_t)B * N * Cp * 2, 1024);
  const uintptr_t f2p = take((size_t)B * lay.n_tiles * PYR_REC * Cp * 2, 1024);
  uintptr_t f2l[PYR_MAX_LEVELS] = {0, 0, 0, 0};
  for (int l = 1; l < lay.L; ++l) f2l[l] = take((size_t)B * C * (lay.H >> l) * (lay.W >> l) * 4, 256);
  if (ws) {
    ws->amax = reinterpret_cast<unsigned int*>(amax);
    ws->bscale = reinterpret_cast<float*>(bscale);
    ws->epi = reinterpret_cast<float*>(epi);
    ws->f1h = reinterpret_cast<__half*>(f1h);
    ws->f2p = reinterpret_cast<__half*>(f2p);
    for (int l = 0; l < PYR_MAX_LEVELS; ++l) ws->f2l[l] = reinterpret_cast<float*>(f2l[l]);
  }
  return p - reinterpret_cast<uintptr_t>(base);
}

template <typename OutT>
static int launch_gemm(const Workspace& ws, int B, int C, const PyrLayout& lay, void* pyr, cudaStream_t st, cudaEvent_t ev0,
                       cudaEvent_t ev1) {
  const int N = lay.H * lay.W, Cp = round_up(C, BK);
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[3] = {(cuuint64_t)Cp, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t str[2] = {(cuuint64_t)Cp * 2, (cuuint64_t)N * Cp * 2};
    cuuint32_t box[3] = {BK, A_SLICE_ROWS, 1};
    int r = make_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, ws.f1h, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) return r;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Cp, (cuuint64_t)B * lay.n_tiles * PYR_REC};
    cuuint64_t str[1] = {(cuuint64_t)Cp * 2};
    cuuint32_t box[2] = {BK, BN};
    int r = make_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, ws.f2p, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) return r;
  }
  GemmParams p;
  p.B = B;
  p.N = N;
  p.KB = Cp / BK;
  p.L = lay.L;
  p.n_recs = lay.n_tiles;
  p.n_qblocks = cdiv(N, BM);
  p.n_groups = lay.n_groups;
  p.group_bytes = lay.group_bytes;
  p.epi_scale = ws.epi;
  p.pyr = static_cast<char*>(pyr);
  // The profiling switches exist only in builds made with -DEZB_PROFILING_HOOKS (tools/gemm_dbg.py builds one under
  // gpurun_out/); the shipped library ignores the environment, so nothing outside can make the product skip work.
#ifdef EZB_PROFILING_HOOKS
  const char* dbg_env = getenv("EZB_GEMM_DEBUG");
  p.dbg = dbg_env ? atoi(dbg_env) : 0;
#else
  p.dbg = 0;
#endif

  auto kern = corr_pyramid_gemm_kernel<OutT>;
  EZB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  // persistent grid = as many clusters as are co-resident (a GPC with an odd SM count leaves one SM out)
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3((unsigned)(num_sms() / CLUSTER * CLUSTER), 1, 1);
  cfg.blockDim = dim3(NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int max_clusters = 0;
  EZB_CUDA_OK(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
  if (max_clusters < 1) return fail(EZB_ERR_CUDA, "no cluster of %d CTAs with %d bytes of shared memory fits on this device", CLUSTER, SMEM_BYTES);
  const long long tiles_per_pair = (long long)cdiv(p.n_recs, CLUSTER) * p.n_qblocks;
  long long n_clusters = std::min<long long>(max_clusters, num_sms() / CLUSTER);
  if (n_clusters > tiles_per_pair) n_clusters = tiles_per_pair;
  cfg.gridDim = dim3((unsigned)(n_clusters * CLUSTER), 1, 1);
  if (ev0) EZB_CUDA_OK(cudaEventRecord(ev0, st));
  EZB_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, p));
  if (ev1) EZB_CUDA_OK(cudaEventRecord(ev1, st));
  return EZB_OK;
}

}  // namespace k1

namespace prep {
int launch_absmax(const float* t0, const float* t1, long long per_b, unsigned int* amax, int B, cudaStream_t st) {
  EZB_REQUIRE(B <= 65535, "batch too large for one launch (split it)");
  EZB_CUDA_OK(cudaMemsetAsync(amax, 0, (size_t)2 * B * 4, st));
  // at least 16 float4 per thread; large inputs: ~8 resident blocks per SM in total walk the tensors grid-stride with 8 loads in
  // flight per thread (a block per 64 KB was bound by the block launch rate: 93 -> 81 us for 8 pairs of 1080p features)
  int gx = (int)std::min<long long>(cdiv64(per_b, 256 * 64), std::max(1, num_sms() * 8 / (2 * B)));
  if (gx < 1) gx = 1;
  k1::absmax_kernel<<<dim3(gx, B, 2), 256, 0, st>>>(t0, t1, per_b, amax, B);
  EZB_LAUNCH_OK();
  return EZB_OK;
}
int launch_convert_kmajor(const float* src, __half* dst, const unsigned int* amax, int B, int C, int Cp, int N, cudaStream_t st) {
  EZB_REQUIRE(B <= 65535 && Cp % 64 == 0, "bad arguments");
  k1::convert_kmajor_kernel<<<dim3(cdiv(N, 32), Cp / 64, B), 256, 0, st>>>(src, dst, nullptr, nullptr, amax, B, C, Cp, N);
  EZB_LAUNCH_OK();
  return EZB_OK;
}
int launch_convert_kmajor2(const float* src0, __half* dst0, const float* src1, __half* dst1, const unsigned int* amax, int B, int C,
                           int Cp, int N, cudaStream_t st) {
  EZB_REQUIRE(2 * B <= 65535 && Cp % 64 == 0, "bad arguments");
  k1::convert_kmajor_kernel<<<dim3(cdiv(N, 32), Cp / 64, 2 * B), 256, 0, st>>>(src0, dst0, src1, dst1, amax, B, C, Cp, N);
  EZB_LAUNCH_OK();
  return EZB_OK;
}
int launch_pool2x2(const float* src, float* dst, long long planes, int hs, int ws, long long src_plane, int hd, int wd,
                   long long dst_plane, cudaStream_t st) {
  const long long total = planes * hd * wd;
  const int gx = (int)std::max<long long>(1, std::min<long long>(cdiv64(total, 256), 148 * 64));
  k1::pool2x2_kernel<<<gx, 256, 0, st>>>(src, dst, planes, hs, ws, src_plane, hd, wd, dst_plane);
  EZB_LAUNCH_OK();
  return EZB_OK;
}
}  // namespace prep
}  // namespace ezb

using namespace ezb;

extern "C" int ezb_corr_pyramid_layout(int H, int W, int L, int dtype, int* h, int* w, int* sub_base, int* n_tiles, int* n_groups,
                                       int64_t* group_bytes, int64_t* pair_bytes) {
  EZB_REQUIRE(H > 0 && W > 0 && L >= 1, "bad pyramid shape H=%d W=%d L=%d", H, W, L);
  EZB_REQUIRE(dtype == EZB_F32 || dtype == EZB_F16, "bad dtype %d", dtype);
  if (L > PYR_MAX_LEVELS) return fail(EZB_ERR_UNSUPPORTED, "num_levels %d > %d is not supported by the fused pyramid", L, PYR_MAX_LEVELS);
  PyrLayout lay;
  EZB_REQUIRE(pyr_layout(H, W, L, dtype, &lay), "feature map %dx%d too small for %d pyramid levels", H, W, L);
  for (int l = 0; l < L; ++l) {
    if (h) h[l] = H >> l;
    if (w) w[l] = W >> l;
    if (sub_base) sub_base[l] = lay.sub_base[l];
  }
  if (sub_base) sub_base[L] = lay.sub_base[PYR_MAX_LEVELS];
  if (n_tiles) *n_tiles = lay.n_tiles;
  if (n_groups) *n_groups = lay.n_groups;
  if (group_bytes) *group_bytes = lay.group_bytes;
  if (pair_bytes) *pair_bytes = lay.pair_bytes;
  return EZB_OK;
}

extern "C" int ezb_corr_pyramid_level_map(int H, int W, int L, int level, int32_t* tile, int32_t* column) {
  EZB_REQUIRE(tile && column, "null pointer");
  EZB_REQUIRE(H > 0 && W > 0 && L >= 1 && level >= 0 && level < L, "bad arguments H=%d W=%d L=%d level=%d", H, W, L, level);
  if (L > PYR_MAX_LEVELS) return fail(EZB_ERR_UNSUPPORTED, "num_levels %d > %d is not supported by the fused pyramid", L, PYR_MAX_LEVELS);
  PyrLayout lay;
  EZB_REQUIRE(pyr_layout(H, W, L, EZB_F16, &lay), "feature map %dx%d too small for %d pyramid levels", H, W, L);
  const int hl = H >> level, wl = W >> level;
  for (int y = 0; y < hl; ++y)
    for (int x = 0; x < wl; ++x) {
      const int s = lay.sub_base[level] + (y / PYR_SY) * lay.sw[level] + x / PYR_SX;
      tile[y * wl + x] = s / PYR_SUBS_PER_TILE;
      column[y * wl + x] = (s % PYR_SUBS_PER_TILE) * PYR_SUB + (y % PYR_SY) * PYR_SX + x % PYR_SX;
    }
  return EZB_OK;
}

extern "C" int64_t ezb_corr_pyramid_record_offset(int dtype, int lane, int column) {
  if ((dtype != EZB_F32 && dtype != EZB_F16) || lane < 0 || lane >= PYR_QG || column < 0 || column >= PYR_REC) return -1;
  return pyr_record_offset(dtype == EZB_F16 ? 2 : 4, lane, column);
}

extern "C" int ezb_corr_pyramid_workspace_bytes(int B, int C, int H, int W, int L, size_t* bytes) {
  EZB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && L >= 1 && bytes, "bad arguments");
  if (L > PYR_MAX_LEVELS) return fail(EZB_ERR_UNSUPPORTED, "num_levels %d > %d is not supported by the fused pyramid", L, PYR_MAX_LEVELS);
  PyrLayout lay;
  EZB_REQUIRE(pyr_layout(H, W, L, EZB_F16, &lay), "feature map %dx%d too small for %d pyramid levels", H, W, L);
  *bytes = k1::ws_layout(B, C, lay, nullptr, nullptr) + 1024;
  return EZB_OK;
}

extern "C" int ezb_corr_pyramid_fwd_ex(const float* fmap1, const float* fmap2, int B, int C, int H, int W, int L, int dtype,
                                       void* pyramid, float* deq_scale, void* workspace, size_t workspace_bytes,
                                       void* stream, void* ev_gemm_start, void* ev_gemm_end) {
  EZB_REQUIRE(fmap1 && fmap2 && pyramid && deq_scale && workspace, "null pointer");
  EZB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && L >= 1, "bad shape B=%d C=%d H=%d W=%d L=%d", B, C, H, W, L);
  EZB_REQUIRE(dtype == EZB_F32 || dtype == EZB_F16, "bad dtype %d", dtype);
  if (round_up(C, k1::BK) > k1::KB_MAX * k1::BK)
    return fail(EZB_ERR_UNSUPPORTED, "all-pairs correlation supports C <= %d (got %d)", k1::KB_MAX * k1::BK, C);
  if (L > PYR_MAX_LEVELS) return fail(EZB_ERR_UNSUPPORTED, "num_levels %d > %d is not supported by the fused pyramid", L, PYR_MAX_LEVELS);
  PyrLayout lay;
  EZB_REQUIRE(pyr_layout(H, W, L, dtype, &lay), "feature map %dx%d too small for %d pyramid levels", H, W, L);
  EZB_REQUIRE((reinterpret_cast<uintptr_t>(pyramid) & 127) == 0, "pyramid buffer must be 128-byte aligned");
  EZB_REQUIRE((long long)B * lay.n_tiles * PYR_REC < (1ll << 31), "batch too large for one launch (split it)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int N = H * W, Cp = round_up(C, k1::BK);

  // carve the workspace from a 1024-aligned base
  uintptr_t wbase = (reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023);
  k1::Workspace ws;
  const size_t need = k1::ws_layout(B, C, lay, reinterpret_cast<void*>(wbase), &ws) + (wbase - reinterpret_cast<uintptr_t>(workspace));
  if (need > workspace_bytes) return fail(EZB_ERR_WORKSPACE, "workspace too small: need %zu, got %zu", need, workspace_bytes);

  if (int r = prep::launch_absmax(fmap1, fmap2, (long long)C * N, ws.amax, B, st)) return r;
  k1::scales_kernel<<<cdiv(B, 128), 128, 0, st>>>(ws.amax, B, C, dtype, ws.bscale, ws.epi, deq_scale);
  EZB_LAUNCH_OK();
  if (int r = prep::launch_convert_kmajor(fmap1, ws.f1h, ws.amax, B, C, Cp, N, st)) return r;
  static const bool split_prep = getenv("EZB_PYRAMID_SPLIT_PREP") != nullptr;  // A/B switch: round-1 pool + pack launches
  if (!split_prep) {
    k1::PoolPackParams pp;
    pp.f2 = fmap2; pp.H = H; pp.W = W; pp.C = C; pp.Cp = Cp; pp.L = L; pp.n_recs = lay.n_tiles; pp.epl = 128 / lay.es;
    pp.rx = pp.ry = 1;
    for (int l = 0; l < PYR_MAX_LEVELS; ++l) {
      pp.h[l] = l < L ? H >> l : 0;
      pp.w[l] = l < L ? W >> l : 0;
      pp.sw[l] = lay.sw[l];
      pp.nsy[l] = l < L ? cdiv(H >> l, PYR_SY) : 0;
      pp.sub_base[l] = lay.sub_base[l];
      if (l < L) {  // regions needed to cover every subtile of this level: a region holds (16 >> l)^2 of its pixels
        pp.rx = std::max(pp.rx, cdiv(pp.sw[l] * PYR_SX, k1::PP_R >> l));
        pp.ry = std::max(pp.ry, cdiv(pp.nsy[l] * PYR_SY, k1::PP_R >> l));
      }
    }
    pp.sub_base[PYR_MAX_LEVELS] = lay.sub_base[PYR_MAX_LEVELS];
    pp.bscale = ws.bscale;
    pp.bp = ws.f2p;
    EZB_REQUIRE(B <= 65535 && Cp / k1::PP_CH <= 65535, "batch too large for one launch (split it)");
    const dim3 grid(pp.rx * pp.ry, Cp / k1::PP_CH, B);
    const bool vec = W % 4 == 0 && (reinterpret_cast<uintptr_t>(fmap2) & 15) == 0;
    auto kern = vec ? k1::pool_pack_kernel<true> : k1::pool_pack_kernel<false>;
    EZB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1::PP_SMEM));
    kern<<<grid, k1::PP_NT, k1::PP_SMEM, st>>>(pp);
    EZB_LAUNCH_OK();
  } else {
    k1::PackParams pk;
    for (int l = 0; l < PYR_MAX_LEVELS; ++l) {
      pk.lvl[l] = l == 0 ? fmap2 : ws.f2l[l];
      pk.h[l] = H >> l;
      pk.w[l] = W >> l;
      pk.sw[l] = lay.sw[l];
      pk.sub_base[l] = lay.sub_base[l];
    }
    pk.sub_base[PYR_MAX_LEVELS] = lay.sub_base[PYR_MAX_LEVELS];
    pk.L = L; pk.C = C; pk.Cp = Cp; pk.n_recs = lay.n_tiles; pk.epl = 128 / lay.es;
    pk.bscale = ws.bscale;
    pk.bp = ws.f2p;
    for (int l = 1; l < L; ++l) {
      const long long total = (long long)B * C * pk.h[l] * pk.w[l];
      const int gx = (int)std::min<long long>(cdiv64(total, 256), 148 * 64);
      k1::pool2x2_kernel<<<gx, 256, 0, st>>>(pk.lvl[l - 1], ws.f2l[l], (long long)B * C, pk.h[l - 1], pk.w[l - 1],
                                             (long long)pk.h[l - 1] * pk.w[l - 1], pk.h[l], pk.w[l], (long long)pk.h[l] * pk.w[l]);
      EZB_LAUNCH_OK();
    }
    k1::pack_record_operand_kernel<<<dim3(lay.n_tiles, Cp / k1::PACK_CH, B), 256, 0, st>>>(pk);
    EZB_LAUNCH_OK();
  }

  cudaEvent_t ev0 = static_cast<cudaEvent_t>(ev_gemm_start), ev1 = static_cast<cudaEvent_t>(ev_gemm_end);
  return dtype == EZB_F16 ? k1::launch_gemm<__half>(ws, B, C, lay, pyramid, st, ev0, ev1)
                          : k1::launch_gemm<float>(ws, B, C, lay, pyramid, st, ev0, ev1);
}

extern "C" int ezb_corr_pyramid_fwd(const float* fmap1, const float* fmap2, int B, int C, int H, int W, int L, int dtype,
                                    void* pyramid, float* deq_scale, void* workspace, size_t workspace_bytes, void* stream) {
  return ezb_corr_pyramid_fwd_ex(fmap1, fmap2, B, C, H, W, L, dtype, pyramid, deq_scale, workspace, workspace_bytes, stream,
                                 nullptr, nullptr);
}
