// K4 on CUDA cores, register-tiled: the PWC-Net shape of the local correlation (9x9 window, stride 1, no dilation) for
// large maps with few channels, exact fp32.
//
// Replaces iter_spatial_correlation_sample (reference ezflow/similarity/correlation/sampler.py:96-129) for
//   stride 1, padding 0, dilation_patch 1, square odd patch P <= 9:
//   out[bg,pi,pj,y,x] = act( scale * sum_c in1[bg,c,y,x] * in2[bg,c, y + pi - md, x + pj - md] ),  md = (P-1)/2, 0 outside.
//
// Why not the tensor-core kernel (local_corr_tc.cu) here: PWC-Net's two finest levels have 32 / 64 channels, i.e. only
// 2.6-5.2 kFMA per output pixel.  The tcgen05 path computes a (queries x targets) band of which 1/5 is kept, after a
// conversion pass that moves as many bytes as the whole problem, and spends its time extracting the band from TMEM
// (138 us for the (8,32,112,256) level: 58 us conversion + 80 us kernel).  On CUDA cores the same level is 0.59 GFMA:
// 16 us at the FMA peak.  The products have to come out of shared memory at >= 16 FMA per 16-byte load to get there;
// the tiling below reaches 10.8 (shared-memory-bound at ~2/3 of the FMA peak), needs no workspace and no operand
// scaling, and is bit-comparable to the fp32 reference up to summation order.
//
//   block  : 32 x 8 output pixels of one image, 64 * ceil(P/3) threads.
//   thread : 4 consecutive pixels (x) x 3 displacement rows (pi) x all P displacement columns: 12 * P accumulators.
//            Per channel it reads its in1 quad (one LDS.128) and, per displacement row, the 12 in2 values
//            x - 4 .. x + 7 of that row (three LDS.128, 128 contiguous bytes per quarter-warp: conflict-free)
//            -> 36 FMAs per three loads.
//   stages : CC = 8 channels of the in1 tile (8 x 32) and of the in2 halo tile (16 x 40), fetched with 16-byte cp.async
//            (zero-filled outside the image / beyond C) through a three-stage ring with one barrier per stage; a thread
//            copies the same (row, quad) chunk for every channel, so its addresses are decoded once.
//   stores : float4 per (pi, pj): 8 lanes write 128 contiguous bytes of the (B, G, P, P, H, W) output.
// The warp-fused entry point (ezb_warp_local_corr_fwd) uses it too where it is the faster kernel: K5 warps in2 into the
// workspace first (a halo-warping variant of this kernel was 3x slower than that: every halo pixel is warped 2.5 times).
#include "ezb_common.cuh"

#include <algorithm>
#include <cstdlib>

namespace ezb {
namespace k4rt {

constexpr int TW = 32;   // output tile width
constexpr int CC = 8;    // channels per stage
constexpr int MDA = 4;   // the halo starts MDA columns left of the tile (aligned for LDS.128), MDA >= md

// TH output rows per tile, PR displacement rows per thread.  <8, 3>: the large-map configuration described above.
// <4, 1>: small maps (fewer than one 32 x 8 tile per SM): half the tile, a third of the accumulators, i.e. 6x the threads
// per pixel, at 9 instead of 10.8 FMAs per shared-memory load.
template <int P, int TH_, int PR_>
struct Cfg {
  static constexpr int TH = TH_, PR = PR_;
  static constexpr int MD = (P - 1) / 2;
  static constexpr int NPG = (P + PR - 1) / PR;       // displacement-row groups
  static constexpr int NT = (TW / 4) * TH * NPG;      // threads
  static constexpr int HALO_W = MDA + TW + 4;         // x0 - 4 .. x0 + 35 (MD <= 4)
  static constexpr int HALO_H = TH + 2 * MD;
  static constexpr int IN2_STAGE = CC * HALO_H * HALO_W;  // floats
  static constexpr int IN1_STAGE = CC * TH * TW;
  static constexpr int STAGE = IN2_STAGE + IN1_STAGE;
  static constexpr int NSTAGE = 3;
  static constexpr size_t SMEM = (size_t)NSTAGE * STAGE * sizeof(float);
  // staging geometry of the vector path: a thread copies the same (row, quad) chunk of the halo for every channel of a stage
  static constexpr int N2PC = HALO_H * (HALO_W / 4);        // 16-byte chunks of one channel's halo
  static constexpr int J2 = (N2PC + NT - 1) / NT;           // ... per thread
  static constexpr int N1PC = TH * (TW / 4);                // chunks of one channel's in1 tile (64)
  static constexpr int CPP = NT / N1PC;                     // in1 channels covered by one pass of the block
  static_assert(P % 2 == 1 && MD <= MDA, "patch size");
};

struct Params {
  const float* in1;
  const float* in2;
  float* out;
  int BG, G, C, H, W;
  float scale, slope;
  long long out_bstride;
};

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t nbytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src, uint32_t nbytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
}

// VEC: W % 4 == 0 and 16-byte aligned tensors, so an aligned quad of pixels is entirely inside or outside the image
template <int P, int TH, int PR, bool VEC>
__global__ void __launch_bounds__(Cfg<P, TH, PR>::NT, 2) local_corr_rt_kernel(const Params Q) {
  using K = Cfg<P, TH, PR>;
  constexpr int MD = K::MD, HALO_W = K::HALO_W, HALO_H = K::HALO_H, NT = K::NT;
  extern __shared__ __align__(16) float sm_rt[];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, bg = blockIdx.z;
  const long long HW = (long long)Q.H * Q.W;
  const float* in1 = Q.in1 + (long long)bg * Q.C * HW;
  const float* in2 = Q.in2 + (long long)bg * Q.C * HW;
  const uint32_t sm_base = (uint32_t)__cvta_generic_to_shared(sm_rt);

  // ---- vector path: the (row, quad) chunks this thread copies in every stage, decoded once
  [[maybe_unused]] int goff2[K::J2];
  [[maybe_unused]] unsigned ok2 = 0;
  [[maybe_unused]] const int ccs1 = tid / K::N1PC, ch1 = tid - (tid / K::N1PC) * K::N1PC;
  [[maybe_unused]] int goff1 = 0;
  [[maybe_unused]] bool ok1 = false;
  if constexpr (VEC) {
#pragma unroll
    for (int j = 0; j < K::J2; ++j) {
      const int ch = tid + j * NT, hy = ch / (HALO_W / 4), hq = ch - hy * (HALO_W / 4);
      const int y = y0 - MD + hy, x = x0 - MDA + 4 * hq;
      goff2[j] = y * Q.W + x;
      if (ch < K::N2PC && y >= 0 && y < Q.H && x >= 0 && x < Q.W) ok2 |= 1u << j;
    }
    const int ty1 = ch1 / (TW / 4), qx1 = ch1 - ty1 * (TW / 4);
    goff1 = (y0 + ty1) * Q.W + x0 + 4 * qx1;
    ok1 = y0 + ty1 < Q.H && x0 + 4 * qx1 < Q.W;
  }

  // ---- staging of one channel chunk into buffer `buf`
  auto stage = [&](int buf, int c0) {
    [[maybe_unused]] float* s2 = sm_rt + buf * K::STAGE;
    const uint32_t a2 = sm_base + (uint32_t)(buf * K::STAGE) * 4u, a1 = a2 + (uint32_t)K::IN2_STAGE * 4u;
    if constexpr (VEC) {
      const float* g2 = in2 + (long long)c0 * HW;
#pragma unroll
      for (int j = 0; j < K::J2; ++j) {
        const int ch = tid + j * NT;
        if (K::N2PC % NT != 0 && ch >= K::N2PC) break;
        const uint32_t dst = a2 + (uint32_t)ch * 16u;
        if ((ok2 >> j) & 1u) {
          const float* src = g2 + goff2[j];
#pragma unroll
          for (int cc = 0; cc < CC; ++cc) {
            if (c0 + cc < Q.C)
              cp_async_16(dst + (uint32_t)(cc * K::N2PC) * 16u, src + (long long)cc * HW, 16u);
            else
              *reinterpret_cast<float4*>(s2 + (cc * K::N2PC + ch) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        } else {  // outside the image: plain zero stores (ordered before the readers by the stage barrier)
#pragma unroll
          for (int cc = 0; cc < CC; ++cc) *reinterpret_cast<float4*>(s2 + (cc * K::N2PC + ch) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    } else {
      constexpr int N2 = CC * HALO_H * HALO_W;
      for (int i = tid; i < N2; i += NT) {
        const int cc = i / (HALO_H * HALO_W), rem = i - cc * (HALO_H * HALO_W), hy = rem / HALO_W, hx = rem - hy * HALO_W;
        const int y = y0 - MD + hy, x = x0 - MDA + hx, c = c0 + cc;
        const bool ok = c < Q.C && y >= 0 && y < Q.H && x >= 0 && x < Q.W;
        cp_async_4(a2 + (uint32_t)i * 4u, ok ? in2 + (long long)c * HW + (long long)y * Q.W + x : in2, ok ? 4u : 0u);
      }
    }
    if constexpr (VEC) {
      const float* g1 = in1 + (long long)c0 * HW + goff1;
      float* s1 = s2 + K::IN2_STAGE;
#pragma unroll
      for (int cc = ccs1; cc < CC; cc += K::CPP) {
        if (ok1 && c0 + cc < Q.C)
          cp_async_16(a1 + (uint32_t)(cc * K::N1PC + ch1) * 16u, g1 + (long long)cc * HW, 16u);
        else
          *reinterpret_cast<float4*>(s1 + (cc * K::N1PC + ch1) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      constexpr int N1 = CC * TH * TW;
      for (int i = tid; i < N1; i += NT) {
        const int cc = i / (TH * TW), rem = i - cc * (TH * TW), ty = rem / TW, xx = rem - ty * TW;
        const int y = y0 + ty, x = x0 + xx, c = c0 + cc;
        const bool ok = c < Q.C && y < Q.H && x < Q.W;
        cp_async_4(a1 + (uint32_t)i * 4u, ok ? in1 + (long long)c * HW + (long long)y * Q.W + x : in1, ok ? 4u : 0u);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // ---- thread geometry
  const int pg = tid / K::N1PC, r64 = tid - pg * K::N1PC, qx = r64 & 7, ty = r64 >> 3;
  const int pi0 = pg * PR;
  float acc[PR][P][4];
#pragma unroll
  for (int r = 0; r < PR; ++r)
#pragma unroll
    for (int pj = 0; pj < P; ++pj)
#pragma unroll
      for (int px = 0; px < 4; ++px) acc[r][pj][px] = 0.f;

  // three-stage cp.async pipeline, one barrier per stage: the barrier of iteration k orders every thread's reads of buffer
  // (k - 1) % 3 before the copies of stage k + 2 that overwrite it
  const int nk = (Q.C + CC - 1) / CC;
  stage(0, 0);
  if (nk > 1) stage(1, CC);
  for (int k = 0; k < nk; ++k) {
    if (k + 1 < nk) {
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (k + 2 < nk) stage((k + 2) % K::NSTAGE, (k + 2) * CC);
    const float* s2 = sm_rt + (k % K::NSTAGE) * K::STAGE;
    const float* s1 = s2 + K::IN2_STAGE;
    // the loads of displacement row r + 1 (or of the next channel's first row and in1 quad) are issued before the 36 FMAs of row r
    const float4* rowp = reinterpret_cast<const float4*>(s2 + (ty + pi0) * HALO_W + 4 * qx);
    const float4* ap = reinterpret_cast<const float4*>(s1 + ty * TW + 4 * qx);
    constexpr int ROW4 = HALO_W / 4, CH4 = HALO_H * HALO_W / 4, A4 = TH * TW / 4;
    const int nrows = (P % PR != 0 && pi0 + PR > P) ? P - pi0 : PR;  // the last group of P = 5, 7 has fewer rows
    float4 a = ap[0], v0 = rowp[0], v1 = rowp[1], v2 = rowp[2];
#pragma unroll 2
    for (int cc = 0; cc < CC; ++cc) {
      const float av[4] = {a.x, a.y, a.z, a.w};
      float4 an = a;
#pragma unroll
      for (int r = 0; r < PR; ++r) {
        if (P % PR != 0 && r >= nrows) break;
        const float v[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
        const bool last_row = (P % PR != 0) ? r + 1 >= nrows : r + 1 == PR;
        if (!last_row) {
          const float4* nx = rowp + cc * CH4 + (r + 1) * ROW4;
          v0 = nx[0]; v1 = nx[1]; v2 = nx[2];
        } else if (cc + 1 < CC) {
          const float4* nx = rowp + (cc + 1) * CH4;
          v0 = nx[0]; v1 = nx[1]; v2 = nx[2];
          an = ap[(cc + 1) * A4];
        }
#pragma unroll
        for (int pj = 0; pj < P; ++pj)
#pragma unroll
          for (int px = 0; px < 4; ++px) acc[r][pj][px] = fmaf(av[px], v[MDA - MD + pj + px], acc[r][pj][px]);
      }
      a = an;
    }
  }

  // ---- epilogue
  const int y = y0 + ty, x = x0 + 4 * qx;
  if (y >= Q.H || x >= Q.W) return;
  const int b = bg / Q.G, g = bg - b * Q.G;
  float* dst = Q.out + (long long)b * Q.out_bstride + ((long long)g * P * P + (long long)pi0 * P) * HW + (long long)y * Q.W + x;
  const bool plain = Q.slope == 1.f;
#pragma unroll
  for (int r = 0; r < PR; ++r) {
    if (pi0 + r >= P) break;
#pragma unroll
    for (int pj = 0; pj < P; ++pj, dst += HW) {
      float o[4];
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        const float v = acc[r][pj][px] * Q.scale;
        o[px] = (plain || v >= 0.f) ? v : v * Q.slope;
      }
      if constexpr (VEC) {
        __stcs(reinterpret_cast<float4*>(dst), make_float4(o[0], o[1], o[2], o[3]));
      } else {
#pragma unroll
        for (int px = 0; px < 4; ++px)
          if (x + px < Q.W) dst[px] = o[px];
      }
    }
  }
}

template <int P, int TH, int PR, bool VEC>
static int launch(const Params& Q, cudaStream_t st) {
  using K = Cfg<P, TH, PR>;
  auto kern = local_corr_rt_kernel<P, TH, PR, VEC>;
  EZB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM));
  dim3 grid(cdiv(Q.W, TW), cdiv(Q.H, TH), Q.BG);
  kern<<<grid, K::NT, K::SMEM, st>>>(Q);
  EZB_LAUNCH_OK();
  return EZB_OK;
}

// ------------------------------------------------------------------------------------ strided forward (DCVNet's fine scale)
// MatryoshkaDilatedCostVolumeList samples its stride-2 feature maps every 4 pixels (learnable_cost.py:420-434: stride =
// 8 // feat_stride), one dilation of 1:  out[bg,pi,pj,y,x] = sum_c in1[bg,c,S*y,S*x] * in2[bg,c,S*y + pi - md, S*x + pj - md].
// That is 81 * C FMAs per *sixteen* input pixels: the op is bound by reading in2 once (151 MB at BASELINE config 4), and the
// tensor-core path's conversion pass alone moved more than that.  Same scheme as local_corr_rt_kernel with the window walking
// S pixels per query:
//   block  : 16 x 4 queries; thread = 4 consecutive queries x 1 displacement row x P displacement columns (4 * P accumulators).
//   stage  : 4 channels of the (S*16 + 8) x (S*3 + P) in2 halo (16-byte cp.async, addresses decoded once per block) and of
//            the 64 in1 query pixels (4-byte gathers), three-stage ring, one barrier per stage.
//   per channel and thread: the in1 quad (one LDS.128), S*3 + P (+ alignment) in2 values as LDS.128, 4 * P FMAs.
//   banks  : a thread's row segment starts 4*S floats after its neighbour's, so the eight lanes of a quarter-warp are dealt
//            QXL query quads x PIL displacement rows and the halo row pitch is 3 (mod 8) 16-byte groups: S = 4 -> lanes
//            (2 quads x 4 rows) hit groups {0,3,6,1} + {0,4}; S = 2 -> (4 quads x 2 rows) hit {0,2,4,6} + {0,3}: conflict-free.
//            Displacement rows are padded to a multiple of PIL; the padding threads only stage.
constexpr int SQW = 16, SQH = 4, SCC = 4;

template <int P, int S>
struct SCfg {
  static constexpr int MD = (P - 1) / 2;
  static constexpr int QXL = S == 4 ? 2 : 4;                    // query quads per quarter-warp
  static constexpr int PIL = 8 / QXL;                           // displacement rows per quarter-warp
  static constexpr int PROWS = ((P + PIL - 1) / PIL) * PIL;
  static constexpr int NQ = (S * SQW + 8) / 4;                  // 16-byte chunks fetched per halo row: S*x0 - 4 .. S*x0 + S*16 + 3
  static constexpr int HALO_W = S == 4 ? 76 : 44;               // row pitch in floats (19 / 11 groups = 3 mod 8)
  static constexpr int HALO_H = S * (SQH - 1) + P;
  static constexpr int NV = ((MDA - MD) + S * 3 + P + 3) & ~3;  // in2 values a thread reads per row
  static constexpr int NT = (SQW / 4) * SQH * PROWS;
  static constexpr int N2PC = HALO_H * NQ;                      // chunks of one channel's halo
  static constexpr int J2 = (N2PC + NT - 1) / NT;
  static constexpr int IN2_STAGE = SCC * HALO_H * HALO_W, IN1_STAGE = SCC * SQH * SQW, STAGE = IN2_STAGE + IN1_STAGE;
  static constexpr int NSTAGE = 3;
  static constexpr size_t SMEM = (size_t)NSTAGE * STAGE * sizeof(float);
  static_assert((S == 2 || S == 4) && S * 12 + NV <= HALO_W && 4 * NQ <= HALO_W && MD <= MDA && (HALO_W / 4) % 8 == 3, "geometry");
};

struct SParams {
  const float* in1;
  const float* in2;
  float* out;
  int BG, G, C, H, W, oh, ow;
  float scale, slope;
  long long out_bstride;
};

template <int P, int S>
__global__ void __launch_bounds__(SCfg<P, S>::NT, 2) local_corr_rt_strided_kernel(const SParams Q) {
  using K = SCfg<P, S>;
  constexpr int MD = K::MD, HALO_W = K::HALO_W, HALO_H = K::HALO_H, NT = K::NT, NQ = K::NQ, NV = K::NV;
  extern __shared__ __align__(16) float sm_srt[];
  const uint32_t sm_base = (uint32_t)__cvta_generic_to_shared(sm_srt);
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * SQW, y0 = blockIdx.y * SQH, bg = blockIdx.z;  // first query of the tile
  const long long HW = (long long)Q.H * Q.W;
  const float* in1 = Q.in1 + (long long)bg * Q.C * HW;
  const float* in2 = Q.in2 + (long long)bg * Q.C * HW;

  // the halo chunks and in1 pixels this thread copies in every stage
  int goff2[K::J2], soff2[K::J2];
  unsigned ok2 = 0, in2range = 0;
#pragma unroll
  for (int j = 0; j < K::J2; ++j) {
    const int ch = tid + j * NT, hy = ch / NQ, hq = ch - hy * NQ;
    const int y = S * y0 - MD + hy, x = S * x0 - MDA + 4 * hq;
    goff2[j] = y * Q.W + x;
    soff2[j] = hy * HALO_W + 4 * hq;
    if (ch < K::N2PC) {
      in2range |= 1u << j;
      if (y >= 0 && y < Q.H && x >= 0 && x < Q.W) ok2 |= 1u << j;
    }
  }
  constexpr int J1 = (K::IN1_STAGE + NT - 1) / NT;
  int goff1[J1];
  unsigned ok1 = 0, in1range = 0;
#pragma unroll
  for (int j = 0; j < J1; ++j) {
    const int i = tid + j * NT, cc = i / (SQH * SQW), rem = i - cc * (SQH * SQW), qy = rem / SQW, qxx = rem - qy * SQW;
    const int y = S * (y0 + qy), x = S * (x0 + qxx);
    goff1[j] = y * Q.W + x;  // plus (c0 + cc) * HW at staging time
    if (i < K::IN1_STAGE) {
      in1range |= 1u << j;
      if (y < Q.H && x < Q.W) ok1 |= 1u << j;
    }
  }

  auto stage = [&](int buf, int c0) {
    float* s2 = sm_srt + buf * K::STAGE;
    float* s1 = s2 + K::IN2_STAGE;
    const uint32_t a2 = sm_base + (uint32_t)(buf * K::STAGE) * 4u, a1 = a2 + (uint32_t)K::IN2_STAGE * 4u;
    const float* g2 = in2 + (long long)c0 * HW;
#pragma unroll
    for (int j = 0; j < K::J2; ++j) {
      if (!((in2range >> j) & 1u)) break;
      const bool ok = (ok2 >> j) & 1u;
#pragma unroll
      for (int cc = 0; cc < SCC; ++cc) {
        if (ok && c0 + cc < Q.C)
          cp_async_16(a2 + (uint32_t)(cc * HALO_H * HALO_W + soff2[j]) * 4u, g2 + (long long)cc * HW + goff2[j], 16u);
        else
          *reinterpret_cast<float4*>(s2 + cc * HALO_H * HALO_W + soff2[j]) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int j = 0; j < J1; ++j) {
      if (!((in1range >> j) & 1u)) break;
      const int i = tid + j * NT, cc = i / (SQH * SQW);
      if (((ok1 >> j) & 1u) && c0 + cc < Q.C)
        cp_async_4(a1 + (uint32_t)i * 4u, in1 + (long long)(c0 + cc) * HW + goff1[j], 4u);
      else
        s1[i] = 0.f;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // thread geometry: lane bits (low to high) = displacement row within its group, query quad within its group
  const int lane8 = tid & 7, upper = tid >> 3;
  const int pi_lo = lane8 % K::PIL, qx_lo = lane8 / K::PIL;
  const int qx_hi = upper % (4 / K::QXL), t2 = upper / (4 / K::QXL), ty = t2 % SQH, pi_hi = t2 / SQH;
  const int pi = pi_hi * K::PIL + pi_lo, qx = qx_hi * K::QXL + qx_lo;
  const bool active = pi < P;
  float acc[P][4];
#pragma unroll
  for (int pj = 0; pj < P; ++pj)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[pj][q] = 0.f;

  const int nk = (Q.C + SCC - 1) / SCC;
  stage(0, 0);
  if (nk > 1) stage(1, SCC);
  for (int k = 0; k < nk; ++k) {
    if (k + 1 < nk) {
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (k + 2 < nk) stage((k + 2) % K::NSTAGE, (k + 2) * SCC);
    if (!active) continue;
    const float* s2 = sm_srt + (k % K::NSTAGE) * K::STAGE;
    const float* s1 = s2 + K::IN2_STAGE;
#pragma unroll
    for (int cc = 0; cc < SCC; ++cc) {
      const float4 a = *reinterpret_cast<const float4*>(s1 + (cc * SQH + ty) * SQW + 4 * qx);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float4* row = reinterpret_cast<const float4*>(s2 + (cc * HALO_H + S * ty + pi) * HALO_W + S * 4 * qx);
      float v[NV];
#pragma unroll
      for (int j = 0; j < NV / 4; ++j) {
        const float4 t = row[j];
        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
      }
#pragma unroll
      for (int pj = 0; pj < P; ++pj)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[pj][q] = fmaf(av[q], v[MDA - MD + pj + S * q], acc[pj][q]);
    }
  }

  const int y = y0 + ty, x = x0 + 4 * qx;
  if (!active || y >= Q.oh || x >= Q.ow) return;
  const long long ohw = (long long)Q.oh * Q.ow;
  const int b = bg / Q.G, g = bg - b * Q.G;
  float* dst = Q.out + (long long)b * Q.out_bstride + ((long long)g * P * P + (long long)pi * P) * ohw + (long long)y * Q.ow + x;
  const bool plain = Q.slope == 1.f;
#pragma unroll
  for (int pj = 0; pj < P; ++pj, dst += ohw) {
    float o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float t = acc[pj][q] * Q.scale;
      o[q] = (plain || t >= 0.f) ? t : t * Q.slope;
    }
    __stcs(reinterpret_cast<float4*>(dst), make_float4(o[0], o[1], o[2], o[3]));  // ow % 4 == 0: the quad is inside or outside
  }
}

template <int P, int S>
static int launch_strided(const SParams& Q, cudaStream_t st) {
  using K = SCfg<P, S>;
  auto kern = local_corr_rt_strided_kernel<P, S>;
  EZB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM));
  kern<<<dim3(cdiv(Q.ow, SQW), cdiv(Q.oh, SQH), Q.BG), K::NT, K::SMEM, st>>>(Q);
  EZB_LAUNCH_OK();
  return EZB_OK;
}

// ------------------------------------------------------------------------------------ strided adjoints (stride 4: DCVNet's fine scale)
// Autograd of the strided forward above (sampler.py:96-129 with stride 4, learnable_cost.py:420-434), gather form, exact fp32:
//   din1[c, 4y, 4x]   = scale * sum_{pi,pj} dout[pi,pj,y,x] * in2[c, 4y + pi - md, 4x + pj - md]          (0 off the query lattice)
//   din2[c, 4Y+i, 4X+j] = scale * sum over the (pi,pj) with pi - md = i, pj - md = j (mod 4) of
//                         dout[pi,pj, Y - (pi-md-i)/4, X - (pj-md-j)/4] * in1[c, 4(Y - ..), 4(X - ..)]
// Both are "one query cell = P*P products per channel": thread = cell (y, x) (lane = x), its P*P upstream gradients live in
// registers for the whole channel loop.  din1: per channel and window row three 16-byte loads (4x-4 .. 4x+7: each perfectly
// coalesced over the warp, two of the three L1 hits) -> P FMAs; the thread writes the whole 4 x 4 block of din1 (the value and
// 15 zeros) as four 16-byte stores, so no memset pass is needed.  din2: every (pi,pj) feeds exactly one of the cell's 4 x 4
// outputs from one of the 3 x 3 neighbouring queries: 9 scalar in1 loads, P*P FMAs, four 16-byte stores per channel.
// The round-1 gather kernels tested every (pi,pj) of every output element for divisibility: 5.5 ms at BASELINE config 4.
struct SBwdParams {
  const float* in1;
  const float* in2;
  const float* dout;
  float* din1;
  float* din2;
  int BG, G, C, H, W, oh, ow, nchunk, cpc;  // nchunk channel chunks of cpc channels (blockIdx.z = bg * nchunk + chunk)
  float scale;
  long long dout_bstride;
};

template <int P, bool SECOND>
__global__ void __launch_bounds__(128) local_corr_bwd_strided_kernel(const SBwdParams Q) {
  constexpr int S = 4, MD = (P - 1) / 2;
  static_assert(MD <= 4, "window");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x = blockIdx.x * 32 + lane, y = blockIdx.y * 4 + warp;
  const int bg = blockIdx.z / Q.nchunk, chunk = blockIdx.z - bg * Q.nchunk;
  if (x >= Q.ow || y >= Q.oh) return;  // no barriers below
  const int b = bg / Q.G, g = bg - b * Q.G;
  const long long ohw = (long long)Q.oh * Q.ow, HW = (long long)Q.H * Q.W;
  const float* dbase = Q.dout + (long long)b * Q.dout_bstride + (long long)g * P * P * ohw;
  const int c_begin = chunk * Q.cpc, c_end = min(Q.C, c_begin + Q.cpc);

  float dv[P][P];
  if constexpr (!SECOND) {
#pragma unroll
    for (int pi = 0; pi < P; ++pi)
#pragma unroll
      for (int pj = 0; pj < P; ++pj) dv[pi][pj] = __ldg(dbase + (long long)(pi * P + pj) * ohw + (long long)y * Q.ow + x) * Q.scale;
    const float* img = Q.in2 + ((long long)bg * Q.C + c_begin) * HW + (long long)(S * y) * Q.W + S * x;
    float* o = Q.din1 + ((long long)bg * Q.C + c_begin) * HW + (long long)(S * y) * Q.W + S * x;
    const bool left = x > 0, right = S * x + 4 < Q.W;
    for (int c = c_begin; c < c_end; ++c, img += HW, o += HW) {
      float acc3[3] = {0.f, 0.f, 0.f};  // three independent FMA chains
#pragma unroll
      for (int pi = 0; pi < P; ++pi) {
        const int yy = S * y + pi - MD;
        if (yy < 0 || yy >= Q.H) continue;
        float& acc = acc3[pi % 3];
        const float* r = img + (long long)(pi - MD) * Q.W;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 a = (MD > 0 && left) ? __ldg(reinterpret_cast<const float4*>(r - 4)) : z;
        const float4 m = __ldg(reinterpret_cast<const float4*>(r));
        const float4 e = (MD == 4 && right) ? __ldg(reinterpret_cast<const float4*>(r + 4)) : z;
        const float v[12] = {a.x, a.y, a.z, a.w, m.x, m.y, m.z, m.w, e.x, e.y, e.z, e.w};
#pragma unroll
        for (int pj = 0; pj < P; ++pj) acc = fmaf(dv[pi][pj], v[4 - MD + pj], acc);
      }
      __stcs(reinterpret_cast<float4*>(o), make_float4((acc3[0] + acc3[1]) + acc3[2], 0.f, 0.f, 0.f));
#pragma unroll
      for (int rr = 1; rr < S; ++rr)
        if (S * y + rr < Q.H) __stcs(reinterpret_cast<float4*>(o + (long long)rr * Q.W), make_float4(0.f, 0.f, 0.f, 0.f));
    }
  } else {
    // displacement pi lands on output row i = (pi - MD) mod 4 of the cell and comes from the query dy = -(pi - MD - i) / 4 cells away
    auto sub = [](int p) { const int t = p - MD; return ((t % S) + S) % S; };
    auto off = [&](int p) { const int t = p - MD; return -(t - sub(p)) / S; };
#pragma unroll
    for (int pi = 0; pi < P; ++pi)
#pragma unroll
      for (int pj = 0; pj < P; ++pj) {
        const int yq = y + off(pi), xq = x + off(pj);
        const bool ok = yq >= 0 && yq < Q.oh && xq >= 0 && xq < Q.ow;
        dv[pi][pj] = ok ? __ldg(dbase + (long long)(pi * P + pj) * ohw + (long long)yq * Q.ow + xq) * Q.scale : 0.f;
      }
    const float* img = Q.in1 + ((long long)bg * Q.C + c_begin) * HW + (long long)(S * y) * Q.W + S * x;
    float* o = Q.din2 + ((long long)bg * Q.C + c_begin) * HW + (long long)(S * y) * Q.W + S * x;
    bool okq[3][3];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) okq[dy + 1][dx + 1] = y + dy >= 0 && y + dy < Q.oh && x + dx >= 0 && x + dx < Q.ow;
    for (int c = c_begin; c < c_end; ++c, img += HW, o += HW) {
      float a[3][3];
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) a[dy + 1][dx + 1] = okq[dy + 1][dx + 1] ? __ldg(img + (long long)(S * dy) * Q.W + S * dx) : 0.f;
      float acc[S][S];
#pragma unroll
      for (int i = 0; i < S; ++i)
#pragma unroll
        for (int j = 0; j < S; ++j) acc[i][j] = 0.f;
#pragma unroll
      for (int pi = 0; pi < P; ++pi)
#pragma unroll
        for (int pj = 0; pj < P; ++pj) acc[sub(pi)][sub(pj)] = fmaf(dv[pi][pj], a[off(pi) + 1][off(pj) + 1], acc[sub(pi)][sub(pj)]);
#pragma unroll
      for (int i = 0; i < S; ++i)
        if (S * y + i < Q.H) __stcs(reinterpret_cast<float4*>(o + (long long)i * Q.W), make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    }
  }
}

template <int P>
static int launch_bwd_strided(SBwdParams Q, cudaStream_t st) {
  // enough channel chunks to fill the GPU a few times over; at least 8 channels per thread to amortise its P*P gradient loads
  const long long cells = (long long)cdiv(Q.ow, 32) * cdiv(Q.oh, 4) * Q.BG;
  int nchunk = (int)std::min<long long>(std::max<long long>(1, (148 * 16) / std::max<long long>(cells, 1)), std::max(1, Q.C / 8));
  nchunk = std::max(1, std::min(nchunk, Q.C));
  Q.cpc = cdiv(Q.C, nchunk);
  Q.nchunk = cdiv(Q.C, Q.cpc);
  const dim3 grid(cdiv(Q.ow, 32), cdiv(Q.oh, 4), Q.BG * Q.nchunk);
  if (Q.din1) { local_corr_bwd_strided_kernel<P, false><<<grid, 128, 0, st>>>(Q); EZB_LAUNCH_OK(); }
  if (Q.din2) { local_corr_bwd_strided_kernel<P, true><<<grid, 128, 0, st>>>(Q); EZB_LAUNCH_OK(); }
  return EZB_OK;
}

// ------------------------------------------------------------------------------------ adjoints (K4b), same regular case
//   SECOND = false: din1[c,y,x] = scale * sum_{pi,pj} dout[pi,pj,y,x] * in2[c, y + pi - md, x + pj - md]
//   SECOND = true : din2[c,y,x] = scale * sum_{pi,pj} dout[P-1-pi, P-1-pj, y + pi - md, x + pj - md] * in1[c, y + pi - md, x + pj - md]
// (autograd of sampler.py:121-127; the second form is the scatter of din2 rewritten as a gather with the flipped displacement).
// Both are "12-value window of the other operand x 9 displacement columns" products like the forward, with the contraction
// over displacements instead of channels:
//   block  : 32 x 8 gradient pixels x 16 channels, 128 threads; thread = 4 consecutive pixels x 8 channels (32 accumulators).
//   staged : the other operand's 16-channel halo tile (16 x 40 per channel) once per block; per displacement row the P planes
//            of dout for that row (8 x 40 with halo, double-buffered, one barrier per displacement row).
//   per displacement row a thread loads its P dout quads (aligned LDS.128 for din1; for din2 the quad sits pj - md pixels off
//   the thread's own, four scalar loads) and per channel the 12-value window (three LDS.128) -> 36 FMAs.
// Vector path only (W % 4 == 0, 16-byte aligned tensors); everything else stays on local_corr_bwd_tiled_kernel.
constexpr int BCB = 16, BCT = 8, BTH = 8;  // channels per block / per thread, tile rows
constexpr int BNT = (TW / 4) * BTH * (BCB / BCT);

struct BwdParams {
  const float* other;  // in2 (din1) / in1 (din2)
  const float* dout;
  float* din;
  int C, H, W;
  float scale;
  long long dout_bstride;
};

template <int P>
struct BwdCfg {
  static constexpr int MD = (P - 1) / 2, HALO_W = MDA + TW + 4, HALO_H = BTH + 2 * MD;
  static constexpr int OTHER = BCB * HALO_H * HALO_W;  // floats
  static constexpr int DROW = P * BTH * HALO_W;        // floats per displacement row
  static constexpr size_t SMEM = (size_t)(OTHER + 2 * DROW) * sizeof(float);
};

template <int P, bool SECOND>
__global__ void __launch_bounds__(BNT, 3) local_corr_bwd_rt_kernel(const BwdParams Q) {
  using K = BwdCfg<P>;
  constexpr int MD = K::MD, HALO_W = K::HALO_W, HALO_H = K::HALO_H, Q4 = HALO_W / 4;
  extern __shared__ __align__(16) float sm_brt[];
  float* so = sm_brt;              // [BCB][HALO_H][HALO_W]
  float* sd = sm_brt + K::OTHER;   // [2][P][BTH][HALO_W]
  const uint32_t so_a = (uint32_t)__cvta_generic_to_shared(so), sd_a = (uint32_t)__cvta_generic_to_shared(sd);
  const int tid = threadIdx.x;
  const int xtiles = (Q.W + TW - 1) / TW;
  const int x0 = (blockIdx.x % xtiles) * TW, y0 = (blockIdx.x / xtiles) * BTH;
  const int c0 = blockIdx.y * BCB, b = blockIdx.z;
  const long long HW = (long long)Q.H * Q.W;
  const float* other = Q.other + ((long long)b * Q.C + c0) * HW;
  const float* dout = Q.dout + (long long)b * Q.dout_bstride;

  // ---- the other operand's halo tile, all 16 channels (zero outside the image / beyond C)
  for (int i = tid; i < HALO_H * Q4; i += BNT) {
    const int hy = i / Q4, hq = i - hy * Q4;
    const int y = y0 - MD + hy, x = x0 - MDA + 4 * hq;
    const bool ok = y >= 0 && y < Q.H && x >= 0 && x < Q.W;
    const float* src = other + (long long)y * Q.W + x;
#pragma unroll
    for (int c = 0; c < BCB; ++c) {
      if (ok && c0 + c < Q.C)
        cp_async_16(so_a + (uint32_t)((c * HALO_H + hy) * HALO_W + 4 * hq) * 4u, src + (long long)c * HW, 16u);
      else
        *reinterpret_cast<float4*>(so + (c * HALO_H + hy) * HALO_W + 4 * hq) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  // ---- dout planes of one displacement row: sd[buf][pj][r][t] = G[pi][pj][y0 + r + dy][x0 - 4 + t]
  //      din1: G = dout, dy = 0 (only t in [4, 36) is read);  din2: G = dout with both displacements flipped, dy = pi - md
  auto stage_d = [&](int buf, int pi) {
    const int dy = SECOND ? pi - MD : 0;
    const float* plane0 = dout + (long long)(SECOND ? (P - 1 - pi) * P + (P - 1) : pi * P) * HW;
    constexpr int QB = SECOND ? 0 : 1, QN = SECOND ? Q4 : Q4 - 2;  // quads of a row that are read
    for (int i = tid; i < P * BTH * QN; i += BNT) {
      const int pj = i / (BTH * QN), rem = i - pj * (BTH * QN), r = rem / QN, hq = QB + rem - r * QN;
      const int y = y0 + r + dy, x = x0 - MDA + 4 * hq;
      const int off = ((buf * P + pj) * BTH + r) * HALO_W + 4 * hq;
      if (y >= 0 && y < Q.H && x >= 0 && x < Q.W)
        cp_async_16(sd_a + (uint32_t)off * 4u, plane0 + (SECOND ? -(long long)pj : (long long)pj) * HW + (long long)y * Q.W + x, 16u);
      else
        *reinterpret_cast<float4*>(sd + off) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage_d(0, 0);  // (the halo copies above belong to this first group)

  const int cg = tid / ((TW / 4) * BTH), r64 = tid - cg * ((TW / 4) * BTH), qx = r64 & 7, ty = r64 >> 3;
  float acc[BCT][4];
#pragma unroll
  for (int k = 0; k < BCT; ++k)
#pragma unroll
    for (int px = 0; px < 4; ++px) acc[k][px] = 0.f;

  for (int pi = 0; pi < P; ++pi) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // row pi has landed; every thread is done with the buffer row pi + 1 goes into
    if (pi + 1 < P) stage_d((pi + 1) & 1, pi + 1);
    const float* dr = sd + (((pi & 1) * P) * BTH + ty) * HALO_W + 4 * qx;
    float d[P][4];
#pragma unroll
    for (int pj = 0; pj < P; ++pj) {
      const float* dp = dr + pj * (BTH * HALO_W);
      if constexpr (!SECOND) {
        const float4 v = *reinterpret_cast<const float4*>(dp + MDA);
        d[pj][0] = v.x; d[pj][1] = v.y; d[pj][2] = v.z; d[pj][3] = v.w;
      } else {
#pragma unroll
        for (int px = 0; px < 4; ++px) d[pj][px] = dp[MDA - MD + pj + px];
      }
    }
    const float4* wrow = reinterpret_cast<const float4*>(so + ((cg * BCT) * HALO_H + ty + pi) * HALO_W + 4 * qx);
#pragma unroll
    for (int k = 0; k < BCT; ++k) {
      const float4 v0 = wrow[k * (HALO_H * Q4)], v1 = wrow[k * (HALO_H * Q4) + 1], v2 = wrow[k * (HALO_H * Q4) + 2];
      const float w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
      for (int pj = 0; pj < P; ++pj)
#pragma unroll
        for (int px = 0; px < 4; ++px) acc[k][px] = fmaf(d[pj][px], w[MDA - MD + pj + px], acc[k][px]);
    }
  }

  const int y = y0 + ty, x = x0 + 4 * qx;
  if (y >= Q.H || x >= Q.W) return;
  float* dst = Q.din + ((long long)b * Q.C + c0 + cg * BCT) * HW + (long long)y * Q.W + x;
#pragma unroll
  for (int k = 0; k < BCT; ++k)
    if (c0 + cg * BCT + k < Q.C)
      *reinterpret_cast<float4*>(dst + (long long)k * HW) = make_float4(acc[k][0] * Q.scale, acc[k][1] * Q.scale, acc[k][2] * Q.scale, acc[k][3] * Q.scale);
}

template <int P>
static int launch_bwd(const float* in1, const float* in2, const float* dout, int B, int C, int H, int W, float scale, long long dout_bstride,
                      float* din1, float* din2, cudaStream_t st) {
  using K = BwdCfg<P>;
  BwdParams Q;
  Q.dout = dout; Q.C = C; Q.H = H; Q.W = W; Q.scale = scale; Q.dout_bstride = dout_bstride;
  const dim3 grid((unsigned)(cdiv(W, TW) * cdiv(H, BTH)), cdiv(C, BCB), B);
  if (din1) {
    Q.other = in2; Q.din = din1;
    EZB_CUDA_OK(cudaFuncSetAttribute(local_corr_bwd_rt_kernel<P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM));
    local_corr_bwd_rt_kernel<P, false><<<grid, BNT, K::SMEM, st>>>(Q);
    EZB_LAUNCH_OK();
  }
  if (din2) {
    Q.other = in1; Q.din = din2;
    EZB_CUDA_OK(cudaFuncSetAttribute(local_corr_bwd_rt_kernel<P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM));
    local_corr_bwd_rt_kernel<P, true><<<grid, BNT, K::SMEM, st>>>(Q);
    EZB_LAUNCH_OK();
  }
  return EZB_OK;
}

}  // namespace k4rt

bool local_corr_rt_supported(int H, int W, int BG, int Ph, int Pw, int sh, int sw, int padh, int padw, int dph, int dpw) {
  return Ph == Pw && (Ph == 3 || Ph == 5 || Ph == 7 || Ph == 9) && sh == 1 && sw == 1 && padh == 0 && padw == 0 && dph == 1 && dpw == 1 &&
         cdiv(H, 4) <= 65535 && BG <= 65535;
}

static long long rt_tiles(int BG, int H, int W) { return (long long)cdiv(W, k4rt::TW) * cdiv(H, 8) * BG; }  // 32 x 8 tiles

// Small maps (fewer 32 x 8 tiles than SMs) run the <4, 1> configuration.
static bool rt_small(int BG, int H, int W) {
  if (const char* e = getenv("EZB_LOCAL_CORR_RT_SMALL")) return atoi(e) != 0;
  return rt_tiles(BG, H, W) < 148;
}

// Preferred over the tensor-core path?  The register-tiled kernel's only parallelism is over pixels, and its cost grows with
// 81 FMAs per channel and pixel where the band GEMM's does not: it wins on PWC-Net's finest levels (few channels, large maps),
// not on the coarse ones (measured per level: profiles/r02_local_corr_rt.md).
bool local_corr_rt_preferred(int BG, int C, int H, int W, bool tc_available) {
  if (const char* e = getenv("EZB_LOCAL_CORR_RT")) return atoi(e) != 0;
  const long long tiles = rt_tiles(BG, H, W);
  if (!tc_available) return tiles >= 4;  // against the generic fp32 kernel (which splits displacement rows over blocks for tiny maps)
  return tiles >= 200 && C <= 96;
}

int local_corr_rt_launch(const float* in1, const float* in2, int BG, int G, int C, int H, int W, int P, float scale,
                         float slope, float* out, long long out_bstride, cudaStream_t st) {
  k4rt::Params Q;
  Q.in1 = in1; Q.in2 = in2; Q.out = out;
  Q.BG = BG; Q.G = G; Q.C = C; Q.H = H; Q.W = W;
  Q.scale = scale; Q.slope = slope; Q.out_bstride = out_bstride;
  const bool vec = W % 4 == 0 && ((reinterpret_cast<uintptr_t>(in1) | reinterpret_cast<uintptr_t>(in2) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
                   out_bstride % 4 == 0;
  const bool small = rt_small(BG, H, W);
#define EZB_RT_CASE(PP)                                                                                                   \
  case PP:                                                                                                                \
    if (small) return vec ? k4rt::launch<PP, 4, 1, true>(Q, st) : k4rt::launch<PP, 4, 1, false>(Q, st);                   \
    return vec ? k4rt::launch<PP, 8, 3, true>(Q, st) : k4rt::launch<PP, 8, 3, false>(Q, st);
  switch (P) {
    EZB_RT_CASE(9)
    EZB_RT_CASE(7)
    EZB_RT_CASE(5)
    EZB_RT_CASE(3)
  }
#undef EZB_RT_CASE
  return fail(EZB_ERR_UNSUPPORTED, "register-tiled local correlation: patch size %d", P);
}

bool local_corr_bwd_rt_supported(const float* in1, const float* in2, const float* dout, const float* din1, const float* din2, int B, int C,
                                 int H, int W, int Ph, int Pw, int sh, int sw, int padh, int padw, int dph, int dpw, long long dout_bstride) {
  if (const char* e = getenv("EZB_LOCAL_CORR_BWD_RT"))
    if (atoi(e) == 0) return false;
  const uintptr_t bits = reinterpret_cast<uintptr_t>(in1) | reinterpret_cast<uintptr_t>(in2) | reinterpret_cast<uintptr_t>(dout) |
                         reinterpret_cast<uintptr_t>(din1) | reinterpret_cast<uintptr_t>(din2);
  return Ph == Pw && (Ph == 3 || Ph == 5 || Ph == 7 || Ph == 9) && sh == 1 && sw == 1 && padh == 0 && padw == 0 && dph == 1 && dpw == 1 &&
         W % 4 == 0 && (bits & 15) == 0 && dout_bstride % 4 == 0 && B <= 65535 && cdiv(C, k4rt::BCB) <= 65535;
}

int local_corr_bwd_rt_launch(const float* in1, const float* in2, const float* dout, int B, int C, int H, int W, int P, float scale,
                             long long dout_bstride, float* din1, float* din2, cudaStream_t st) {
  switch (P) {
    case 9: return k4rt::launch_bwd<9>(in1, in2, dout, B, C, H, W, scale, dout_bstride, din1, din2, st);
    case 7: return k4rt::launch_bwd<7>(in1, in2, dout, B, C, H, W, scale, dout_bstride, din1, din2, st);
    case 5: return k4rt::launch_bwd<5>(in1, in2, dout, B, C, H, W, scale, dout_bstride, din1, din2, st);
    case 3: return k4rt::launch_bwd<3>(in1, in2, dout, B, C, H, W, scale, dout_bstride, din1, din2, st);
  }
  return fail(EZB_ERR_UNSUPPORTED, "register-tiled local correlation adjoint: patch size %d", P);
}

// Strided forward: one dilation of 1, stride 2 or 4 in both directions, no padding, square window 9 (DCVNet) .. 3.
bool local_corr_rt_strided_supported(const float* in1, const float* in2, const float* out, int BG, int C, int H, int W, int Ph, int Pw,
                                     int sh, int sw, int padh, int padw, int dph, int dpw, long long out_bstride) {
  if (const char* e = getenv("EZB_LOCAL_CORR_RT"))
    if (atoi(e) == 0) return false;
  const int ow = cdiv(W, sw);
  const uintptr_t bits = reinterpret_cast<uintptr_t>(in1) | reinterpret_cast<uintptr_t>(in2) | reinterpret_cast<uintptr_t>(out);
  return Ph == Pw && (Ph == 9 || Ph == 7 || Ph == 5 || Ph == 3) && sh == sw && (sh == 2 || sh == 4) && padh == 0 && padw == 0 && dph == 1 && dpw == 1 &&
         W % 4 == 0 && ow % 4 == 0 && (bits & 15) == 0 && out_bstride % 4 == 0 && BG <= 65535 && cdiv(cdiv(H, sh), k4rt::SQH) <= 65535 &&
         (sh == 4 || C <= 128);
}

int local_corr_rt_strided_launch(const float* in1, const float* in2, int BG, int G, int C, int H, int W, int P, int S, float scale, float slope,
                                 float* out, long long out_bstride, cudaStream_t st) {
  k4rt::SParams Q;
  Q.in1 = in1; Q.in2 = in2; Q.out = out;
  Q.BG = BG; Q.G = G; Q.C = C; Q.H = H; Q.W = W; Q.oh = cdiv(H, S); Q.ow = cdiv(W, S);
  Q.scale = scale; Q.slope = slope; Q.out_bstride = out_bstride;
#define EZB_RTS_CASE(PP)                                                                              \
  case PP: return S == 4 ? k4rt::launch_strided<PP, 4>(Q, st) : k4rt::launch_strided<PP, 2>(Q, st);
  switch (P) {
    EZB_RTS_CASE(9)
    EZB_RTS_CASE(7)
    EZB_RTS_CASE(5)
    EZB_RTS_CASE(3)
  }
#undef EZB_RTS_CASE
  return fail(EZB_ERR_UNSUPPORTED, "strided register-tiled local correlation: patch size %d", P);
}

// Strided adjoints: stride 4 in both directions, one dilation of 1, no padding, square odd window <= 9, W % 4 == 0.
bool local_corr_bwd_strided_supported(const float* in1, const float* in2, const float* dout, const float* din1, const float* din2, int BG,
                                      int C, int H, int W, int Ph, int Pw, int sh, int sw, int padh, int padw, int dph, int dpw) {
  if (const char* e = getenv("EZB_LOCAL_CORR_BWD_RT"))
    if (atoi(e) == 0) return false;
  const uintptr_t bits = reinterpret_cast<uintptr_t>(in1) | reinterpret_cast<uintptr_t>(in2) | reinterpret_cast<uintptr_t>(din1) |
                         reinterpret_cast<uintptr_t>(din2);
  (void)dout;  // read with scalar loads: no alignment requirement
  return Ph == Pw && (Ph == 9 || Ph == 7 || Ph == 5 || Ph == 3) && sh == 4 && sw == 4 && padh == 0 && padw == 0 && dph == 1 && dpw == 1 &&
         W % 4 == 0 && (bits & 15) == 0 && cdiv(cdiv(H, 4), 4) <= 65535 && (long long)BG * C <= 65535;
}

int local_corr_bwd_strided_launch(const float* in1, const float* in2, const float* dout, int BG, int G, int C, int H, int W, int P, float scale,
                                  long long dout_bstride, float* din1, float* din2, cudaStream_t st) {
  k4rt::SBwdParams Q;
  Q.in1 = in1; Q.in2 = in2; Q.dout = dout; Q.din1 = din1; Q.din2 = din2;
  Q.BG = BG; Q.G = G; Q.C = C; Q.H = H; Q.W = W; Q.oh = cdiv(H, 4); Q.ow = W / 4;
  Q.nchunk = 1; Q.cpc = C; Q.scale = scale; Q.dout_bstride = dout_bstride;
  switch (P) {
    case 9: return k4rt::launch_bwd_strided<9>(Q, st);
    case 7: return k4rt::launch_bwd_strided<7>(Q, st);
    case 5: return k4rt::launch_bwd_strided<5>(Q, st);
    case 3: return k4rt::launch_bwd_strided<3>(Q, st);
  }
  return fail(EZB_ERR_UNSUPPORTED, "strided local correlation adjoint: patch size %d", P);
}

}  // namespace ezb
