// K1b on tensor cores: the two GEMMs of the all-pairs correlation adjoint, tcgen05 kind::tf32.
//
// Autograd of MutliScalePairwise4DCorr.corr (reference ezflow/similarity/correlation/pairwise.py:78-88):
//   df1[b,c,q] = sum_t g[b,q,t] f2[b,c,t] / sqrt(C)        df2[b,c,t] = sum_q g[b,q,t] f1[b,c,q] / sqrt(C)
// g = level-0 gradient volume (fp32, 4.2 GB per 1080p pair) after the pooled levels have been folded into it.
// Both GEMMs read g, f1 and f2 as they are in HBM — fp32, through TMA, consumed as tf32 — so there is no conversion
// or transposition pass over the 4.2 GB volume:
//   df1:  D[q, c]  M = 256 queries (two 128-row accumulators), N = 256 channels, K = t
//         A = g rows (K-major: t contiguous), B = f2 (K-major: (C, N) rows are channels)
//   df2:  D[c, t]  M = 256 channels (two accumulators), N = 256 targets, K = q
//         A = f1 (K-major), B = g read as an MN-major operand (t contiguous, q strided; eight 32x32 TMA boxes per
//         stage written with the 32-byte-atom 128B swizzle = the SWIZZLE_128B_BASE32B layout tf32 MN-major needs)
// One CTA owns a 256x256 output tile and runs the whole K loop (N/32 k-blocks of 64 KB through a 3-stage TMA ring);
// the two accumulators fill TMEM (512 columns).  HBM-bound: g is read exactly once per GEMM.
#include "ezb_common.cuh"
#include "ezb_ptx.cuh"
#include "ezb_tma.cuh"

namespace ezb {
namespace k1btc {

constexpr int BM = 128, BN = 256, BK = 32;  // BK fp32 = one 128-byte swizzle row
constexpr int STAGES = 3;
constexpr int A_BYTES = BM * BK * 4;         // 16 KB per 128-row half
constexpr int B_BYTES = BN * BK * 4;         // 32 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + B_BYTES;
constexpr int NUM_THREADS = 256;             // warp 0 TMA, 1 MMA, 2 TMEM alloc, 4-7 epilogue
constexpr int SMEM_BAR = STAGES * STAGE_BYTES;
constexpr int SMEM_BYTES = SMEM_BAR + 256 + 1024;
constexpr int TMEM_COLS = 512;

struct Params {
  int mode;  // 0: df1 (rows = queries, cols = channels)   1: df2 (rows = channels, cols = targets)
  int C, N;
  float alpha;
  float* out;  // (B, C, N)
};

__device__ __forceinline__ uint64_t desc_mnmajor_sw128_base32(uint32_t smem_addr, uint32_t lbo_bytes) {
  // MN-major 32-bit operands only exist in the SWIZZLE_128B_BASE32B layout (cute::UMMA::Layout_MN_SW128_32B_Atom): an
  // atom is 4 K-rows x 128 bytes of MN whose 32-byte granules are XOR-swizzled with the row index — what TMA writes
  // with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  LBO = distance between atoms along MN, SBO = between groups of 4 K-rows.
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version 1
  d |= static_cast<uint64_t>(1) << 61;  // SWIZZLE_128B_BASE32B
  return d;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
    corr_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmF, const Params P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  const uint32_t sBar = base + SMEM_BAR;
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (STAGES + s); };
  const uint32_t done_bar = sBar + 8u * (2 * STAGES);
  volatile uint32_t* tmem_ptr_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + SMEM_BAR + 8 * (2 * STAGES + 1));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int tile0 = blockIdx.x * 256;  // first query (mode 0) / first target (mode 1) of this tile
  const int nk = (P.N + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmG);
    ptx::prefetch_tmap(&tmF);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(done_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_ptr_slot)), TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_slot;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    uint32_t stage = 0, phase = 0;
    for (int k = 0; k < nk; ++k) {
      ptx::mbar_wait(empty_bar(stage), phase ^ 1);
      const uint32_t sA = base + stage * STAGE_BYTES, sB = sA + 2 * A_BYTES;
      ptx::mbar_expect_tx(full_bar(stage), STAGE_BYTES);
      if (P.mode == 0) {
        // A = g[b, tile0 + (0..255), k*32 ..]: two boxes of 128 query rows; B = f2[b, 0..255, k*32 ..]
        ptx::tma_load_2d_hint(sA, &tmG, full_bar(stage), k * BK, b * P.N + tile0, ptx::kEvictFirst);
        ptx::tma_load_2d_hint(sA + A_BYTES, &tmG, full_bar(stage), k * BK, b * P.N + tile0 + BM, ptx::kEvictFirst);
        ptx::tma_load_3d_hint(sB, &tmF, full_bar(stage), k * BK, 0, b, ptx::kEvictLast);
      } else {
        // A = f1[b, 0..255, k*32 ..]: two boxes of 128 channel rows; B = g[b, k*32 + (0..31), tile0 + (0..255)] as eight
        // MN-major atoms columns (32 targets x 32 query rows each)
        ptx::tma_load_3d_hint(sA, &tmF, full_bar(stage), k * BK, 0, b, ptx::kEvictLast);
        ptx::tma_load_3d_hint(sA + A_BYTES, &tmF, full_bar(stage), k * BK, BM, b, ptx::kEvictLast);
#pragma unroll
        for (int j = 0; j < BN / 32; ++j)
          ptx::tma_load_2d_hint(sB + j * (BK * 128), &tmG, full_bar(stage), tile0 + 32 * j, b * P.N + k * BK, ptx::kEvictFirst);
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = ptx::umma_idesc(/*tf32*/ 2, BM, BN) | (P.mode == 1 ? (1u << 16) : 0u);  // bit 16: B is MN-major
    uint32_t stage = 0, phase = 0;
    for (int k = 0; k < nk; ++k) {
      ptx::mbar_wait(full_bar(stage), phase);
      ptx::tc_fence_after();
      const uint32_t sA = base + stage * STAGE_BYTES, sB = sA + 2 * A_BYTES;
      const uint64_t b_desc = P.mode == 0 ? ptx::umma_desc_kmajor_sw128(sB) : desc_mnmajor_sw128_base32(sB, BK * 128);
      const uint32_t b_step = P.mode == 0 ? 2u : 64u;  // 8 tf32 of K: 32 bytes along a K-major row / one 1024-byte atom
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint64_t a_desc = ptx::umma_desc_kmajor_sw128(sA + h * A_BYTES);
#pragma unroll
        for (int kk = 0; kk < BK / 8; ++kk)
          ptx::umma_tf32(tmem_base + h * BN, a_desc + 2 * kk, b_desc + b_step * kk, idesc, (k | kk) != 0);
      }
      ptx::umma_commit(empty_bar(stage));
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    ptx::umma_commit(done_bar);
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int quad = warp & 3;
    ptx::mbar_wait(done_bar, 0);
    ptx::tc_fence_after();
    float* out_b = P.out + (long long)b * P.C * P.N;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const int row = h * BM + quad * 32 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + h * BN + c0 + ((uint32_t)(quad * 32) << 16), r);
        ptx::tmem_ld_wait();
        if (P.mode == 0) {
          // row = query, columns = channels: for a fixed channel the warp writes 32 consecutive queries
          const int q = tile0 + row;
          if (q < P.N) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < P.C) out_b[(long long)(c0 + j) * P.N + q] = __uint_as_float(r[j]) * P.alpha;
          }
        } else {
          // row = channel, columns = targets: each lane writes 32 consecutive targets of its channel row
          const int t0 = tile0 + c0;
          if (row < P.C && t0 < P.N) {
            float* dst = out_b + (long long)row * P.N + t0;
            if (t0 + 32 <= P.N && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]) * P.alpha, __uint_as_float(r[j + 1]) * P.alpha,
                                                                  __uint_as_float(r[j + 2]) * P.alpha, __uint_as_float(r[j + 3]) * P.alpha);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (t0 + j < P.N) dst[j] = __uint_as_float(r[j]) * P.alpha;
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace k1btc

// true when the tcgen05 path covers this problem (TMA needs 16-byte-aligned row strides)
bool corr_bwd_tc_supported(int C, int N, long long g_qstride) { return C <= 256 && (N % 4) == 0 && (g_qstride % 4) == 0; }

// g: level-0 gradient volume, row (b*N + q) at g + row*g_qstride, t contiguous.  F: the other feature map (B,C,N).
// mode 0: out = dfmap1 from F = fmap2;  mode 1: out = dfmap2 from F = fmap1.
int corr_bwd_tc_launch(int mode, const float* g, long long g_qstride, const float* F, int B, int C, int N, float alpha, float* out,
                       cudaStream_t st) {
  using namespace k1btc;
  EZB_REQUIRE((long long)B * N < (1ll << 31), "batch too large for one launch (split it)");
  CUtensorMap tmG, tmF;
  {
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)B * N};
    cuuint64_t str[1] = {(cuuint64_t)g_qstride * 4};
    cuuint32_t box[2] = {BK, (cuuint32_t)(mode == 0 ? BM : BK)};
    if (int r = make_tmap(&tmG, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(g), dims, str, box,
                          mode == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
      return r;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)C, (cuuint64_t)B};
    cuuint64_t str[2] = {(cuuint64_t)N * 4, (cuuint64_t)C * N * 4};
    cuuint32_t box[3] = {BK, (cuuint32_t)(mode == 0 ? BN : BM), 1};
    if (int r = make_tmap(&tmF, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(F), dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return r;
  }
  Params P;
  P.mode = mode; P.C = C; P.N = N; P.alpha = alpha; P.out = out;
  EZB_CUDA_OK(cudaFuncSetAttribute(corr_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  dim3 grid(cdiv(N, 256), B);
  EZB_REQUIRE(grid.y <= 65535, "batch too large for one launch (split it)");
  corr_bwd_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(tmG, tmF, P);
  EZB_LAUNCH_OK();
  return EZB_OK;
}

}  // namespace ezb
