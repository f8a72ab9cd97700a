// K1b on tensor cores: the two GEMMs of the all-pairs correlation adjoint, tcgen05 kind::tf32.
//
// Autograd of MutliScalePairwise4DCorr.corr (reference ezflow/similarity/correlation/pairwise.py:78-88):
//   df1[b,c,q] = sum_t g[b,q,t] f2[b,c,t] / sqrt(C)        df2[b,c,t] = sum_q g[b,q,t] f1[b,c,q] / sqrt(C)
// (K2b is folded in by linearity: level l of the pyramid is f1^T pool_l(f2) / sqrt(C), so with g_l the gradient of level l
//   df1 = sum_l g_l pool_l(f2)^T / sqrt(C),      df2 = sum_l unpool_l( g_l^T f1 ) / sqrt(C)
//  i.e. the df1 GEMM simply runs its K loop over the targets of ALL levels against the pooled f2, and the df2 GEMM
//  produces one small (C x N_l) matrix per level that a tiny kernel un-pools and adds; the 5.6 GB gradient pyramid is
//  never folded level by level.)
// Both GEMMs read g, f1 and f2 as they are in HBM — fp32, through TMA, consumed as tf32 — so there is no conversion
// or transposition pass over the gradient volume (4.2 GB for level 0 of a 1080p pair):
//   df1:  D[q, c]  M = 256 queries (two 128-row accumulators), N = 256 channels, K = t
//         A = g rows (K-major: t contiguous), B = f2 (K-major: (C, N) rows are channels)
//   df2:  D[c, t]  M = 256 channels (two accumulators), N = 256 targets, K = q
//         A = f1 (K-major), B = g read as an MN-major operand (t contiguous, q strided; eight 32x32 TMA boxes per
//         stage written with the 32-byte-atom 128B swizzle = the SWIZZLE_128B_BASE32B layout tf32 MN-major needs)
// One CTA owns a 256x256 output tile and runs the whole K loop (N/32 k-blocks of 64 KB through a 3-stage TMA ring);
// the two accumulators fill TMEM (512 columns).  HBM-bound: g is read exactly once per GEMM.
#include "ezb_common.cuh"

#include <algorithm>
#include <cstdlib>
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

constexpr int MAXL = 4;
struct Maps {
  CUtensorMap g[MAXL];  // gradient level l: 2-D (t_l, B*N rows); K-major boxes (mode 0) / MN-major boxes (mode 1)
  CUtensorMap f[MAXL];  // mode 0: level l of (pooled) fmap2 as (N_l, C, B);  mode 1: f[0] = fmap1 (N, C, B)
};
struct Params {
  int mode;  // 0: df1 (rows = queries, cols = channels)   1: df2 (rows = channels, cols = targets of one level)
  int C, N, L;
  int Nl[MAXL];             // targets per level
  int tile_base[MAXL + 1];  // mode 1: first grid tile of level l (a level has cdiv(Nl, 256) * split[l] of them)
  int split[MAXL];          // mode 1: CTAs sharing one target tile of level l, each with a slice of the query (K) range; their
                            // partial sums go to out[l] + slice * B * C * Nl[l] and are added by reduce_slices_kernel.  The coarse
                            // levels have few target tiles with (under ranges) long K loops: without the split they are the tail.
  float alpha;
  float* out[MAXL];         // mode 0: out[0] = dfmap1 (B,C,N);  mode 1: out[l] = (B, C, Nl[l]) (out[0] = dfmap2)
  // Optional support of the gradient pyramid (ezb_corr_grad_ranges): for every (pair, 256-query tile T, level l) the
  // half-open range [t0, t1) of 256-target tiles outside which the rows of T are exact zeros.  mode 0 restricts its K
  // loop to it, mode 1 skips the query blocks whose range misses its target tile.  NULL: dense.
  const int2* ranges;       // [B][nT][MAXL]
  int nT;
};
__device__ __forceinline__ int2 range_of(const Params& P, int b, int T, int l) {
  return __ldg(P.ranges + ((long long)b * P.nT + T) * MAXL + l);
}

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
    corr_bwd_tc_kernel(const __grid_constant__ Maps M, const Params P) {
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
  int lvl = 0;  // mode 1: the level this tile belongs to
  if (P.mode == 1)
    while (lvl + 1 < P.L && (int)blockIdx.x >= P.tile_base[lvl + 1]) ++lvl;
  const int lt = P.mode == 0 ? (int)blockIdx.x : (int)blockIdx.x - P.tile_base[lvl];     // tile (x split slice) inside the level
  const int nsplit = P.mode == 0 ? 1 : P.split[lvl], slice = lt % nsplit;
  const int tile0 = (lt / nsplit) * 256;  // first query / first target
  // K loop: mode 0 walks the targets of every level, mode 1 the queries
  const int l_begin = P.mode == 0 ? 0 : lvl, l_end = P.mode == 0 ? P.L : lvl + 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&M.g[lvl]);
    ptx::prefetch_tmap(&M.f[0]);
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

  // k-blocks (BK = 32 targets / queries) this CTA's K loop visits at level l: [k0, k1), and in mode 1 only the blocks whose
  // 256-query tile has this CTA's target tile in its range
  auto k_bounds = [&](int l, int& k0, int& k1) {
    const int nk = ((P.mode == 0 ? P.Nl[l] : P.N) + BK - 1) / BK;
    k0 = 0;
    k1 = nk;
    if (P.ranges != nullptr && P.mode == 0) {
      const int2 r = range_of(P, b, blockIdx.x, l);
      k0 = min(nk, r.x * (256 / BK));
      k1 = max(k0, min(nk, r.y * (256 / BK)));
    }
    if (P.mode == 1 && nsplit > 1) {  // this slice's 256-query tiles
      k0 = min(nk, (int)((long long)P.nT * slice / nsplit) * (256 / BK));
      k1 = min(nk, (int)((long long)P.nT * (slice + 1) / nsplit) * (256 / BK));
    }
  };
  const int my_tt = tile0 >> 8;  // mode 1: this CTA's 256-target tile of level lvl
  auto k_live = [&](int l, int k) {
    if (P.ranges == nullptr || P.mode == 0) return true;
    const int2 r = range_of(P, b, (k * BK) >> 8, l);
    return my_tt >= r.x && my_tt < r.y;
  };
  // does the K loop visit anything at all?  (an all-zero gradient support leaves the accumulators unwritten)
  __shared__ int s_any;
  if (threadIdx.x == 0) {
    int any = 0;
    for (int l = l_begin; l < l_end && !any; ++l) {
      int k0, k1;
      k_bounds(l, k0, k1);
      if (P.ranges == nullptr || P.mode == 0) {
        any = k1 > k0;
      } else {
        for (int k = k0; k < k1 && !any; k += 256 / BK) any = k_live(l, k);
      }
    }
    s_any = any;
  }
  __syncthreads();
  const bool any_k = s_any != 0;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    uint32_t stage = 0, phase = 0;
    for (int l = l_begin; l < l_end; ++l) {
      int k0, k1;
      k_bounds(l, k0, k1);
      for (int k = k0; k < k1; ++k) {
        if (!k_live(l, k)) { k |= 256 / BK - 1; continue; }  // skip the whole 256-query tile
        ptx::mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t sA = base + stage * STAGE_BYTES, sB = sA + 2 * A_BYTES;
        ptx::mbar_expect_tx(full_bar(stage), STAGE_BYTES);
        if (P.mode == 0) {
          // A = g_l[b, tile0 + (0..255), k*32 ..]: two boxes of 128 query rows; B = pool_l(f2)[b, 0..255, k*32 ..]
          ptx::tma_load_2d_hint(sA, &M.g[l], full_bar(stage), k * BK, b * P.N + tile0, ptx::kEvictFirst);
          ptx::tma_load_2d_hint(sA + A_BYTES, &M.g[l], full_bar(stage), k * BK, b * P.N + tile0 + BM, ptx::kEvictFirst);
          ptx::tma_load_3d_hint(sB, &M.f[l], full_bar(stage), k * BK, 0, b, ptx::kEvictLast);
        } else {
          // A = f1[b, 0..255, k*32 ..]: two boxes of 128 channel rows; B = g_l[b, k*32 + (0..31), tile0 + (0..255)] as
          // eight MN-major atom columns (32 targets x 32 query rows each)
          ptx::tma_load_3d_hint(sA, &M.f[0], full_bar(stage), k * BK, 0, b, ptx::kEvictLast);
          ptx::tma_load_3d_hint(sA + A_BYTES, &M.f[0], full_bar(stage), k * BK, BM, b, ptx::kEvictLast);
#pragma unroll
          for (int j = 0; j < BN / 32; ++j)
            ptx::tma_load_2d_hint(sB + j * (BK * 128), &M.g[l], full_bar(stage), tile0 + 32 * j, b * P.N + k * BK, ptx::kEvictFirst);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = ptx::umma_idesc(/*tf32*/ 2, BM, BN) | (P.mode == 1 ? (1u << 16) : 0u);  // bit 16: B is MN-major
    uint32_t stage = 0, phase = 0, first = 1;
    for (int l = l_begin; l < l_end; ++l) {
      int k0, k1;
      k_bounds(l, k0, k1);
      for (int k = k0; k < k1; ++k) {
        if (!k_live(l, k)) { k |= 256 / BK - 1; continue; }
        ptx::mbar_wait(full_bar(stage), phase);
        ptx::tc_fence_after();
        const uint32_t sA = base + stage * STAGE_BYTES, sB = sA + 2 * A_BYTES;
        const uint64_t b_desc = P.mode == 0 ? ptx::umma_desc_kmajor_sw128(sB) : desc_mnmajor_sw128_base32(sB, BK * 128);
        const uint32_t b_step = P.mode == 0 ? 2u : 64u;  // 8 tf32 of K: 32 bytes along a K-major row / two 512-byte atoms
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint64_t a_desc = ptx::umma_desc_kmajor_sw128(sA + h * A_BYTES);
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk)
            ptx::umma_tf32(tmem_base + h * BN, a_desc + 2 * kk, b_desc + b_step * kk, idesc, (first == 0) || kk != 0);
        }
        first = 0;
        ptx::umma_commit(empty_bar(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    if (any_k) ptx::umma_commit(done_bar);
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int quad = warp & 3;
    if (any_k) {
      ptx::mbar_wait(done_bar, 0);
      ptx::tc_fence_after();
    }
    const int ncol = P.mode == 0 ? P.N : P.Nl[lvl];  // row length of the output matrix
    float* out_b = P.out[lvl] + ((long long)slice * gridDim.y + b) * P.C * ncol;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const int row = h * BM + quad * 32 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        if (any_k) {
          ptx::tmem_ld_32x32b_x32(tmem_base + h * BN + c0 + ((uint32_t)(quad * 32) << 16), r);
          ptx::tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        if (P.mode == 0) {
          // row = query, columns = channels: for a fixed channel the warp writes 32 consecutive queries
          const int q = tile0 + row;
          if (q < P.N) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < P.C) out_b[(long long)(c0 + j) * ncol + q] = __uint_as_float(r[j]) * P.alpha;
          }
        } else {
          // row = channel, columns = targets: each lane writes 32 consecutive targets of its channel row
          const int t0 = tile0 + c0;
          if (row < P.C && t0 < ncol) {
            float* dst = out_b + (long long)row * ncol + t0;
            if (t0 + 32 <= ncol && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]) * P.alpha, __uint_as_float(r[j + 1]) * P.alpha,
                                                                  __uint_as_float(r[j + 2]) * P.alpha, __uint_as_float(r[j + 3]) * P.alpha);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (t0 + j < ncol) dst[j] = __uint_as_float(r[j]) * P.alpha;
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

// dB_l[0][i] += dB_l[1][i] + ... + dB_l[n - 1][i]: the K slices of a coarse level, added in a fixed order (deterministic)
__global__ void __launch_bounds__(256) reduce_slices_kernel(float* __restrict__ p, long long part, int n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < part; i += (long long)gridDim.x * blockDim.x) {
    float sum = p[i];
    for (int sl = 1; sl < n; ++sl) sum += __ldg(p + sl * part + i);
    p[i] = sum;
  }
}

// dfmap2[b,c,y,x] += sum_{l>=1} dB_l[b,c,y>>l,x>>l] / 4^l   (adjoint of l floor 2x2 average pools; pixels outside the
// pooled area of a level get nothing from it)
struct UnpoolParams {
  const float* dB[MAXL];
  int h[MAXL], w[MAXL];
  int L, H, W;
  long long planes;
};
__global__ void __launch_bounds__(256) unpool_add_kernel(float* __restrict__ df2, const UnpoolParams U) {
  const long long total = U.planes * U.H * U.W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % U.W);
    const long long r = i / U.W;
    const int y = (int)(r % U.H);
    const long long pl = r / U.H;
    float acc = 0.f, wgt = 0.25f;
    for (int l = 1; l < U.L; ++l, wgt *= 0.25f) {
      const int yl = y >> l, xl = x >> l;
      if (yl < U.h[l] && xl < U.w[l]) acc += wgt * __ldg(U.dB[l] + (pl * U.h[l] + yl) * U.w[l] + xl);
    }
    df2[i] += acc;
  }
}

}  // namespace k1btc

// true when the tcgen05 path covers this problem (TMA needs 16-byte-aligned row strides)
bool corr_bwd_tc_supported(int C, int N) { return C <= 256 && (N % 4) == 0; }

// K-split of the df2 GEMM per level: 1, 4, 16, 32 slices (at most one per 256-query tile)
static int df2_split(int l, int N) {
  static const bool off = getenv("EZB_BWD_NOSPLIT") != nullptr;  // A/B switch
  return (l == 0 || off) ? 1 : std::max(1, std::min(std::min(1 << (2 * l), 32), cdiv(N, 256)));
}

// scratch: pooled fmap2 levels (padded rows) and the per-level dB matrices (one per K slice)
size_t corr_bwd_tc_workspace_bytes(int B, int C, int H, int W, int L) {
  size_t n = 1024;
  for (int l = 1; l < L; ++l) {
    const long long nl = (long long)(H >> l) * (W >> l);
    n += (size_t)B * C * round_up((int)nl, 4) * 4 + 256;          // pool_l(f2)
    n += (size_t)df2_split(l, H * W) * B * C * nl * 4 + 256;     // dB_l partial sums
  }
  return n;
}

// dlvl[l]: gradient level l, row (b*N + q) at dlvl[l] + row*qstride[l], targets contiguous.  Computes dfmap1 and/or dfmap2.
int corr_bwd_tc_launch(float* const* dlvl, const long long* qstride, const float* fmap1, const float* fmap2, int B, int C, int H,
                       int W, int L, float alpha, float* dfmap1, float* dfmap2, void* workspace, size_t workspace_bytes,
                       const int32_t* ranges, cudaStream_t st) {
  using namespace k1btc;
  const int N = H * W;
  EZB_REQUIRE(L <= MAXL, "at most %d levels", MAXL);
  EZB_REQUIRE((long long)B * N < (1ll << 31), "batch too large for one launch (split it)");
  if (corr_bwd_tc_workspace_bytes(B, C, H, W, L) > workspace_bytes)
    return fail(EZB_ERR_WORKSPACE, "workspace too small: need %zu, got %zu", corr_bwd_tc_workspace_bytes(B, C, H, W, L), workspace_bytes);
  uintptr_t wp = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
  auto take = [&](size_t bytes) {
    uintptr_t r = wp;
    wp = (wp + bytes + 255) & ~uintptr_t(255);
    return reinterpret_cast<float*>(r);
  };
  int Nl[MAXL], hl[MAXL], wl[MAXL], Npad[MAXL];
  const float* f2l[MAXL];
  float* dB[MAXL];
  for (int l = 0; l < L; ++l) {
    hl[l] = H >> l; wl[l] = W >> l; Nl[l] = hl[l] * wl[l]; Npad[l] = round_up(Nl[l], 4);
  }
  f2l[0] = fmap2; Npad[0] = N; dB[0] = dfmap2;
  for (int l = 1; l < L; ++l) {
    float* p = take((size_t)B * C * Npad[l] * 4);
    dB[l] = take((size_t)df2_split(l, N) * B * C * Nl[l] * 4);
    if (dfmap1) {
      if (int r = prep::launch_pool2x2(f2l[l - 1], p, (long long)B * C, hl[l - 1], wl[l - 1], Npad[l - 1], hl[l], wl[l], Npad[l], st)) return r;
    }
    f2l[l] = p;
  }
  EZB_CUDA_OK(cudaFuncSetAttribute(corr_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  for (int mode = 0; mode < 2; ++mode) {
    if (mode == 0 ? dfmap1 == nullptr : dfmap2 == nullptr) continue;
    Maps M;
    memset(&M, 0, sizeof(M));
    Params P;
    P.mode = mode; P.C = C; P.N = N; P.L = L; P.alpha = alpha;
    P.ranges = reinterpret_cast<const int2*>(ranges); P.nT = cdiv(N, 256);
    int tiles = 0;
    for (int l = 0; l < MAXL; ++l) {
      P.Nl[l] = l < L ? Nl[l] : 0;
      P.tile_base[l] = tiles;
      P.split[l] = l < L ? df2_split(l, N) : 1;
      if (l < L) tiles += cdiv(Nl[l], 256) * (mode == 1 ? P.split[l] : 1);
      P.out[l] = mode == 0 ? dfmap1 : (l < L ? dB[l] : nullptr);
    }
    P.tile_base[MAXL] = tiles;
    for (int l = 0; l < L; ++l) {
      cuuint64_t dims[2] = {(cuuint64_t)Nl[l], (cuuint64_t)B * N};
      cuuint64_t str[1] = {(cuuint64_t)qstride[l] * 4};
      cuuint32_t box[2] = {BK, (cuuint32_t)(mode == 0 ? BM : BK)};
      if (int r = make_tmap(&M.g[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dlvl[l], dims, str, box,
                            mode == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))
        return r;
    }
    const int nf = mode == 0 ? L : 1;
    for (int l = 0; l < nf; ++l) {
      const float* F = mode == 0 ? f2l[l] : fmap1;
      const int n = mode == 0 ? Nl[l] : N, npad = mode == 0 ? Npad[l] : N;
      cuuint64_t dims[3] = {(cuuint64_t)n, (cuuint64_t)C, (cuuint64_t)B};
      cuuint64_t str[2] = {(cuuint64_t)npad * 4, (cuuint64_t)C * npad * 4};
      cuuint32_t box[3] = {BK, (cuuint32_t)(mode == 0 ? BN : BM), 1};
      if (int r = make_tmap(&M.f[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(F), dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return r;
    }
    dim3 grid(mode == 0 ? cdiv(N, 256) : tiles, B);
    EZB_REQUIRE(grid.y <= 65535, "batch too large for one launch (split it)");
    corr_bwd_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(M, P);
    EZB_LAUNCH_OK();
    if (mode == 1 && L > 1) {
      UnpoolParams U;
      for (int l = 1; l < L; ++l) {  // the K slices of the coarse levels -> slice 0
        const int ns = df2_split(l, N);
        if (ns == 1) continue;
        const long long part = (long long)B * C * Nl[l];
        reduce_slices_kernel<<<(int)std::min<long long>(cdiv64(part, 256), 148 * 8), 256, 0, st>>>(dB[l], part, ns);
        EZB_LAUNCH_OK();
      }
      for (int l = 0; l < MAXL; ++l) { U.dB[l] = l < L ? dB[l] : nullptr; U.h[l] = l < L ? hl[l] : 0; U.w[l] = l < L ? wl[l] : 0; }
      U.L = L; U.H = H; U.W = W; U.planes = (long long)B * C;
      const long long total = U.planes * N;
      unpool_add_kernel<<<(int)std::min<long long>(cdiv64(total, 256), 148 * 16), 256, 0, st>>>(dfmap2, U);
      EZB_LAUNCH_OK();
    }
  }
  return EZB_OK;
}

}  // namespace ezb
