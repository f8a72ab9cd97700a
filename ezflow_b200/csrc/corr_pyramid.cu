// K1 + K2: RAFT all-pairs correlation with the average-pool pyramid fused into the GEMM epilogue.
//
// Replaces MutliScalePairwise4DCorr.__init__/.corr (reference ezflow/similarity/correlation/
// pairwise.py:26-41, :78-88):  corr[b,q,t] = sum_c f1[b,c,q] f2[b,c,t] / sqrt(C), then L-1 floor
// 2x2 average pools over the target axes (t = h2*W + w2).
//
// Design:
//   prep   : per-pair absmax -> power-of-two pre-scale -> fp32 (B,C,N) to fp16 K-major (B,N,Cp)
//   gemm   : persistent, warp-specialised tcgen05 kernel.  One CTA tile = 128 queries (TMEM lanes)
//            x one 8x32 *patch* of target pixels (256 TMEM columns, column = y*32+x inside the patch).
//            The patch of f2 (B operand, up to 128 KB) stays resident in shared memory while the CTA
//            sweeps query blocks (A operand streamed through a 4-stage TMA ring).  Because every
//            thread of the epilogue owns one query row and a whole 8x32 target patch, all four
//            pyramid levels (8x32 -> 4x16 -> 2x8 -> 1x4) are reduced in registers from the fp32
//            accumulators and written once, as contiguous 4 KB sections (bulk shared->global copies)
//            of the warp-strip-major pyramid layout (ezb_common.cuh): the pyramid is never re-read.
#include "ezb_common.cuh"
#include "ezb_ptx.cuh"

namespace ezb {
namespace k1 {

constexpr int BM = 128;           // queries per tile  (UMMA M, TMEM lanes)
constexpr int PH = 8, PW = 32;    // target patch      (UMMA N = 256, TMEM columns)
constexpr int BN = PH * PW;
constexpr int BK = 64;            // fp16 elements per 128-byte swizzle row
constexpr int KB_MAX = 4;         // resident B k-blocks  (C <= 256)
constexpr int STAGES = 4;         // A ring
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KB
constexpr int B_KB_BYTES = BN * BK * 2;      // 32 KB
constexpr int SLOT_BYTES = 4096;             // one staged section: 32 queries x 128 B
constexpr int SLOTS_PER_WARP = 2;
constexpr int EPI_WARPS = 4;          // one warp per TMEM lane quadrant
constexpr int NUM_THREADS = 128 + EPI_WARPS * 32;
constexpr int SMEM_B = 0;
constexpr int SMEM_A = SMEM_B + KB_MAX * B_KB_BYTES;
constexpr int SMEM_STG = SMEM_A + STAGES * A_STAGE_BYTES;
constexpr int SMEM_BAR = SMEM_STG + EPI_WARPS * SLOTS_PER_WARP * SLOT_BYTES;
constexpr int SMEM_BYTES = SMEM_BAR + 256 + 1024;  // + barrier block + alignment slack
constexpr int TMEM_COLS = 512;                     // two 256-column accumulators

struct GemmParams {
  int B, N, H, W, KB, L;
  int n_pw, n_patches, n_qblocks, n_groups;
  long long total_tiles, group_bytes;
  const float* epi_scale;  // [B]
  char* pyr;               // pyramid buffer, group blocks ordered [pair][patch][group]
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
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      float4 v = __ldg(x4 + i);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_b;
         i += (long long)gridDim.x * blockDim.x)
      m = fmaxf(m, fabsf(x[i]));
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

// epi_scale[b] multiplies the fp32 accumulator in the GEMM epilogue, deq[b] multiplies stored
// values at lookup time;  epi*deq == 2^(e1+e2)/sqrt(C) always.
__global__ void scales_kernel(const unsigned int* __restrict__ amax, int B, int C, int dtype, float* __restrict__ epi,
                              float* __restrict__ deq) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int e = pow2_exponent(amax[b]) + pow2_exponent(amax[B + b]);
  const float true_scale = pow2f(e) * rsqrtf((float)C);  // exact for power-of-four C, <=1ulp otherwise
  if (dtype == EZB_F16) {
    // operands are pre-scaled into (-2,2): |acc| < 4C.  Store acc * F with F the largest power of two
    // keeping |stored| < 2^15, so fp16 keeps full relative precision over ~19 binades below that.
    int lg = 0;
    while ((1 << lg) < 4 * C) ++lg;
    const float F = pow2f(15 - lg);
    epi[b] = F;
    deq[b] = true_scale / F;
  } else {
    epi[b] = true_scale;
    deq[b] = 1.0f;
  }
}

// fp32 (B,C,N) -> fp16 (B,N,Cp) scaled by 2^-e (per pair, per tensor); channels >= C zero-filled.
__global__ void __launch_bounds__(256) convert_kmajor_kernel(const float* __restrict__ f1, const float* __restrict__ f2,
                                                             __half* __restrict__ o1, __half* __restrict__ o2,
                                                             const unsigned int* __restrict__ amax, int B, int C, int Cp,
                                                             int N) {
  __shared__ float tile[64][33];
  const int which = blockIdx.z / B, b = blockIdx.z % B;
  const float* src = (which ? f2 : f1) + (long long)b * C * N;
  __half* dst = (which ? o2 : o1) + (long long)b * N * Cp;
  const float s = pow2f(-pow2_exponent(amax[which * B + b]));
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

// ----------------------------------------------------------------------------- epilogue helpers
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
  return r;
}
__device__ __forceinline__ void st_global_stream_v4(void* p, const uint4& v) {
  // .cs: streaming / evict-first — the pyramid is written once and must not evict the GEMM operands from L2
  asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// One section of a group block (32 lines, one per lane, NWORDS*4 bytes each).  Every lane writes its own
// line into the warp's staging slot with the layout's XOR chunk swizzle (bank-conflict free); the slot
// is then byte-identical to the section in global memory, and the warp streams it out with fully
// coalesced 16-byte stores: each store instruction writes 512 contiguous bytes.
template <int NWORDS>
__device__ __forceinline__ void staged_section_store(uint32_t slot, int lane, const uint32_t (&w)[NWORDS], char* gdst) {
  static_assert(NWORDS == 32 || NWORDS == 16, "lines are 128 or 64 bytes");
  constexpr uint32_t kSwz = NWORDS / 4 - 1;
  __syncwarp();  // the previous section's read-back of this slot is complete in every lane
#pragma unroll
  for (int c = 0; c < NWORDS / 4; ++c)
    st_shared_v4(slot + lane * (NWORDS * 4) + ((c ^ (lane & kSwz)) << 4), w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
  __syncwarp();
#pragma unroll
  for (int c = 0; c < NWORDS / 4; ++c) {
    const uint4 v = ld_shared_v4(slot + c * 512 + lane * 16);
    st_global_stream_v4(gdst + c * 512 + lane * 16, v);
  }
}

// Epilogue of one tile for one warp: 32 queries (one per lane) x one 8x32 target patch -> one group block.
// Strip s holds patch rows 2s and 2s+1; level-1 row s is their 2x2 average.  Levels 2 and 3 follow
// from level 1, all from the fp32 accumulators (the reference pools fp32 values too).
template <typename OutT>
__device__ __forceinline__ void epilogue_tile(uint32_t t_base, uint32_t stg_base, uint32_t& slot_ctr, int lane,
                                              float scale, uint32_t tmem_empty_bar, int L, char* gblock, int vr1,
                                              int vc1) {
  // vr1 / vc1: number of valid level-1 rows / columns in this tile (4 / 16 for interior tiles).  Level-0
  // values outside the image are exact zeros (TMA zero-fills the f2 patch); pooled values outside their
  // level are forced to zero here so the lookup never needs per-element masking.
  constexpr bool kHalf = sizeof(OutT) == 2;
  auto next_slot = [&]() { return stg_base + (slot_ctr++ & (SLOTS_PER_WARP - 1)) * SLOT_BYTES; };
  float l1[4][16];
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    uint32_t r0[32], r1[32];
    ptx::tmem_ld_32x32b_x32(t_base + (2 * s) * 32, r0);
    ptx::tmem_ld_32x32b_x32(t_base + (2 * s + 1) * 32, r1);
    ptx::tmem_ld_wait();
    if (s == 3) {  // accumulator fully drained into registers: hand the TMEM buffer back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tmem_empty_bar);
    }
    if (gblock == nullptr) continue;  // query group entirely beyond N: nothing to store
    float v0[32], v1[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      v0[j] = __uint_as_float(r0[j]) * scale;
      v1[j] = __uint_as_float(r1[j]) * scale;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j)
      l1[s][j] = (s < vr1 && j < vc1) ? 0.25f * ((v0[2 * j] + v0[2 * j + 1]) + (v1[2 * j] + v1[2 * j + 1])) : 0.f;

    if constexpr (kHalf) {
      uint32_t w[32];  // line = [row 2s: 32 halfs][row 2s+1: 32 halfs]        (section s)
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        w[j] = pack_half2(v0[2 * j], v0[2 * j + 1]);
        w[16 + j] = pack_half2(v1[2 * j], v1[2 * j + 1]);
      }
      staged_section_store<32>(next_slot(), lane, w, gblock + s * 4096);
    } else {
      uint32_t w[32];  // line = one row of 32 floats                            (sections 2s, 2s+1)
#pragma unroll
      for (int j = 0; j < 32; ++j) w[j] = __float_as_uint(v0[j]);
      staged_section_store<32>(next_slot(), lane, w, gblock + (2 * s) * 4096);
#pragma unroll
      for (int j = 0; j < 32; ++j) w[j] = __float_as_uint(v1[j]);
      staged_section_store<32>(next_slot(), lane, w, gblock + (2 * s + 1) * 4096);
      if (L > 1 && (s & 1)) {  // level-1 rows s-1, s                              (section 8 + s/2)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          w[j] = __float_as_uint(l1[s - 1][j]);
          w[16 + j] = __float_as_uint(l1[s][j]);
        }
        staged_section_store<32>(next_slot(), lane, w, gblock + (8 + (s >> 1)) * 4096);
      }
    }
  }
  if (gblock == nullptr || L < 2) return;
  if constexpr (kHalf) {
    uint32_t w[32];  // 4 rows x 16 halfs                                          (section 4)
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) w[r * 8 + j] = pack_half2(l1[r][2 * j], l1[r][2 * j + 1]);
    staged_section_store<32>(next_slot(), lane, w, gblock + 4 * 4096);
  }
  if (L < 3) return;
  float l2[2][8], l3[4];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      l2[r][j] = (r < (vr1 >> 1) && j < (vc1 >> 1))
                     ? 0.25f * ((l1[2 * r][2 * j] + l1[2 * r][2 * j + 1]) + (l1[2 * r + 1][2 * j] + l1[2 * r + 1][2 * j + 1]))
                     : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    l3[j] = ((vr1 >> 2) > 0 && j < (vc1 >> 2)) ? 0.25f * ((l2[0][2 * j] + l2[0][2 * j + 1]) + (l2[1][2 * j] + l2[1][2 * j + 1])) : 0.f;
  if constexpr (kHalf) {
    uint32_t w[16];  // [level 2: 2 x 8 halfs = 32 B][level 3: 4 halfs = 8 B][pad]   (section 5, 64-byte lines)
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int j = 0; j < 4; ++j) w[r * 4 + j] = pack_half2(l2[r][2 * j], l2[r][2 * j + 1]);
    w[8] = pack_half2(l3[0], l3[1]);
    w[9] = pack_half2(l3[2], l3[3]);
#pragma unroll
    for (int j = 10; j < 16; ++j) w[j] = 0u;
    staged_section_store<16>(next_slot(), lane, w, gblock + 5 * 4096);
  } else {
    uint32_t w[32];  // [level 2: 2 x 8 floats = 64 B][level 3: 4 floats = 16 B][pad] (section 10)
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int j = 0; j < 8; ++j) w[r * 8 + j] = __float_as_uint(l2[r][j]);
#pragma unroll
    for (int j = 0; j < 4; ++j) w[16 + j] = __float_as_uint(l3[j]);
#pragma unroll
    for (int j = 20; j < 32; ++j) w[j] = 0u;
    staged_section_store<32>(next_slot(), lane, w, gblock + 10 * 4096);
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
  const uint32_t sB = base + SMEM_B, sA = base + SMEM_A, sStg = base + SMEM_STG, sBar = base + SMEM_BAR;
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
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(b_full, 1);
    ptx::mbar_init(b_empty, 1);
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tmem_full(a), 1);
      ptx::mbar_init(tmem_empty(a), 4);  // one arrive per epilogue warp
    }
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

  // Pairs are processed one after the other by the whole grid (so the L2 working set is ONE pair's
  // operands); within a pair every CTA owns a contiguous, balanced range of tiles.
  // tile t of pair b -> (patch = t / n_qblocks, qb = t % n_qblocks), bp = b * n_patches + patch.
  const long long tiles_per_pair = (long long)p.n_patches * p.n_qblocks;
  const long long t_begin = tiles_per_pair * blockIdx.x / gridDim.x;
  const long long t_end = tiles_per_pair * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    uint32_t stage = 0, phase = 0, uphase = 0;
    long long cur_bp = -1;
    for (int b = 0; b < p.B; ++b)
    for (long long t = t_begin; t < t_end; ++t) {
      const int patch = (int)(t / p.n_qblocks);
      const int qb = (int)(t - (long long)patch * p.n_qblocks);
      const long long bp = (long long)b * p.n_patches + patch;
      if (bp != cur_bp) {
        const int ph = patch / p.n_pw, pw = patch - ph * p.n_pw;
        if (cur_bp >= 0) {  // all MMAs reading the previous patch have completed
          ptx::mbar_wait(b_empty, uphase);
          uphase ^= 1;
        }
        ptx::mbar_expect_tx(b_full, (uint32_t)(p.KB * B_KB_BYTES));
        for (int kb = 0; kb < p.KB; ++kb)
          ptx::tma_load_4d_hint(sB + kb * B_KB_BYTES, &tmB, b_full, kb * BK, pw * PW, ph * PH, b, ptx::kEvictLast);
        cur_bp = bp;
      }
      for (int kb = 0; kb < p.KB; ++kb) {
        ptx::mbar_wait(empty_bar(stage), phase ^ 1);
        ptx::mbar_expect_tx(full_bar(stage), A_STAGE_BYTES);
        // fmap1 is re-read by every patch sweep (136x at 1080p): pin it in L2
        ptx::tma_load_3d_hint(sA + stage * A_STAGE_BYTES, &tmA, full_bar(stage), kb * BK, qb * BM, b, ptx::kEvictLast);
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
      const long long bp = (long long)b * p.n_patches + t / p.n_qblocks;
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
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)  // UMMA_K = 16 fp16 = 32 B = +2 in the descriptor's 16-byte units
          ptx::umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
        ptx::umma_commit(empty_bar(stage));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      ptx::umma_commit(tmem_full(acc));
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access
    const uint32_t stg_base = sStg + (warp - 4) * (SLOTS_PER_WARP * SLOT_BYTES);
    uint32_t slot_ctr = 0, tile = 0;
    long long cur_bp = -1;
    char* patch_base = nullptr;
    float scale = 1.f;
    int vr1 = 4, vc1 = 16;
    for (int b = 0; b < p.B; ++b)
    for (long long t = t_begin; t < t_end; ++t, ++tile) {
      const int patch = (int)(t / p.n_qblocks);
      const int qb = (int)(t - (long long)patch * p.n_qblocks);
      const long long bp = (long long)b * p.n_patches + patch;
      if (bp != cur_bp) {
        const int ph = patch / p.n_pw, pw = patch - ph * p.n_pw;
        vr1 = max(0, min(PH / 2, (p.H >> 1) - ph * (PH / 2)));  // valid level-1 rows / cols of this patch
        vc1 = max(0, min(PW / 2, (p.W >> 1) - pw * (PW / 2)));
        scale = __ldg(p.epi_scale + b);
        patch_base = p.pyr + bp * p.n_groups * p.group_bytes;  // blocks are ordered [pair][patch][group]
        cur_bp = bp;
      }
      const uint32_t acc = tile & 1, aphase = (tile >> 1) & 1;
      ptx::mbar_wait(tmem_full(acc), aphase);
      ptx::tc_fence_after();
      const uint32_t t_base = tmem_base + acc * BN + ((uint32_t)(quad * 32) << 16);
      const int grp = qb * (BM / 32) + quad;  // query group of this warp
      epilogue_tile<OutT>(t_base, stg_base, slot_ctr, lane, scale, tmem_empty(acc), p.L,
                          grp < p.n_groups ? patch_base + (long long)grp * p.group_bytes : nullptr, vr1, vc1);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

// --------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

static int make_tmap(CUtensorMap* tm, CUtensorMapDataType dt, int rank, void* ptr, const cuuint64_t* dims,
                     const cuuint64_t* strides_bytes /*rank-1*/, const cuuint32_t* box, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(EZB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(tm, dt, (cuuint32_t)rank, ptr, dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(EZB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return EZB_OK;
}

struct Workspace {
  unsigned int* amax;  // [2][B]
  float* epi;          // [B]
  __half* f1h;         // (B,N,Cp)
  __half* f2h;
};

static size_t ws_layout(int B, int C, int N, void* base, Workspace* ws) {
  const int Cp = round_up(C, BK);
  uintptr_t p = reinterpret_cast<uintptr_t>(base);
  auto take = [&](size_t bytes, size_t align) {
    p = (p + align - 1) / align * align;
    uintptr_t r = p;
    p += bytes;
    return r;
  };
  const uintptr_t amax = take((size_t)2 * B * 4, 256);
  const uintptr_t epi = take((size_t)B * 4, 256);
  const uintptr_t f1h = take((size_t)B * N * Cp * 2, 1024);
  const uintptr_t f2h = take((size_t)B * N * Cp * 2, 1024);
  if (ws) {
    ws->amax = reinterpret_cast<unsigned int*>(amax);
    ws->epi = reinterpret_cast<float*>(epi);
    ws->f1h = reinterpret_cast<__half*>(f1h);
    ws->f2h = reinterpret_cast<__half*>(f2h);
  }
  return p - reinterpret_cast<uintptr_t>(base);
}

template <typename OutT>
static int launch_gemm(const Workspace& ws, int B, int C, const PyrLayout& lay, void* pyr, cudaStream_t st, cudaEvent_t ev0,
                       cudaEvent_t ev1) {
  const int H = lay.H, W = lay.W, N = H * W, Cp = round_up(C, BK);
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[3] = {(cuuint64_t)Cp, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t str[2] = {(cuuint64_t)Cp * 2, (cuuint64_t)N * Cp * 2};
    cuuint32_t box[3] = {BK, BM, 1};
    int r = make_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, ws.f1h, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) return r;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)Cp * 2, (cuuint64_t)W * Cp * 2, (cuuint64_t)N * Cp * 2};
    cuuint32_t box[4] = {BK, PW, PH, 1};
    int r = make_tmap(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, ws.f2h, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (r) return r;
  }
  GemmParams p;
  p.B = B;
  p.N = N;
  p.H = H;
  p.W = W;
  p.KB = Cp / BK;
  p.L = lay.L;
  p.n_pw = lay.n_pw;
  p.n_patches = lay.n_patches;
  p.n_qblocks = cdiv(N, BM);
  p.n_groups = lay.n_groups;
  p.group_bytes = lay.group_bytes;
  p.total_tiles = (long long)B * p.n_patches * p.n_qblocks;
  p.epi_scale = ws.epi;
  p.pyr = static_cast<char*>(pyr);

  auto kern = corr_pyramid_gemm_kernel<OutT>;
  EZB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  long long grid = num_sms();
  if (grid > p.total_tiles / B) grid = p.total_tiles / B;  // tiles of one pair
  if (ev0) EZB_CUDA_OK(cudaEventRecord(ev0, st));
  kern<<<(unsigned)grid, NUM_THREADS, SMEM_BYTES, st>>>(tmA, tmB, p);
  EZB_LAUNCH_OK();
  if (ev1) EZB_CUDA_OK(cudaEventRecord(ev1, st));
  return EZB_OK;
}

}  // namespace k1
}  // namespace ezb

using namespace ezb;

extern "C" int ezb_corr_pyramid_layout(int H, int W, int L, int dtype, int* h, int* w, int* n_ph, int* n_pw, int* n_groups,
                                       int64_t* group_bytes, int64_t* pair_bytes) {
  EZB_REQUIRE(H > 0 && W > 0 && L >= 1, "bad pyramid shape H=%d W=%d L=%d", H, W, L);
  EZB_REQUIRE(dtype == EZB_F32 || dtype == EZB_F16, "bad dtype %d", dtype);
  if (L > PYR_MAX_LEVELS) return fail(EZB_ERR_UNSUPPORTED, "num_levels %d > %d is not supported by the fused pyramid", L, PYR_MAX_LEVELS);
  PyrLayout lay;
  EZB_REQUIRE(pyr_layout(H, W, L, dtype, &lay), "feature map %dx%d too small for %d pyramid levels", H, W, L);
  for (int l = 0; l < L; ++l) {
    if (h) h[l] = H >> l;
    if (w) w[l] = W >> l;
  }
  if (n_ph) *n_ph = lay.n_ph;
  if (n_pw) *n_pw = lay.n_pw;
  if (n_groups) *n_groups = lay.n_groups;
  if (group_bytes) *group_bytes = lay.group_bytes;
  if (pair_bytes) *pair_bytes = lay.pair_bytes;
  return EZB_OK;
}

extern "C" int64_t ezb_corr_pyramid_elem_offset(int dtype, int level, int lane, int yy, int xx) {
  if ((dtype != EZB_F32 && dtype != EZB_F16) || level < 0 || level >= PYR_MAX_LEVELS || lane < 0 || lane >= PYR_QG || yy < 0 ||
      yy >= (PYR_PH >> level) || xx < 0 || xx >= (PYR_PW >> level))
    return -1;
  return pyr_elem_offset(dtype == EZB_F16 ? 2 : 4, level, lane, yy, xx);
}

extern "C" int ezb_corr_pyramid_workspace_bytes(int B, int C, int H, int W, size_t* bytes) {
  EZB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && bytes, "bad arguments");
  *bytes = k1::ws_layout(B, C, H * W, nullptr, nullptr) + 1024;
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
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int N = H * W, Cp = round_up(C, k1::BK);

  // carve the workspace from a 1024-aligned base
  uintptr_t wbase = (reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023);
  k1::Workspace ws;
  const size_t need = k1::ws_layout(B, C, N, reinterpret_cast<void*>(wbase), &ws) + (wbase - reinterpret_cast<uintptr_t>(workspace));
  if (need > workspace_bytes) return fail(EZB_ERR_WORKSPACE, "workspace too small: need %zu, got %zu", need, workspace_bytes);

  EZB_CUDA_OK(cudaMemsetAsync(ws.amax, 0, (size_t)2 * B * 4, st));
  {
    const long long per_b = (long long)C * N;
    int gx = (int)std::min<long long>(cdiv64(per_b, 256 * 16), 4096);
    if (gx < 1) gx = 1;
    k1::absmax_kernel<<<dim3(gx, B, 2), 256, 0, st>>>(fmap1, fmap2, per_b, ws.amax, B);
    EZB_LAUNCH_OK();
  }
  k1::scales_kernel<<<cdiv(B, 128), 128, 0, st>>>(ws.amax, B, C, dtype, ws.epi, deq_scale);
  EZB_LAUNCH_OK();
  k1::convert_kmajor_kernel<<<dim3(cdiv(N, 32), Cp / 64, 2 * B), 256, 0, st>>>(fmap1, fmap2, ws.f1h, ws.f2h, ws.amax, B, C,
                                                                              Cp, N);
  EZB_LAUNCH_OK();

  cudaEvent_t ev0 = static_cast<cudaEvent_t>(ev_gemm_start), ev1 = static_cast<cudaEvent_t>(ev_gemm_end);
  return dtype == EZB_F16 ? k1::launch_gemm<__half>(ws, B, C, lay, pyramid, st, ev0, ev1)
                          : k1::launch_gemm<float>(ws, B, C, lay, pyramid, st, ev0, ev1);
}

extern "C" int ezb_corr_pyramid_fwd(const float* fmap1, const float* fmap2, int B, int C, int H, int W, int L, int dtype,
                                    void* pyramid, float* deq_scale, void* workspace, size_t workspace_bytes, void* stream) {
  return ezb_corr_pyramid_fwd_ex(fmap1, fmap2, B, C, H, W, L, dtype, pyramid, deq_scale, workspace, workspace_bytes, stream,
                                 nullptr, nullptr);
}
