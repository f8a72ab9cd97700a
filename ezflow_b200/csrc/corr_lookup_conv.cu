// K3 + convc1 fused: radius-r pyramid lookup -> 1x1 convolution (+ bias, ReLU) in ONE kernel.
//
// Replaces, for RAFT's update block, the pair
//   corr = MutliScalePairwise4DCorr.__call__(coords)      ezflow/similarity/correlation/pairwise.py:43-69
//   cor  = F.relu(self.convc1(corr))                      ezflow/decoder/iterative/recurrent_lookup.py:182,204
// (convc1 = Conv2d(L*(2r+1)^2, 256, 1); SmallMotionEncoder: 96 outputs, :153) so that the (B, L*(2r+1)^2, H, W) lookup
// features — 42 MB per 1080p pair and iteration — are never written to HBM nor re-read by the convolution.
//
// A 1x1 convolution is a GEMM over the channel axis:  out[b, n, q] = act( sum_k Wt[n, k] * feat[b, k, q] + bias[n] ).
// Per tile of 128 queries the lookup features are produced level by level straight into a shared-memory A operand
// (fp16, K-major, SWIZZLE_64B; one level = 81 taps padded to 96 = 3 k-blocks of 32) and one tcgen05.mma chain per level
// accumulates against that level's slice of the weights (TMA, L2-resident) into a 128 x Cout fp32 accumulator in
// tensor memory.  The epilogue applies the scales, bias and ReLU and stores NCHW with 128-byte coalesced rows.
//
//   warps 0-15  gather (cp.async windows, two (tile, level) stages ahead in flight) + bilinear taps of the current stage -> A
//   warps 16-19 epilogue (one TMEM lane quadrant each)
//   warp  20    weight-slice TMA + MMA issue (one elected lane), TMEM allocation
//
// Values: A holds the STORED pyramid domain (fp16 storage scale), the weights are pre-scaled by a power of two into
// (-2, 2); out = act(acc * deq_scale[b] * wscale + bias).  fp16 operands / fp32 accumulation: <= 1e-3 relative to the
// fp32 chain (tests/test_gpu_lookup_conv.py).  fp16 pyramids only (the default); callers fall back to the two
// kernels otherwise.
#include "ezb_common.cuh"
#include "ezb_ptx.cuh"
#include "ezb_tma.cuh"

#include <type_traits>

namespace ezb {
namespace k3c {

constexpr int TQ = 128;                 // queries per tile = UMMA M = TMEM lanes
constexpr int MAX_R = 4;                // (2r+1)^2 <= 81 taps per level
constexpr int KL = 96;                  // K slots per level (81 padded to a multiple of UMMA_K = 16)
constexpr int KB_EL = 32;               // fp16 elements per SWIZZLE_64B row
constexpr int KBS = KL / KB_EL;         // k-blocks per level (3)
constexpr int N_MAX = 256;              // output channels (UMMA N)
constexpr int A_KB_BYTES = TQ * 64;     // 8 KB
constexpr int B_KB_BYTES = N_MAX * 64;  // 16 KB
constexpr int ROW_EL = 24;              // fetched pixels per window row: 3 chunks of 8 (covers 2r+2 <= 10 taps at any alignment)
constexpr int QSTRIDE = 496;            // bytes per query in a staging buffer: 10 rows x 48 B = 480, odd number of 16-byte chunks
constexpr int STG_BYTES = TQ * QSTRIDE; // 63488
constexpr int COMPUTE_WARPS = 16, EPI_WARPS = 4;  // 4 compute warps per scheduler hide the tap arithmetic's latencies
constexpr int NUM_THREADS = (COMPUTE_WARPS + EPI_WARPS + 1) * 32;  // 672
constexpr int MMA_WARP = COMPUTE_WARPS + EPI_WARPS;
constexpr int QPW = TQ / COMPUTE_WARPS;  // queries whose windows one warp fetches per stage
constexpr int SMEM_STG = 0;
constexpr int SMEM_A = SMEM_STG + 2 * STG_BYTES;           // 126976
constexpr int SMEM_B = SMEM_A + 2 * KBS * A_KB_BYTES;      // + 49152
constexpr int SMEM_BIAS = SMEM_B + KBS * B_KB_BYTES;       // + 49152
constexpr int SMEM_BAR = SMEM_BIAS + N_MAX * 4;
constexpr int SMEM_BYTES = SMEM_BAR + 128 + 1024;
static_assert(SMEM_A % 1024 == 0 && SMEM_B % 1024 == 0 && SMEM_BYTES <= 232448, "shared-memory plan");

struct Params {
  const char* pyr;
  int H, W, N, L, r, B;
  int sw[PYR_MAX_LEVELS], sub_base[PYR_MAX_LEVELS];
  int n_tiles, n_groups;
  long long group_bytes;
  const float* deq;     // [B]
  const float* coords;  // (B,2,N)
  const float* wscale;  // [1]
  const float* bias;    // [Cout] or null
  float* out;           // (B, Cout, N)
  int Cout, Np;         // Np = Cout rounded up to 16 (UMMA N)
  int relu;
  int tiles_per_pair, n_work;
};

__device__ __forceinline__ int floor_clamped(float v) {
  float f = floorf(v);
  f = fminf(fmaxf(f, -16777216.f), 16777216.f);
  return (f == f) ? (int)f : -(1 << 24);
}
__device__ __forceinline__ void cp_async_16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// Waits that normally last microseconds (a whole stage / tile): mbarrier.try_wait with a suspend-time hint parks the thread in
// hardware until the phase completes (or the hint expires), instead of polling and competing with the compute warps for issue slots.
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity, unsigned ns) {
  uint32_t spins = 0, ok = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(ns)
        : "memory");
    if (ok) break;
    if (++spins > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void compute_bar() { asm volatile("bar.sync 1, %0;" ::"n"(COMPUTE_WARPS * 32) : "memory"); }

// K-major SWIZZLE_64B shared-memory matrix descriptor: rows of 64 B, 8-row groups 512 B apart
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;  // SWIZZLE_64B
  return d;
}
// byte offset of element (row, k) inside one level's A operand (KBS k-blocks of [TQ rows][32 fp16], SWIZZLE_64B:
// the 16-byte chunk index is XORed with address bits [7,9) = (row >> 1) & 3)
__device__ __forceinline__ uint32_t a_offset(int row, int k) {
  const int kb = k >> 5, kk = k & 31;
  return (uint32_t)(kb * A_KB_BYTES + row * 64 + ((((kk >> 3) ^ (row >> 1)) & 3) << 4) + (kk & 7) * 2);
}

// K slot of window tap (i, j) inside a level's block: every window column i owns K + 1 slots (K is odd), so that the taps
// (i, 2m), (i, 2m + 1) are an aligned fp16 pair = one 4-byte shared-memory store; the odd slot out stays zero.
__host__ __device__ __forceinline__ int tap_slot(int K, int i, int j) { return i * (K + 1) + j; }

// weights (Cout, L*K*K) fp32 -> (N_MAX rows, L*KL) fp16 scaled by 2^-e (e = floor(log2 max|W|)); k slot l*KL + tap_slot(i, j)
// holds input channel l*K*K + i*K + j, everything else is zero.  One block; wscale[0] = 2^e.
__global__ void __launch_bounds__(1024) pack_weights_kernel(const float* __restrict__ w, int Cout, int K, int L, __half* __restrict__ wp,
                                                            float* __restrict__ wscale) {
  const int KK = K * K;
  __shared__ float red[32];
  const int cin = L * KK, total = Cout * cin;
  float m = 0.f;
  for (int i = threadIdx.x; i < total; i += blockDim.x) m = fmaxf(m, fabsf(w[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
  const int eb = (int)((__float_as_uint(m) >> 23) & 0xff);
  const int e = (eb == 0 || eb == 255) ? 0 : max(-126, min(126, eb - 127));
  const float s = __uint_as_float((unsigned int)(127 - e) << 23);
  if (threadIdx.x == 0) wscale[0] = __uint_as_float((unsigned int)(127 + e) << 23);
  const int kp = L * KL;
  for (int i = threadIdx.x; i < N_MAX * kp; i += blockDim.x) {
    const int n = i / kp, k = i - n * kp, l = k / KL, t = k - l * KL;
    const int wi = t / (K + 1), wj = t - wi * (K + 1);  // slot -> tap
    wp[i] = __float2half_rn((n < Cout && wi < K && wj < K) ? w[(long long)n * cin + l * KK + wi * K + wj] * s : 0.f);
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1) lookup_conv1x1_kernel(const __grid_constant__ CUtensorMap tmW, const Params P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  const uint32_t sStg = base + SMEM_STG, sA = base + SMEM_A, sB = base + SMEM_B, sBar = base + SMEM_BAR;
  float* s_bias = reinterpret_cast<float*>(base_ptr + SMEM_BIAS);
  auto a_full = [&](int i) { return sBar + 8u * i; };         // 2
  auto a_empty = [&](int i) { return sBar + 8u * (2 + i); };  // 2
  const uint32_t b_full = sBar + 8u * 4, b_empty = sBar + 8u * 5;
  auto d_full = [&](int i) { return sBar + 8u * (6 + i); };   // 2
  auto d_empty = [&](int i) { return sBar + 8u * (8 + i); };  // 2
  volatile uint32_t* tmem_ptr_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + SMEM_BAR + 8 * 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = P.r, D = 2 * r + 2, K = 2 * r + 1;

  if (warp == MMA_WARP) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmW);
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(a_full(i), COMPUTE_WARPS);
        ptx::mbar_init(a_empty(i), 1);
        ptx::mbar_init(d_full(i), 1);
        ptx::mbar_init(d_empty(i), EPI_WARPS);
      }
      ptx::mbar_init(b_full, 1);
      ptx::mbar_init(b_empty, 1);
      ptx::fence_mbar_init();
    }
    __syncwarp();
    ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_ptr_slot)), 512);
    ptx::tmem_relinquish();
  }
  // zero both A buffers once: the K slots [KK, KL) of every level are never written again
  for (int i = threadIdx.x; i < 2 * KBS * A_KB_BYTES / 16; i += NUM_THREADS)
    reinterpret_cast<uint4*>(base_ptr + SMEM_A)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < N_MAX; i += NUM_THREADS) s_bias[i] = (P.bias != nullptr && i < P.Cout) ? __ldg(P.bias + i) : 0.f;
  fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_slot;

  const int n_my = (P.n_work - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles of this CTA
  const int n_stages = n_my * P.L;

  if (warp < COMPUTE_WARPS) {
    // ================================ gather + taps ================================
    // stage s = (tile ti = s / L of this CTA, level l = s % L).  Iteration s: wait for stage s's windows, issue the
    // copies of stage s + 1 into the other staging buffer, compute stage s into A[s & 1].
    // (pair, first query) of tile ti of this CTA; the integer division is paid once per tile, not per stage
    auto tile_of = [&](int ti, int& b, int& q0) {
      const int work = (int)blockIdx.x + ti * (int)gridDim.x;
      b = work / P.tiles_per_pair;
      q0 = (work - b * P.tiles_per_pair) * TQ;
    };
    int g_ti = 0, g_l = 0, g_b, g_q0;  // stage the next issue_gather call fetches
    tile_of(0, g_b, g_q0);
    auto issue_gather = [&]() {
      const int l = g_l, b = g_b, q0 = g_q0;
      const int s = g_ti * P.L + g_l;
      if (++g_l == P.L) {
        g_l = 0;
        ++g_ti;
        tile_of(g_ti, g_b, g_q0);
      }
      const float* cx_ptr = P.coords + (long long)b * 2 * P.N;
      const float* cy_ptr = cx_ptr + P.N;
      const uint32_t stg = sStg + (uint32_t)(s & 1) * STG_BYTES;
      const float inv = 1.0f / (float)(1 << l);
      const int hl = P.H >> l, wl = P.W >> l;
      // lanes 0..QPW-1: window origin of query warp*QPW + lane of the tile
      int org_packed;
      {
        const int q = min(q0 + warp * QPW + (lane & (QPW - 1)), P.N - 1);
        const int x0 = floor_clamped(__ldg(cx_ptr + q) * inv), y0 = floor_clamped(__ldg(cy_ptr + q) * inv);
        const int ox = max(-D, min(wl, x0 - r)), oy = max(-D, min(hl, y0 - r));
        // first 8-pixel chunk << 19 | (chunks the 2r+2 window columns span - 1) << 16 | first row: a window reaches into the ng-th
        // chunk only when it starts in the last pixels of a chunk; the chunks beyond its last column (each a DRAM line of its
        // own) are not fetched
        const int cf = ((ox + 64) >> 3) - 8;
        org_packed = (cf << 19) | (((((ox + D - 1) + 64) >> 3) - 8 - cf) << 16) | (oy & 0xffff);
      }
      const int ng = (7 + D + 7) >> 3;  // chunks per window row (3 for r = 4)
      const int row = lane / ng >= D ? 0 : lane / ng, gi = lane - (lane / ng) * ng;
      const bool task_ok = lane < D * ng;
      const unsigned int tile_stride = (unsigned int)(P.n_groups * P.group_bytes);
      const char* pair_base = P.pyr + (long long)b * P.n_tiles * P.n_groups * P.group_bytes;
      const uint32_t dst_lane = stg + (uint32_t)(warp * QPW) * QSTRIDE + (row * ROW_EL + gi * 8) * 2;
      const int sb = P.sub_base[l], swl = P.sw[l];
#pragma unroll
      for (int qi = 0; qi < QPW; ++qi) {
        const int pk = __shfl_sync(0xffffffffu, org_packed, qi);
        if (task_ok) {
          const int q = min(q0 + warp * QPW + qi, P.N - 1);
          const int y = (int)(short)(pk & 0xffff) + row;
          const int xc = (pk >> 19) + gi;  // 8-pixel chunk = subtile column
          // a chunk is one row of one 8x8 subtile: record column (s % 4) * 64 + (y % 8) * 8, i.e. section s % 4 of the
          // query's group block, 16 bytes at (y % 8) * 16 of its 128-byte line (ezb_common.cuh)
          const unsigned sidx = (unsigned)(sb + (y >> 3) * swl + xc);
          const bool ok = (unsigned)y < (unsigned)hl && xc >= 0 && xc * 8 < wl && gi <= ((pk >> 16) & 7);
          const char* src = pair_base + (unsigned long long)(sidx >> 2) * tile_stride + (long long)(q >> 5) * P.group_bytes +
                            (((sidx & 3u) << 12) + ((unsigned)(q & 31) << 7) + ((unsigned)(y & 7) << 4));
          cp_async_16(dst_lane + qi * QSTRIDE, ok ? src : pair_base, ok ? 16u : 0u);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };

    if (n_stages > 0) issue_gather();
    const int gq = warp & 3, part = warp >> 2;  // query group (= TMEM lane quadrant) and share of the window columns
    const int qrow = gq * 32 + lane;  // A row / TMEM lane of this thread's query
    // this warp's share of the window columns i: K columns dealt to the 4 parts (at most 3 each for K <= 9).  The share is
    // fixed for the whole kernel, so the swizzled byte offsets of the thread's tap-slot pairs are computed once.
    const int i0 = (K * part) >> 2, ni = ((K * (part + 1)) >> 2) - i0;
    const int swz = (qrow >> 1) & 3;
    uint32_t soff[3][5];  // [column a][pair m]: offset (inside the A row's k-blocks) of slots (i0 + a, 2m), (i0 + a, 2m + 1)
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int m = 0; m < 5; ++m) {
        const int t = tap_slot(K, i0 + a, 2 * m);
        soff[a][m] = (uint32_t)((t >> 5) * A_KB_BYTES + ((((t >> 3) & 3) ^ swz) << 4) + (t & 7) * 2);
      }
    int c_l = 0, c_ti = 0, b, q0;  // stage being computed
    tile_of(0, b, q0);
    for (int s = 0; s < n_stages; ++s) {
      const int l = c_l;
      // stage s + 1 goes into the buffer stage s - 1 was read from (everybody left it at the barrier that ended iteration
      // s - 1), BEFORE waiting for stage s: two stages of windows are in flight while this one is awaited
      if (s + 1 < n_stages) issue_gather();
      else asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      compute_bar();  // stage s has landed for every thread

      const int q = min(q0 + qrow, P.N - 1);
      const float* cx_ptr = P.coords + (long long)b * 2 * P.N;
      const float inv = 1.0f / (float)(1 << l);
      const float cx = __ldg(cx_ptr + q) * inv, cy = __ldg(cx_ptr + P.N + q) * inv;
      const float fx = cx - floorf(cx), fy = cy - floorf(cy);
      const float wl0 = 1.f - fx, wl1 = fx;
      const int xs = max(-D, min(P.W >> l, floor_clamped(cx) - r));
      const int off = xs - ((((xs + 64) >> 3) - 8) << 3);  // taps start `off` pixels into the fetched row
      const __half* pl = reinterpret_cast<const __half*>(base_ptr + SMEM_STG + (s & 1) * STG_BYTES + qrow * QSTRIDE) + off + i0;

      const int abuf = s & 1;
      ptx::mbar_wait(a_empty(abuf), (((uint32_t)s >> 1) & 1u) ^ 1u);  // the MMAs that read this buffer two stages ago are done
      uint8_t* arow = base_ptr + SMEM_A + abuf * (KBS * A_KB_BYTES) + qrow * 64;
      // Walk down the window rows once: a row's ni + 1 pixels give the ni horizontally interpolated values, two
      // consecutive rows give the taps (i, j); taps (i, 2m), (i, 2m + 1) leave as one half2.
      float top[3], held[3];
      auto hrow = [&](const __half* p, float (&h)[3]) {
        float x[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) x[a] = (a <= ni) ? __half2float(p[a]) : 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) h[a] = wl0 * x[a] + wl1 * x[a + 1];
      };
      hrow(pl, top);
      const __half* p = pl;
      auto body = [&](int j, auto unrolled_tag) {
        p += ROW_EL;
        float bot[3];
        hrow(p, bot);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float v = top[a] + fy * (bot[a] - top[a]);
          top[a] = bot[a];
          if (a < ni) {
            uint32_t so;
            if constexpr (decltype(unrolled_tag)::value) {
              so = soff[a][j >> 1];  // compile-time indices: registers
            } else {
              const int t = tap_slot(K, i0 + a, j & ~1);
              so = (uint32_t)((t >> 5) * A_KB_BYTES + ((((t >> 3) & 3) ^ swz) << 4) + (t & 7) * 2);
            }
            if (j & 1) *reinterpret_cast<__half2*>(arow + so) = __floats2half2_rn(held[a], v);
            else if (j == K - 1) *reinterpret_cast<__half*>(arow + so) = __float2half_rn(v);  // odd K: the last tap is alone
            else held[a] = v;
          }
        }
      };
      if (K == 9) {
#pragma unroll
        for (int j = 0; j < 9; ++j) body(j, std::true_type{});
      } else {
        for (int j = 0; j < K; ++j) body(j, std::false_type{});
      }
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(a_full(abuf));
      compute_bar();  // everybody is done reading this stage's windows: the buffer may be refilled
      if (++c_l == P.L) {
        c_l = 0;
        ++c_ti;
        tile_of(c_ti, b, q0);
      }
    }
  } else if (warp == MMA_WARP) {
    // ================================ weight TMA + MMA issue ================================
    if (lane == 0) {
      const uint32_t idesc = ptx::umma_idesc(/*f16*/ 0, TQ, (uint32_t)P.Np);
      for (int s = 0; s < n_stages; ++s) {
        const int ti = s / P.L, l = s - ti * P.L;
        mbar_wait_backoff(b_empty, ((uint32_t)s & 1u) ^ 1u, 20000);  // MMAs of stage s - 1 (the previous user of the weight buffer) are done
        ptx::mbar_expect_tx(b_full, (uint32_t)(KBS * B_KB_BYTES));
        for (int kb = 0; kb < KBS; ++kb)
          ptx::tma_load_2d_hint(sB + kb * B_KB_BYTES, &tmW, b_full, l * KL + kb * KB_EL, 0, ptx::kEvictLast);
        const int dbuf = ti & 1;
        if (l == 0) {
          mbar_wait_backoff(d_empty(dbuf), (((uint32_t)ti >> 1) & 1u) ^ 1u, 20000);
          ptx::tc_fence_after();
        }
        mbar_wait_backoff(b_full, (uint32_t)s & 1u, 20000);
        mbar_wait_backoff(a_full(s & 1), ((uint32_t)s >> 1) & 1u, 20000);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + dbuf * N_MAX;
        const uint32_t a_base = sA + (s & 1) * (KBS * A_KB_BYTES);
#pragma unroll
        for (int kb = 0; kb < KBS; ++kb) {
          const uint64_t a_desc = umma_desc_kmajor_sw64(a_base + kb * A_KB_BYTES);
          const uint64_t b_desc = umma_desc_kmajor_sw64(sB + kb * B_KB_BYTES);
#pragma unroll
          for (int kk = 0; kk < KB_EL / 16; ++kk) ptx::umma_f16(d_tmem, a_desc + 2 * kk, b_desc + 2 * kk, idesc, (l | kb | kk) != 0);
        }
        ptx::umma_commit(a_empty(s & 1));
        ptx::umma_commit(b_empty);
        if (l == P.L - 1) ptx::umma_commit(d_full(dbuf));
      }
    }
  } else {
    // ================================ epilogue: lane = query ================================
    const int quad = warp - COMPUTE_WARPS;  // warps 16..19 -> TMEM lane quadrants 0..3 (= warp % 4)
    const float wscale = __ldg(P.wscale);
    for (int ti = 0; ti < n_my; ++ti) {
      const int work = (int)blockIdx.x + ti * (int)gridDim.x;
      const int b = work / P.tiles_per_pair, q0 = (work - b * P.tiles_per_pair) * TQ;
      const int q = q0 + quad * 32 + lane;
      const float sc = __ldg(P.deq + b) * wscale;
      const int dbuf = ti & 1;
      mbar_wait_backoff(d_full(dbuf), ((uint32_t)ti >> 1) & 1u, 100000);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + dbuf * N_MAX + ((uint32_t)(quad * 32) << 16);
      float* o = P.out + (long long)b * P.Cout * P.N + q;
      for (int c0 = 0; c0 < P.Np; c0 += 32) {
        uint32_t v[32];
        if (c0 + 32 <= P.Np) {
          ptx::tmem_ld_32x32b_x32(taddr + c0, v);
        } else {  // Np is a multiple of 16
          uint32_t v16[16];
          ptx::tmem_ld_32x32b_x16(taddr + c0, v16);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = v16[j];
#pragma unroll
          for (int j = 16; j < 32; ++j) v[j] = 0u;
        }
        ptx::tmem_ld_wait();
        if (c0 + 32 >= P.Np) {  // the accumulator is in registers: hand it back to the MMA warp
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(d_empty(dbuf));
        }
        if (q < P.N) {
          float* oc = o + (long long)c0 * P.N;
          if (c0 + 32 <= P.Cout) {  // whole chunk: no per-channel predicate
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float y = fmaf(__uint_as_float(v[j]), sc, s_bias[c0 + j]);
              if (P.relu) y = fmaxf(y, 0.f);
              *oc = y;
              oc += P.N;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (c0 + j < P.Cout) {
                float y = fmaf(__uint_as_float(v[j]), sc, s_bias[c0 + j]);
                if (P.relu) y = fmaxf(y, 0.f);
                oc[(long long)j * P.N] = y;
              }
            }
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) ptx::tmem_dealloc(tmem_base, 512);
}

}  // namespace k3c
}  // namespace ezb

using namespace ezb;

extern "C" int ezb_lookup_conv_packed_weight_bytes(int L, size_t* bytes) {
  EZB_REQUIRE(L >= 1 && L <= PYR_MAX_LEVELS && bytes, "bad arguments");
  *bytes = (size_t)k3c::N_MAX * L * k3c::KL * 2;
  return EZB_OK;
}

extern "C" int ezb_lookup_conv_pack_weights(const float* weight, int Cout, int L, int radius, void* packed, float* wscale, void* stream) {
  EZB_REQUIRE(weight && packed && wscale, "null pointer");
  EZB_REQUIRE(L >= 1 && radius >= 0, "bad arguments");
  if (L > PYR_MAX_LEVELS || radius > k3c::MAX_R || Cout < 1 || Cout > k3c::N_MAX)
    return fail(EZB_ERR_UNSUPPORTED, "fused lookup + 1x1 conv supports num_levels <= %d, corr_radius <= %d, 1 <= out_channels <= %d", PYR_MAX_LEVELS,
                k3c::MAX_R, k3c::N_MAX);
  EZB_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 127) == 0, "packed weight buffer must be 128-byte aligned");
  const int K = 2 * radius + 1;
  k3c::pack_weights_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(weight, Cout, K, L, static_cast<__half*>(packed), wscale);
  EZB_LAUNCH_OK();
  return EZB_OK;
}

extern "C" int ezb_corr_lookup_conv1x1_fwd(const void* pyramid, const float* deq_scale, const float* coords, int B, int H, int W, int L,
                                           int radius, int dtype, const void* packed_weight, const float* wscale, const float* bias,
                                           int Cout, int relu, float* out, void* stream) {
  EZB_REQUIRE(pyramid && deq_scale && coords && packed_weight && wscale && out, "null pointer");
  EZB_REQUIRE(B > 0 && H > 0 && W > 0 && L >= 1 && radius >= 0, "bad shape B=%d H=%d W=%d L=%d r=%d", B, H, W, L, radius);
  if (dtype != EZB_F16) return fail(EZB_ERR_UNSUPPORTED, "fused lookup + 1x1 conv needs an fp16 pyramid");
  if (L > PYR_MAX_LEVELS || radius > k3c::MAX_R || Cout < 1 || Cout > k3c::N_MAX)
    return fail(EZB_ERR_UNSUPPORTED, "fused lookup + 1x1 conv supports num_levels <= %d, corr_radius <= %d, 1 <= out_channels <= %d", PYR_MAX_LEVELS,
                k3c::MAX_R, k3c::N_MAX);
  PyrLayout lay;
  EZB_REQUIRE(pyr_layout(H, W, L, dtype, &lay), "feature map %dx%d too small for %d pyramid levels", H, W, L);
  EZB_REQUIRE((reinterpret_cast<uintptr_t>(pyramid) & 127) == 0, "pyramid buffer must be 128-byte aligned");
  EZB_REQUIRE((long long)lay.n_groups * lay.group_bytes < (1ll << 32), "feature map too large (H*W=%d)", H * W);
  const int N = H * W;
  k3c::Params P;
  P.pyr = static_cast<const char*>(pyramid);
  P.H = H; P.W = W; P.N = N; P.L = L; P.r = radius; P.B = B;
  for (int l = 0; l < PYR_MAX_LEVELS; ++l) { P.sw[l] = lay.sw[l]; P.sub_base[l] = lay.sub_base[l]; }
  P.n_tiles = lay.n_tiles; P.n_groups = lay.n_groups; P.group_bytes = lay.group_bytes;
  P.deq = deq_scale; P.coords = coords; P.wscale = wscale; P.bias = bias; P.out = out;
  P.Cout = Cout; P.Np = round_up(Cout, 16); P.relu = relu;
  P.tiles_per_pair = cdiv(N, k3c::TQ);
  const long long n_work = (long long)P.tiles_per_pair * B;
  EZB_REQUIRE(n_work < (1ll << 31), "batch too large for one launch (split it)");
  P.n_work = (int)n_work;

  CUtensorMap tmW;
  {
    cuuint64_t dims[2] = {(cuuint64_t)L * k3c::KL, (cuuint64_t)k3c::N_MAX};
    cuuint64_t str[1] = {(cuuint64_t)L * k3c::KL * 2};
    cuuint32_t box[2] = {k3c::KB_EL, k3c::N_MAX};
    if (int r = make_tmap(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(packed_weight), dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B))
      return r;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EZB_CUDA_OK(cudaFuncSetAttribute(k3c::lookup_conv1x1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k3c::SMEM_BYTES));
  const int grid = (int)std::min<long long>(n_work, num_sms());
  k3c::lookup_conv1x1_kernel<<<grid, k3c::NUM_THREADS, k3c::SMEM_BYTES, st>>>(tmW, P);
  EZB_LAUNCH_OK();
  return EZB_OK;
}
