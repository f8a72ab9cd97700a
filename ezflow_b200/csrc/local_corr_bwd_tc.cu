// K4b on tensor cores: both adjoints of the local displacement-window correlation as banded tcgen05 GEMMs.
//
// Autograd of iter_spatial_correlation_sample (reference ezflow/similarity/correlation/sampler.py:96-129) for the regular
// case — one patch dilation (handled as a lattice decomposition when > 1), stride 1, no padding, C <= 256, odd P <= 25:
//   din1[b,c,y,x]   = scale * sum_{pi,pj} dout[b,pi,pj,y,x]          * in2[b,c,y+dy,x+dx]        (dy,dx) = (pi-md, pj-md) * dil
//   din2[b,c,y',x'] = scale * sum_{pi,pj} dout[b,pi,pj,y'-dy,x'-dx]  * in1[b,c,y'-dy,x'-dx]
// Both are  D[q, c] = sum_t G[q, t] * F[t, c]  with G the window band of the all-pairs matrix — the transpose of what the
// forward kernel (local_corr_tc.cu) extracts: per 8x16 tile of output pixels q (UMMA M = 128) the CTA sweeps the same 8x24 /
// 4x40 tiles of the other feature map (targets t), and
//   A = G tile   [128 q][BN t]  fp16, K-major, SWIZZLE_128B, BUILT in shared memory from dout: a lane owns a query row, writes
//                the window rows that fall into the tile at their column positions and zeros elsewhere (double-buffered);
//   B = F tile   [BN t][64 c]   the K-major fp16 feature rows exactly as the forward's TMA ring loads them — read by the
//                tensor core as an MN-major operand (channels contiguous), one 64-channel chunk per ring stage;
//   D = 128 x C fp32 in tensor memory, accumulated over all target tiles (two accumulators: the epilogue of one query tile
//                overlaps the next one's MMAs), written to NCHW with 64-byte coalesced rows.
// din2 is the same computation with the roles of in1 / in2 swapped and dout read at the partner pixel with the flipped
// displacement:  G2[t, (pi,pj)] = dout[b, P-1-pi, P-1-pj, t + (pi-md, pj-md)].
// Operands are scaled by one power of two per (batch, group) and tensor (the contraction runs over pixels, so per-pixel
// scales cannot be factored out here); fp16 operands / fp32 accumulation: <= 1e-3 relative ('tc' mode); 'exact' mode keeps the
// fp32 CUDA-core kernels of local_corr.cu.
#include "ezb_common.cuh"
#include "ezb_ptx.cuh"
#include "ezb_tma.cuh"

#include <cstdlib>

namespace ezb {
namespace k4btc {

constexpr int QTH = 8, QTW = 16, BM = QTH * QTW;
constexpr int CH = 64;                  // channels per B stage = UMMA N per instruction
constexpr int KBC_MAX = 4;              // C <= 256
constexpr int STAGES = 4;
constexpr int BUILD_WARPS = 16, EPI_WARPS = 4;
constexpr int NUM_THREADS = (4 + BUILD_WARPS + EPI_WARPS) * 32;  // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4..19 build, 20..23 epilogue
constexpr int ACC_COLS = 256;

template <int TTW_, int TTH_>
struct Geo {
  static constexpr int TTW = TTW_, TTH = TTH_, BN = TTW * TTH;
  static constexpr int A_KBLOCKS = (BN + 63) / 64;
  static constexpr int A_BYTES = A_KBLOCKS * BM * 128;
  static constexpr int B_STAGE_BYTES = BN * 128;
  static constexpr int CPW = BN / 4;        // columns (targets) each of a quadrant's 4 builder warps fills
  static constexpr int RPW = CPW / TTW;     // = whole tile rows
  static constexpr int SMEM_A = 0;
  static constexpr int SMEM_B = 2 * A_BYTES;
  static constexpr int SMEM_BAR = SMEM_B + STAGES * B_STAGE_BYTES;
  static constexpr int SMEM_BYTES = SMEM_BAR + 256 + 1024;
  static_assert(BN % 16 == 0 && CPW % TTW == 0 && (TTW * 2) % 16 == 0 && B_STAGE_BYTES % 1024 == 0 && SMEM_BYTES <= 232448, "tile geometry");
};

struct Params {
  int G, P, lat, BG;
  int Hf, Wf;         // full-resolution image
  int Hu, Wu;         // size of residue class (0,0) in lattice units (= Hf, Wf when lat == 1)
  int qtx, qty, ncls, KBC, C, n_items;
  int modes;          // bit 0: din1, bit 1: din2
  int accumulate;     // add to the outputs instead of overwriting them (K6: one call per dilation)
  const float* dout;  // (B, DG, P, P, Hf, Wf) view: batch stride dout_bstride, this dilation's first plane at dout
  long long dout_bstride;
  const unsigned int* amax;  // [3][BG]: float bits of max|in1|, max|in2| per (batch, group), max|dout| per batch entry (index b)
  float scale;
  float* din1;
  float* din2;
};

__device__ __forceinline__ int pow2_exp(unsigned int bits) {
  const int e = (int)((bits >> 23) & 0xff);
  return (e == 0 || e == 255) ? 0 : max(-126, min(126, e - 127));
}
__device__ __forceinline__ float pow2f(int e) { return __uint_as_float((unsigned int)(max(-126, min(126, e)) + 127) << 23); }

// max |dout| per batch entry (dout is a strided view: `planes` contiguous planes of `plane` elements per entry)
__global__ void __launch_bounds__(256) absmax_dout_kernel(const float* __restrict__ dout, long long bstride, long long per_b,
                                                          unsigned int* __restrict__ amax) {
  const float* x = dout + (long long)blockIdx.y * bstride;
  float m = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x, t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {  // 16-byte loads, four in flight per thread
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const long long n4 = per_b >> 2;
    long long i = t0;
    for (; i + 3 * stride < n4; i += 4 * stride) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(x4 + i + u * stride);
#pragma unroll
      for (int u = 0; u < 4; ++u) m = fmaxf(m, fmaxf(fmaxf(fabsf(v[u].x), fabsf(v[u].y)), fmaxf(fabsf(v[u].z), fabsf(v[u].w))));
    }
    for (; i < n4; i += stride) {
      const float4 v = __ldg(x4 + i);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    for (long long k = (n4 << 2) + t0; k < per_b; k += stride) m = fmaxf(m, fabsf(__ldg(x + k)));
  } else {
    for (long long i = t0; i < per_b; i += stride) m = fmaxf(m, fabsf(__ldg(x + i)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax + blockIdx.y, __float_as_uint(m));
}

// MN-major fp16 operand, SWIZZLE_128B: rows (K) of 128 bytes = 64 contiguous MN elements, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t desc_mnmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;            // LBO: one 64-element atom along MN, unused
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO: next group of 8 K rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;            // SWIZZLE_128B
  return d;
}

template <int TTW_, int TTH_>
__global__ void __launch_bounds__(NUM_THREADS, 1)
    local_corr_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmF2, const __grid_constant__ CUtensorMap tmF1, const Params P) {
  using G = Geo<TTW_, TTH_>;
  constexpr int TTW = G::TTW, TTH = G::TTH, BN = G::BN, A_BYTES = G::A_BYTES, B_STAGE_BYTES = G::B_STAGE_BYTES, RPW = G::RPW;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  const uint32_t sA = base + G::SMEM_A, sB = base + G::SMEM_B, sBar = base + G::SMEM_BAR;
  auto b_full = [&](int s) { return sBar + 8u * s; };
  auto b_empty = [&](int s) { return sBar + 8u * (STAGES + s); };
  auto a_full = [&](int i) { return sBar + 8u * (2 * STAGES + i); };
  auto a_empty = [&](int i) { return sBar + 8u * (2 * STAGES + 2 + i); };
  auto d_full = [&](int i) { return sBar + 8u * (2 * STAGES + 4 + i); };
  auto d_empty = [&](int i) { return sBar + 8u * (2 * STAGES + 6 + i); };
  volatile uint32_t* tmem_ptr_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + G::SMEM_BAR + 8 * (2 * STAGES + 8));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int md = (P.P - 1) >> 1;

  // geometry of a work item (every role walks the same sequence): item -> (mode, bg, class, query tile)
  struct Item {
    int mode, bg, cy, cx, qy0, qx0, H_, W_, y_lo, x_lo, ntiles;
    bool valid;
  };
  auto item_geometry = [&](int item) {
    Item it;
    const int per_mode = P.n_items / ((P.modes == 3) ? 2 : 1);
    const int mi = item / per_mode;
    it.mode = (P.modes == 3) ? mi : (P.modes == 2 ? 1 : 0);
    int rest = item - mi * per_mode;
    const int bx = rest % P.qtx;
    rest /= P.qtx;
    const int by = rest % P.qty;
    rest /= P.qty;
    const int cls = rest % P.ncls;
    it.bg = rest / P.ncls;
    it.cy = cls / P.lat;
    it.cx = cls - it.cy * P.lat;
    it.qx0 = bx * QTW;
    it.qy0 = by * QTH;
    it.H_ = P.lat > 1 ? (P.Hf - it.cy + P.lat - 1) / P.lat : P.Hf;
    it.W_ = P.lat > 1 ? (P.Wf - it.cx + P.lat - 1) / P.lat : P.Wf;
    it.valid = it.qy0 < it.H_ && it.qx0 < it.W_;
    it.y_lo = max(0, it.qy0 - md);
    const int y_hi = min(it.H_ - 1, min(it.qy0 + QTH, it.H_) - 1 + md);
    it.x_lo = max(0, it.qx0 - md);
    it.ntiles = it.valid ? (y_hi - it.y_lo) / TTH + 1 : 0;
    return it;
  };

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmF2);
    ptx::prefetch_tmap(&tmF1);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(b_full(s), 1);
      ptx::mbar_init(b_empty(s), 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(a_full(i), BUILD_WARPS);
      ptx::mbar_init(a_empty(i), 1);
      ptx::mbar_init(d_full(i), 1);
      ptx::mbar_init(d_empty(i), EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_ptr_slot)), 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_slot;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer: 64-channel chunks of the other map's tiles =====================
    uint32_t stage = 0, phase = 0;
    for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      const Item it = item_geometry(item);
      const CUtensorMap* tm = it.mode == 0 ? &tmF2 : &tmF1;
      for (int tt = 0; tt < it.ntiles; ++tt) {
        const int ty0 = it.y_lo + tt * TTH;
        for (int kb = 0; kb < P.KBC; ++kb) {
          ptx::mbar_wait(b_empty(stage), phase ^ 1);
          ptx::mbar_expect_tx(b_full(stage), B_STAGE_BYTES);
          ptx::tma_load_4d_hint(sB + stage * B_STAGE_BYTES, tm, b_full(stage), kb * CH, it.cx + P.lat * it.x_lo, it.cy + P.lat * ty0, it.bg,
                                ptx::kEvictLast);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ptx::umma_idesc(/*f16*/ 0, BM, CH) | (1u << 16);  // bit 16: B is MN-major
    uint32_t stage = 0, phase = 0;
    int icount = 0, tcount = 0;
    for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      const Item it = item_geometry(item);
      if (it.ntiles <= 0) continue;
      const uint32_t acc = icount & 1;
      ptx::mbar_wait(d_empty(acc), (((uint32_t)icount >> 1) & 1u) ^ 1u);
      ++icount;
      ptx::tc_fence_after();
      for (int tt = 0; tt < it.ntiles; ++tt, ++tcount) {
        const uint32_t abuf = tcount & 1;
        ptx::mbar_wait(a_full(abuf), ((uint32_t)tcount >> 1) & 1u);
        ptx::tc_fence_after();
        for (int kb = 0; kb < P.KBC; ++kb) {
          ptx::mbar_wait(b_full(stage), phase);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * ACC_COLS + kb * CH;
#pragma unroll
          for (int ks = 0; ks < BN / 16; ++ks) {
            const uint64_t a_desc = ptx::umma_desc_kmajor_sw128(sA + abuf * A_BYTES + (ks >> 2) * (BM * 128)) + 2 * (ks & 3);
            const uint64_t b_desc = desc_mnmajor_sw128(sB + stage * B_STAGE_BYTES + ks * 2048);
            ptx::umma_f16(d_tmem, a_desc, b_desc, idesc, (tt | ks) != 0);
          }
          ptx::umma_commit(b_empty(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(a_empty(abuf));
      }
      ptx::umma_commit(d_full(acc));
    }
  } else if (warp >= 4 && warp < 4 + BUILD_WARPS) {
    // ===================== A builders: lane = query row, 4 warps per 32-query quadrant split the tile rows =====================
    const int bw = warp - 4, quad = bw & 3, part = bw >> 2;
    const int q = quad * 32 + lane;
    const long long plane = (long long)P.Hf * P.Wf;
    int tcount = 0;
    for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      const Item it = item_geometry(item);
      if (it.ntiles <= 0) continue;
      const int oy = it.qy0 + q / QTW, ox = it.qx0 + q % QTW;
      const bool q_ok = oy < it.H_ && ox < it.W_;
      const int b = it.bg / P.G, g = it.bg - b * P.G;
      const float sd = pow2f(-pow2_exp(__ldg(P.amax + 2 * P.BG + b)));
      // window columns whose partner pixel lies inside the image
      const int b_lo = max(0, md - ox), b_hi = min(P.P - 1, it.W_ - 1 - ox + md);
      const int xoff = ox - md - it.x_lo;  // tile column of window column 0
      const float* dbase = P.dout + (long long)b * P.dout_bstride + (long long)g * P.P * P.P * plane;
      // mode 0 reads dout at this query's own pixel
      const long long pix_q = (long long)(it.cy + P.lat * oy) * P.Wf + (it.cx + P.lat * ox);
      for (int tt = 0; tt < it.ntiles; ++tt, ++tcount) {
        const int ty0 = it.y_lo + tt * TTH;
        const uint32_t abuf = tcount & 1;
        ptx::mbar_wait(a_empty(abuf), (((uint32_t)tcount >> 1) & 1u) ^ 1u);
        uint8_t* arow = base_ptr + G::SMEM_A + abuf * A_BYTES + q * 128;
        const int swz = q & 7;
        auto a_off = [&](int k) { return (uint32_t)((k >> 6) * (BM * 128) + ((((k >> 3) & 7) ^ swz) << 4) + (k & 7) * 2); };
#pragma unroll
        for (int u = 0; u < RPW; ++u) {
          const int tr = part * RPW + u;
          // zero this tile row's columns of the query's operand row (TTW * 2 bytes, 16-byte aligned)
#pragma unroll
          for (int c = 0; c < TTW / 8; ++c) *reinterpret_cast<uint4*>(arow + a_off(tr * TTW + c * 8)) = make_uint4(0, 0, 0, 0);
          const int y2 = ty0 + tr;                // partner row (lattice units)
          const int pi = y2 - oy + md;            // window row that reaches it
          if (q_ok && y2 < it.H_ && pi >= 0 && pi < P.P) {
            const float* src;
            long long pj_stride;
            if (it.mode == 0) {                   // dout[b, pi, pj, oy, ox]
              src = dbase + (long long)(pi * P.P) * plane + pix_q;
              pj_stride = plane;
            } else {                              // dout[b, P-1-pi, P-1-pj, y2, x2],  x2 = ox - md + pj
              src = dbase + (long long)((P.P - 1 - pi) * P.P + (P.P - 1)) * plane + (long long)(it.cy + P.lat * y2) * P.Wf +
                    (it.cx + P.lat * (ox - md));
              pj_stride = (long long)P.lat - plane;
            }
            uint8_t* dstrow = arow;
            const int col0 = tr * TTW + xoff;
            int pj = b_lo;
            for (; pj + 6 <= b_hi; pj += 7) {     // seven independent loads in flight (a 21-wide window row = 3 rounds)
              float v[7];
#pragma unroll
              for (int u = 0; u < 7; ++u) v[u] = __ldg(src + (pj + u) * pj_stride);
#pragma unroll
              for (int u = 0; u < 7; ++u) *reinterpret_cast<__half*>(dstrow + a_off(col0 + pj + u)) = __float2half_rn(v[u] * sd);
            }
            for (; pj + 2 <= b_hi; pj += 3) {     // three independent loads in flight
              const float v0 = __ldg(src + pj * pj_stride), v1 = __ldg(src + (pj + 1) * pj_stride), v2 = __ldg(src + (pj + 2) * pj_stride);
              *reinterpret_cast<__half*>(dstrow + a_off(col0 + pj)) = __float2half_rn(v0 * sd);
              *reinterpret_cast<__half*>(dstrow + a_off(col0 + pj + 1)) = __float2half_rn(v1 * sd);
              *reinterpret_cast<__half*>(dstrow + a_off(col0 + pj + 2)) = __float2half_rn(v2 * sd);
            }
            for (; pj <= b_hi; ++pj) *reinterpret_cast<__half*>(dstrow + a_off(col0 + pj)) = __float2half_rn(__ldg(src + pj * pj_stride) * sd);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(a_full(abuf));
      }
    }
  } else if (warp >= 4 + BUILD_WARPS) {
    // ===================== epilogue: lane = pixel, NCHW stores =====================
    const int quad = warp - (4 + BUILD_WARPS);  // warps 20..23: warp % 4 = TMEM lane quadrant
    const int q = quad * 32 + lane;
    const long long plane = (long long)P.Hf * P.Wf;
    int icount = 0;
    for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      const Item it = item_geometry(item);
      if (it.ntiles <= 0) continue;
      const int oy = it.qy0 + q / QTW, ox = it.qx0 + q % QTW;
      const bool q_ok = oy < it.H_ && ox < it.W_;
      const int b = it.bg / P.G;
      const int e_other = pow2_exp(__ldg(P.amax + (it.mode == 0 ? P.BG : 0) + it.bg));
      const float deq = P.scale * pow2f(e_other + pow2_exp(__ldg(P.amax + 2 * P.BG + b)));
      float* o = (it.mode == 0 ? P.din1 : P.din2) + (long long)it.bg * P.C * plane + (long long)(it.cy + P.lat * oy) * P.Wf + (it.cx + P.lat * ox);
      const uint32_t acc = icount & 1;
      ptx::mbar_wait(d_full(acc), ((uint32_t)icount >> 1) & 1u);
      ++icount;
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + acc * ACC_COLS + ((uint32_t)(quad * 32) << 16);
      const int ncol = P.KBC * CH;
      for (int c0 = 0; c0 < ncol; c0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32b_x32(taddr + c0, v);
        ptx::tmem_ld_wait();
        if (c0 + 32 >= ncol) {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(d_empty(acc));
        }
        if (q_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c = c0 + j;
            if (c < P.C) {
              float* p = o + (long long)c * plane;
              const float y = __uint_as_float(v[j]) * deq;
              *p = P.accumulate ? *p + y : y;
            }
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

struct Workspace {
  unsigned int* amax;  // [3][BG]
  __half* in1h;        // (BG, H, W, Cp)
  __half* in2h;
};
static size_t ws_layout(int BG, int C, int H, int W, void* basep, Workspace* ws) {
  const int Cp = round_up(C, CH);
  uintptr_t p = reinterpret_cast<uintptr_t>(basep);
  auto take = [&](size_t bytes, size_t align) {
    p = (p + align - 1) / align * align;
    uintptr_t r = p;
    p += bytes;
    return r;
  };
  const uintptr_t amax = take((size_t)3 * BG * 4, 256);
  const uintptr_t a = take((size_t)BG * H * W * Cp * 2, 1024);
  const uintptr_t b = take((size_t)BG * H * W * Cp * 2, 1024);
  if (ws) {
    ws->amax = reinterpret_cast<unsigned int*>(amax);
    ws->in1h = reinterpret_cast<__half*>(a);
    ws->in2h = reinterpret_cast<__half*>(b);
  }
  return p - reinterpret_cast<uintptr_t>(basep);
}

}  // namespace k4btc

// true when the tensor-core adjoint covers this problem
bool local_corr_bwd_tc_supported(int C, int H, int W, int Ph, int Pw, int sh, int sw, int padh, int padw, int dph, int dpw) {
  if (C > k4btc::KBC_MAX * k4btc::CH || Ph != Pw || (Ph & 1) == 0 || Ph > 25 || sh != 1 || sw != 1 || padh != 0 || padw != 0 || dph != dpw || dph < 1 || dph > 8)
    return false;
  if (dph > 1 && (cdiv(H, dph) < k4btc::QTH || cdiv(W, dph) < k4btc::QTW)) return false;  // residue classes must still fill a query tile
  return getenv("EZB_LOCAL_CORR_BWD_SIMT") == nullptr;
}

size_t local_corr_bwd_tc_workspace_bytes(int BG, int C, int H, int W) { return k4btc::ws_layout(BG, C, H, W, nullptr, nullptr) + 1024; }

// dout: first plane of this dilation's (G, P, P, H, W) block of batch entry 0; dout_bstride: elements between batch entries
int local_corr_bwd_tc_launch(const float* in1, const float* in2, const float* dout, long long dout_bstride, int BG, int G, int C, int H, int W,
                             int Pn, int dil, float scale, int accumulate, float* din1, float* din2, void* workspace, size_t workspace_bytes,
                             cudaStream_t st) {
  using namespace k4btc;
  const int Cp = round_up(C, CH), N = H * W, B = BG / G;
  uintptr_t wbase = (reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023);
  Workspace ws;
  const size_t need = ws_layout(BG, C, H, W, reinterpret_cast<void*>(wbase), &ws) + (wbase - reinterpret_cast<uintptr_t>(workspace));
  if (need > workspace_bytes) return fail(EZB_ERR_WORKSPACE, "workspace too small: need %zu, got %zu", need, workspace_bytes);
  EZB_REQUIRE(BG <= 32767, "batch too large for one launch (split it)");

  // one power-of-two scale per (batch, group) and tensor; per batch entry for dout
  if (int r = prep::launch_absmax(in1, in2, (long long)C * N, ws.amax, BG, st)) return r;
  EZB_CUDA_OK(cudaMemsetAsync(ws.amax + 2 * BG, 0, (size_t)BG * 4, st));
  {
    const long long per_b = (long long)G * Pn * Pn * N;
    // at least 16 float4 per thread; ~8 resident blocks per SM in total
    const int gx = (int)std::max<long long>(1, std::min<long long>(cdiv64(per_b, 256 * 64), std::max(1, num_sms() * 8 / B)));
    absmax_dout_kernel<<<dim3(gx, B), 256, 0, st>>>(dout, dout_bstride, per_b, ws.amax + 2 * BG);
    EZB_LAUNCH_OK();
  }
  if (int r = prep::launch_convert_kmajor2(in1, ws.in1h, in2, ws.in2h, ws.amax, BG, C, Cp, N, st)) return r;

  const int lat = dil;
  const int TTW = (QTW + Pn - 1) <= 24 ? 24 : 40, TTH = TTW == 40 ? 4 : 8;
  CUtensorMap tm1, tm2;
  for (int which = 0; which < 2; ++which) {
    cuuint64_t dims[4] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)BG};
    cuuint64_t str[3] = {(cuuint64_t)Cp * 2, (cuuint64_t)W * Cp * 2, (cuuint64_t)H * W * Cp * 2};
    cuuint32_t box[4] = {CH, (cuuint32_t)(TTW * lat), (cuuint32_t)(TTH * lat), 1};
    cuuint32_t es[4] = {1, (cuuint32_t)lat, (cuuint32_t)lat, 1};
    if (int r = make_tmap_ex(which ? &tm2 : &tm1, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, which ? ws.in2h : ws.in1h, dims, str, box, es,
                             CU_TENSOR_MAP_SWIZZLE_128B))
      return r;
  }
  Params P;
  P.G = G; P.P = Pn; P.lat = lat; P.BG = BG; P.Hf = H; P.Wf = W;
  P.Hu = cdiv(H, lat); P.Wu = cdiv(W, lat);
  P.qtx = cdiv(P.Wu, QTW); P.qty = cdiv(P.Hu, QTH); P.ncls = lat * lat; P.KBC = Cp / CH; P.C = C;
  P.modes = (din1 ? 1 : 0) | (din2 ? 2 : 0);
  P.accumulate = accumulate;
  P.dout = dout; P.dout_bstride = dout_bstride; P.amax = ws.amax; P.scale = scale; P.din1 = din1; P.din2 = din2;
  const long long n_items = (long long)P.qtx * P.qty * P.ncls * BG * (P.modes == 3 ? 2 : 1);
  EZB_REQUIRE(n_items < (1ll << 31), "output too large for one launch");
  P.n_items = (int)n_items;
  // the dout block of one (batch entry, group) and the outputs must be addressable with the pointer arithmetic above
  const int grid = (int)std::min<long long>(n_items, num_sms());
  if (TTW == 24) {
    EZB_CUDA_OK(cudaFuncSetAttribute(local_corr_bwd_tc_kernel<24, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<24, 8>::SMEM_BYTES));
    local_corr_bwd_tc_kernel<24, 8><<<grid, NUM_THREADS, Geo<24, 8>::SMEM_BYTES, st>>>(tm2, tm1, P);
  } else {
    EZB_CUDA_OK(cudaFuncSetAttribute(local_corr_bwd_tc_kernel<40, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<40, 4>::SMEM_BYTES));
    local_corr_bwd_tc_kernel<40, 4><<<grid, NUM_THREADS, Geo<40, 4>::SMEM_BYTES, st>>>(tm2, tm1, P);
  }
  EZB_LAUNCH_OK();
  return EZB_OK;
}

}  // namespace ezb
