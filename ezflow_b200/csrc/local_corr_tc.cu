// K4 / K6 on tensor cores: local displacement-window correlation as a windowed all-pairs tcgen05 GEMM.
//
// Replaces iter_spatial_correlation_sample (reference ezflow/similarity/correlation/sampler.py:29-129) and
// MatryoshkaDilatedCostVolume.forward (ezflow/similarity/learnable_cost.py:306-325) for C <= 256 per group:
//   out[b,d,g,pi,pj,y,x] = act( scale * sum_c in1p[bg,c,y*sh,x*sw] * in2p[bg,c, y*sh + pi*dph_d - mdh_d, x*sw + pj*dpw_d - mdw_d] )
//
// FlowNetC (C=256, 441 displacements) and DCVNet (C=256, 486 displacements) have 60 flop per byte of HBM
// traffic, far above what CUDA cores sustain, so the contraction runs on tcgen05:
//   prep    : ONE launch converts both tensors: fp32 (BG,C,H,W) -> fp16 K-major (BG,H,W,Cp), every image ROW scaled by its own
//             power of two (absmax of the row, found in a first sweep the second one re-reads from L2), so there is no
//             separate absmax pass over the inputs; the epilogue multiplies by (query row scale) x (target row scale)
//             (for stride > 1 / padding > 0 / lattice mode only the query pixels of in1 are gathered: (BG,oh,ow,Cp))
//   kernel  : one CTA per 8x16 tile of output pixels ("queries", UMMA M = 128).  The query operand (<= 64 KB)
//             is TMA-loaded once; negative box coordinates / OOB zero fill of the target loads implement the
//             zero padding and the image border.  The CTA then sweeps 8x16 tiles of in2 pixels ("targets", UMMA N = 128) over the
//             bounding box of all displacements of all dilations, clipped to the image, through a 4-stage TMA
//             ring; 4 TMEM accumulators decouple the MMA warp from the epilogue.
//   epilogue: lane = query.  The 128 accumulator columns of a tile go TMEM -> registers -> a private row of
//             shared memory (dynamic column indexing), then every (dilation, pi, pj) whose target falls into this
//             tile is picked up and stored: for a fixed (pi,pj) the 16 queries of a tile row write 64
//             contiguous bytes of the (B, D*G, Ph, Pw, oh, ow) output.  With several dilations only the tiles some
//             displacement reaches are visited (per-dilation row / column tables from the launcher: Params::rowtab/coltab)
//             and, for D >= 4, whole (tile, dilation) pairs are dealt to the warps of a TMEM lane quadrant.
//   zeros   : displacements that point outside the image: act(0) = 0, written from a per-lane bit mask of the out-of-image
//             columns — by warps 2-3 concurrently with the sweep when D >= 4, else by the epilogue warps before it.
//             Every output element is written exactly once.
// Only 1/5 to 1/40 of the computed products are kept (the window is a band of the all-pairs matrix), which is
// still 20-40x faster than fp32 FMAs on CUDA cores for these shapes (profiles/r01_local_corr.md).
#include "ezb_common.cuh"
#include "ezb_ptx.cuh"
#include "ezb_tma.cuh"
#include "ezb_warp.cuh"

#include <climits>
#include <cstring>
#include <type_traits>
#include <cstdlib>

namespace ezb {
namespace k4tc {

constexpr int QTH = 8, QTW = 16, BM = QTH * QTW;  // query tile  -> UMMA M = 128 (TMEM lanes)
constexpr int BK = 64, KB_MAX = 4;                // C <= 256
constexpr int A_KB_BYTES = BM * BK * 2;           // 16 KB
constexpr int EPI_SPLIT = 4;                      // epilogue warps per TMEM lane quadrant
constexpr int EPI_WARPS = 4 * EPI_SPLIT;
constexpr int NUM_THREADS = 128 + EPI_WARPS * 32;  // warp 0 TMA, 1 MMA, 2 TMEM alloc, 2-3 zero fill (D >= 4), 4.. epilogue
constexpr int MAX_DIL = 8;  // also the bit width of the sparse-sweep tables (one bit per dilation)
constexpr int TMEM_COLS = 512;
constexpr int SWEEP_TAB = 288;  // entries of the sparse-sweep tables: 2 * max margin + query tile extent must fit

// Target-tile geometry (rows x columns of in2 pixels = UMMA N).  8x16 (N = 128) is the general tile.  When the x-extent of a
// query tile's bounding box is small, the tile spans the whole box width instead, so that no displacement straddles two
// tiles in x and fewer extraction slots are wasted: 8x24 (N = 192; PWC-Net's 9x9 window), 4x40 (N = 160; FlowNetC's 21x21
// window in lattice mode).
template <int TTW_, int TTH_>
struct Geo {
  static constexpr int TTW = TTW_, TTH = TTH_, BN = TTH * TTW;
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGES = BN <= 128 ? 4 : BN <= 160 ? 3 : 2;
  static constexpr int ACC_STRIDE = BN <= 128 ? 128 : 256;  // TMEM columns between accumulators
  static constexpr int NACC = TMEM_COLS / ACC_STRIDE;
  static constexpr int S_STRIDE = BN + 1;  // words; odd => lane = query accesses are conflict-free
  static constexpr int CPW = BN / EPI_SPLIT;  // accumulator columns each epilogue warp stages
  static constexpr int SMEM_A = 0;
  static constexpr int SMEM_B = SMEM_A + KB_MAX * A_KB_BYTES;
  static constexpr int SMEM_S = SMEM_B + STAGES * B_STAGE_BYTES;
  static constexpr int SMEM_TS = SMEM_S + BM * S_STRIDE * 4;  // [4 quadrants][BN] power-of-two scales of the tile's target pixels
  static constexpr int SMEM_BAR = SMEM_TS + 4 * BN * 4;
  static constexpr int SMEM_BYTES = SMEM_BAR + 256 + 1024;
  static_assert(CPW % 8 == 0 && CPW <= 56 && BN % 16 == 0 && SMEM_BYTES <= 232448 && (2 * STAGES + 4 + 2 * NACC) * 8 + 4 <= 256, "tile geometry");
};

// n / d for 0 <= n < 2^31 as a multiply-high and a shift (the divisors are launch constants: the item decode would
// otherwise cost every warp half a dozen integer divisions per work item)
struct FastDiv {
  unsigned mul, shr;
  int d;
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f{0u, 0u, d};
  if (d > 1) {
    int lg = 31 - __builtin_clz((unsigned)d);
    if (d & (d - 1)) ++lg;  // ceil(log2 d)
    const int p = 31 + lg;
    f.mul = (unsigned)(((1ull << p) + (unsigned)d - 1) / (unsigned)d);
    f.shr = (unsigned)(p - 32);
  }
  return f;
}
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) { return f.d == 1 ? n : (int)(__umulhi((unsigned)n, f.mul) >> f.shr); }

struct Params {
  int G, H, W, oh, ow, Ph, Pw, sh, sw, padh, padw, D, KB;
  FastDiv d_qtx, d_qty, d_nsplit, d_ncls, d_lat, d_G;
  int regular;  // D == 1, dilation 1 (layout units), stride 1, no padding, window x-extent inside ONE target tile: row-wise extraction
  int nsplit;  // work items sharing one query tile (each sweeps a slice of its target tiles): balances the load
  // Persistent CTAs: work item = (query tile x, query tile y, z = (bg * lat^2 + class) * nsplit + split), item index
  // x-fastest; CTA c takes items c, c + gridDim.x, ...
  int qtx, qty, n_items;
  // Lattice mode (lat > 1; one dilation = lat, stride 1, no padding): queries and targets with the same residues
  // (cy, cx) mod lat only ever meet each other, so every residue class is an independent dilation-1 problem on the
  // sub-lattice  pixel = (cy, cx) + lat * (y, x);  H, W, oh, ow, dph, dpw below are then in lattice units of class (0,0).
  int lat, BG;
  int Hf, Wf;        // full-resolution input size (lattice mode)
  int out_h, out_w;  // full output plane
  int dph[MAX_DIL], dpw[MAX_DIL];
  int mh[MAX_DIL], mw[MAX_DIL];             // dph * (Ph - 1) / 2, dpw * (Pw - 1) / 2
  float inv_dph[MAX_DIL], inv_dpw[MAX_DIL];  // 1 / dph, 1 / dpw (round to nearest)
  float slope;
  const float* qscale;  // [BG * lat^2][oh * ow] 2^e of every query pixel: its fp16 operand row was scaled by 2^-e
  const float* tscale;  // [BG][Hf * Wf]         2^e of every in2 pixel
  float scale;          // user scale (CorrelationLayer: 1/C)
  float* out;
  long long out_bstride;
  // Sparse sweep (general extraction): a target tile of the bounding box is visited only if some displacement of some
  // dilation takes a query of the tile into it.  Separable per dilation: bit d of rowtab[ty0 - query-box y0 + max margin]
  // and of coltab[tx0 - query-box x0 + max margin].  DCVNet's {1,2,3,5,9,16} needs 123 of the 153 tiles of its box.
  int sparse;
  int zero_idle;     // 1: the out-of-image zeros are written by the two idle warps, 0: by the epilogue warps before the sweep
  int xclip_unroll;  // regular case, window clipped in x: warp-uniform unrolled walk (1) or the per-lane loop (0)
  unsigned char rowtab[SWEEP_TAB], coltab[SWEEP_TAB];
};

// ---------------------------------------------------------------------------------------------------------------
// Operand conversion: fp32 NCHW -> fp16 K-major with a power-of-two scale per PIXEL (= per operand row of the GEMM).
//   dst[img, oy, ox, c] = 2^-e * value(c, oy, ox),   rowscale[img * oh * ow + oy * ow + ox] = 2^e,  e = floor(log2 max_c |value|)
//   value = src[bg, c, oy*sh - padh, ox*sw - padw]  (0 outside the image or for c >= C), or, with a flow field, the
//   bilinear warp of K5 (ezflow/utils/warp.py:5-42) evaluated at that pixel — the warped fp32 tensor is never written.
// In lattice mode (queries of a dilation-lat problem) img = bg * lat^2 + class and the class (cy, cx) acts as a negative
// padding with stride lat.  One pass over the inputs: no absmax kernel, no second read.  Up to two jobs (targets, queries)
// share one launch (blockIdx.z ranges).
struct ConvJob {
  const float* src;
  const float* flow;  // (B,2,H,W) or null
  __half* dst;
  float* rowscale;  // one power-of-two scale per destination pixel
  int C, Cp, H, W, oh, ow, sh, sw, padh, padw, lat, nimg, G;
};
struct ConvJobs {
  ConvJob j[2];
  int groups;  // 32-pixel groups a block converts at once
};

constexpr int CONV_THREADS = 256;
constexpr int CONV_MAX_PX = 256;  // destination pixels per block at most (few channels: more pixels, same bytes)

// Block = `npx` consecutive destination pixels of one image x all channels, npx = 32 * groups (JJ.groups: 1 at Cp = 256, 4 at
// Cp = 64, ...), so that a block moves ~32 KB whatever the channel count.  Thread -> (pixel, channel slice): one sweep loads
// the values (coalesced along x, up to 8 loads in flight per thread) into shared memory [Cp][npx + 1] while tracking every pixel's
// absmax, then a warp per pixel writes the K-major row scaled by the pixel's own power of two.
__global__ void __launch_bounds__(CONV_THREADS) convert_pixels_kmajor_kernel(const ConvJobs JJ) {
  extern __shared__ float pxbuf[];  // [Cp][npx + 1]
  __shared__ float red[CONV_THREADS];
  __shared__ float s_scale[CONV_MAX_PX];
  const bool second = (int)blockIdx.z >= JJ.j[0].nimg;
  const ConvJob& J = JJ.j[second ? 1 : 0];
  const int img = (int)blockIdx.z - (second ? JJ.j[0].nimg : 0);
  const int total_px = J.oh * J.ow;
  const int npx = 32 * JJ.groups, pitch = npx + 1, nsl = CONV_THREADS / npx;  // pixels per block, channel slices
  const int n0 = blockIdx.x * npx;
  if (n0 >= total_px) return;
  int sh = J.sh, sw = J.sw, padh = J.padh, padw = J.padw, bg = img;
  if (J.lat > 1) {
    const int ncls = J.lat * J.lat, cls = img % ncls;
    bg = img / ncls;
    sh = sw = J.lat;
    padh = -(cls / J.lat);
    padw = -(cls % J.lat);
  }
  const int HW = J.H * J.W;
  const float* src = J.src + (long long)bg * J.C * HW;
  const float* fl = J.flow ? J.flow + (long long)(bg / J.G) * 2 * HW : nullptr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = CONV_THREADS / 32;

  // ---- sweep 1: thread = (pixel, channel slice); consecutive threads take consecutive pixels
  const int pl = threadIdx.x % npx, sl = threadIdx.x / npx;
  const int n = n0 + pl;
  const int oy = n / J.ow, ox = n - oy * J.ow;
  const int y = oy * sh - padh, x = ox * sw - padw;
  const bool in_img = n < total_px && y >= 0 && y < J.H && x >= 0 && x < J.W;
  float m = 0.f;
  if (fl) {
    k5::Gather g;
    g.zero = true;
    if (in_img) {
      const k5::Taps t = k5::make_taps(__ldg(fl + y * J.W + x), __ldg(fl + HW + y * J.W + x), x, y, J.H, J.W);
      g = k5::make_gather(t, J.W);
    }
    for (int c = sl; c < J.Cp; c += nsl) {
      float v = 0.f;
      if (c < J.C && !g.zero) {
        const float* sp = src + (long long)c * HW;
        v = (g.w00 * __ldg(sp + g.oaa) + g.w01 * __ldg(sp + g.oab)) + (g.w10 * __ldg(sp + g.oba) + g.w11 * __ldg(sp + g.obb));
      }
      m = fmaxf(m, fabsf(v));
      pxbuf[c * pitch + pl] = v;
    }
  } else {
    const float* sp = src + (long long)y * J.W + x;
    int c = sl;
    constexpr int LD = 16;  // independent loads in flight per thread (channels >= C are zero-filled)
    for (; c + (LD - 1) * nsl < J.Cp; c += LD * nsl) {
      float a[LD];
#pragma unroll
      for (int u = 0; u < LD; ++u) a[u] = (in_img && c + u * nsl < J.C) ? __ldg(sp + (long long)(c + u * nsl) * HW) : 0.f;
#pragma unroll
      for (int u = 0; u < LD; ++u) {
        m = fmaxf(m, fabsf(a[u]));
        pxbuf[(c + u * nsl) * pitch + pl] = a[u];
      }
    }
    for (; c < J.Cp; c += nsl) {
      const float v = (in_img && c < J.C) ? __ldg(sp + (long long)c * HW) : 0.f;
      m = fmaxf(m, fabsf(v));
      pxbuf[c * pitch + pl] = v;
    }
  }
  red[threadIdx.x] = m;  // [slice][pixel]
  __syncthreads();
  if ((int)threadIdx.x < npx) {
    float mm = red[threadIdx.x];
    for (int i = 1; i < nsl; ++i) mm = fmaxf(mm, red[i * npx + threadIdx.x]);
    const int eb = (int)((__float_as_uint(mm) >> 23) & 0xff);
    const int e = (eb == 0 || eb == 255) ? 0 : max(-126, min(126, eb - 127));  // 0 for an all-zero pixel
    s_scale[threadIdx.x] = __uint_as_float((unsigned int)(127 - e) << 23);
    if (n0 + (int)threadIdx.x < total_px) J.rowscale[(long long)img * total_px + n0 + threadIdx.x] = __uint_as_float((unsigned int)(127 + e) << 23);
  }
  __syncthreads();

  // ---- sweep 2: a warp per pixel, a lane per channel pair of a 64-channel chunk (128-byte stores)
  __half* dimg = J.dst + (long long)img * total_px * J.Cp;
  for (int p = warp; p < npx && n0 + p < total_px; p += NW) {
    const float sc = s_scale[p];
    for (int c0 = 0; c0 < J.Cp; c0 += 64) {
      const int c = c0 + 2 * lane;
      *reinterpret_cast<__half2*>(dimg + (long long)(n0 + p) * J.Cp + c) = __floats2half2_rn(pxbuf[c * pitch + p] * sc, pxbuf[(c + 1) * pitch + p] * sc);
    }
  }
}

// floor(a / b) and ceil(a / b) for small |a| and b > 0 through a reciprocal: (a + 0.5) / b is never within 0.5/b >= 2^-5 of an
// integer, far more than the rounding error of the float product, so the floor is exact (|a| < 2^15)
__device__ __forceinline__ int floor_div(int a, float inv_b) { return __float2int_rd(((float)a + 0.5f) * inv_b); }
__device__ __forceinline__ int ceil_div(int a, float inv_b) { return -floor_div(-a, inv_b); }

// bit nb set: a row of nb window columns wastes fewer slots in chunks of 3 than in chunks of 4
constexpr unsigned ch3_mask() {
  unsigned m = 0;
  for (int nb = 1; nb < 32; ++nb)
    if ((nb + 2) / 3 * 3 < (nb + 3) / 4 * 4) m |= 1u << nb;
  return m;
}
constexpr unsigned kCh3Mask = ch3_mask();

template <int TTW_, int TTH_>
__global__ void __launch_bounds__(NUM_THREADS, 1)
    local_corr_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Params P) {
  using G = Geo<TTW_, TTH_>;
  constexpr int TTW = G::TTW, TTH = G::TTH, BN = G::BN, STAGES = G::STAGES, NACC = G::NACC, B_STAGE_BYTES = G::B_STAGE_BYTES;
  constexpr int S_STRIDE = G::S_STRIDE, CPW = G::CPW, ACC_STRIDE = G::ACC_STRIDE;
  constexpr int SMEM_A = G::SMEM_A, SMEM_B = G::SMEM_B, SMEM_S = G::SMEM_S, SMEM_BAR = G::SMEM_BAR;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw_addr);
  const uint32_t sA = base + SMEM_A, sB = base + SMEM_B, sBar = base + SMEM_BAR;
  float* S = reinterpret_cast<float*>(base_ptr + SMEM_S);
  float* TS = reinterpret_cast<float*>(base_ptr + G::SMEM_TS);
  auto full_bar = [&](int s) { return sBar + 8u * s; };
  auto empty_bar = [&](int s) { return sBar + 8u * (STAGES + s); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // the query operand is double-buffered when it takes at most half of its shared-memory area (C <= 128)
  const int na = P.KB <= KB_MAX / 2 ? 2 : 1;
  auto a_full = [&](int s) { return sBar + 8u * (2 * STAGES + s); };
  auto a_empty = [&](int s) { return sBar + 8u * (2 * STAGES + 2 + s); };
  auto tmem_full = [&](int a) { return sBar + 8u * (2 * STAGES + 4 + a); };
  auto tmem_empty = [&](int a) { return sBar + 8u * (2 * STAGES + 4 + NACC + a); };
  volatile uint32_t* tmem_ptr_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + SMEM_BAR + 8 * (2 * STAGES + 4 + 2 * NACC));

  int mdh_max = 0, mdw_max = 0;
  for (int d = 0; d < P.D; ++d) {
    mdh_max = max(mdh_max, P.dph[d] * (P.Ph - 1) / 2);
    mdw_max = max(mdw_max, P.dpw[d] * (P.Pw - 1) / 2);
  }
  const int ncls = P.lat * P.lat;

  // geometry of a work item (every role walks the same item sequence and derives the same numbers)
  struct Item {
    int qx0, qy0, zq, split, bg, cy, cx, H_, W_, oh_, ow_, y_lo, x_lo, ntx, tt_begin, ntiles;
    bool valid;
  };
  auto item_geometry = [&](int item) {
    Item it;
    const int rest = fdiv(item, P.d_qtx), bx = item - rest * P.qtx, bz = fdiv(rest, P.d_qty), by = rest - bz * P.qty;
    it.qx0 = bx * QTW;
    it.qy0 = by * QTH;
    it.zq = fdiv(bz, P.d_nsplit);  // bg * lat^2 + class
    it.split = bz - it.zq * P.nsplit;
    it.bg = fdiv(it.zq, P.d_ncls);
    const int cls = it.zq - it.bg * ncls;
    it.cy = fdiv(cls, P.d_lat);
    it.cx = cls - it.cy * P.lat;
    // image / output size in the units the tiles are laid out in (lattice mode: the lattice of this residue class)
    it.H_ = P.lat > 1 ? fdiv(P.Hf - it.cy + P.lat - 1, P.d_lat) : P.H;
    it.W_ = P.lat > 1 ? fdiv(P.Wf - it.cx + P.lat - 1, P.d_lat) : P.W;
    it.oh_ = P.lat > 1 ? it.H_ : P.oh;
    it.ow_ = P.lat > 1 ? it.W_ : P.ow;
    it.valid = it.qy0 < it.oh_ && it.qx0 < it.ow_;  // (lattice mode) a class smaller than class (0,0) has fewer tiles
    // bounding box (in2 pixel coordinates, clipped to the image) of every target any query of this tile needs
    it.y_lo = max(0, it.qy0 * P.sh - P.padh - mdh_max);
    const int y_hi = min(it.H_ - 1, (min(it.qy0 + QTH, it.oh_) - 1) * P.sh - P.padh + mdh_max);
    it.x_lo = max(0, it.qx0 * P.sw - P.padw - mdw_max);
    const int x_hi = min(it.W_ - 1, (min(it.qx0 + QTW, it.ow_) - 1) * P.sw - P.padw + mdw_max);
    const int nty = y_hi >= it.y_lo ? (y_hi - it.y_lo) / TTH + 1 : 0;
    it.ntx = x_hi >= it.x_lo ? (x_hi - it.x_lo) / TTW + 1 : 0;
    const int ntiles_all = it.valid ? nty * it.ntx : 0;
    it.tt_begin = fdiv(ntiles_all * it.split, P.d_nsplit);
    it.ntiles = fdiv(ntiles_all * (it.split + 1), P.d_nsplit) - it.tt_begin;  // this item sweeps target tiles [tt_begin, tt_begin + ntiles)
    return it;
  };

  // CTA-uniform, identical in every role: does any displacement reach the target tile at (ty0, tx0) from this query tile?
  auto tile_needed = [&](const Item& it, int ty0, int tx0) {
    if (!P.sparse) return true;
    const int ky = ty0 - (it.qy0 * P.sh - P.padh) + mdh_max, kx = tx0 - (it.qx0 * P.sw - P.padw) + mdw_max;
    return (P.rowtab[ky] & P.coltab[kx]) != 0;
  };

  // Displacements that point outside the image: act(0) = 0, written once.  A query tile whose whole displacement box lies
  // inside the image has none.  The calling warp(s) cover queries q0, q0 + qstep, ... < BM and window rows a0, a0 + astep, ...;
  // the loops over (pi, pj) are warp-uniform with a per-lane bit mask of the out-of-image columns, so a store instruction
  // still covers contiguous output pixels.  All 32 lanes of a warp must call it together.
  auto zero_fill = [&](const Item& it, int q0, int qstep, int a0, int astep) {
    if (!it.valid) return;  // the items sharing a query tile (nsplit) share its zeros too: the callers deal them window rows
    const int qx0 = it.qx0, qy0 = it.qy0, H_ = it.H_, W_ = it.W_;
    const bool box_inside = qy0 * P.sh - P.padh - mdh_max >= 0 && (min(qy0 + QTH, it.oh_) - 1) * P.sh - P.padh + mdh_max < H_ &&
                            qx0 * P.sw - P.padw - mdw_max >= 0 && (min(qx0 + QTW, it.ow_) - 1) * P.sw - P.padw + mdw_max < W_;
    if (box_inside) return;
    const unsigned uplane = (unsigned)(P.out_h * P.out_w);
    const int b = fdiv(it.bg, P.d_G), g = it.bg - b * P.G;
    for (int q = q0; q < BM; q += qstep) {
      const int oy = qy0 + q / QTW, ox = qx0 + q % QTW;
      const bool q_ok = oy < it.oh_ && ox < it.ow_;
      float* out_q = P.out + (long long)b * P.out_bstride + (long long)(it.cy + P.lat * oy) * P.out_w + (it.cx + P.lat * ox);
      for (int d = 0; d < P.D; ++d) {
        const int dph = P.dph[d], dpw = P.dpw[d];
        const int y2_0 = oy * P.sh - P.padh - P.mh[d], x2_0 = ox * P.sw - P.padw - P.mw[d];
        const unsigned o_d = (unsigned)((d * P.G + g) * P.Ph * P.Pw) * uplane;  // offsets inside a batch entry fit 32 bits (launcher)
        for (int bb0 = 0; bb0 < P.Pw; bb0 += 32) {  // 32 window columns at a time (one pass for every window <= 31 wide)
          const int nb = min(32, P.Pw - bb0);
          unsigned zx = 0;  // columns of this lane's window that fall outside the image
          for (int u = 0; u < nb; ++u) {
            const int x2 = x2_0 + (bb0 + u) * dpw;
            zx |= (unsigned)(x2 < 0 || x2 >= W_) << u;
          }
          const unsigned all = nb == 32 ? 0xffffffffu : (1u << nb) - 1u;
          for (int a = a0; a < P.Ph; a += astep) {
            const int y2 = y2_0 + a * dph;
            const unsigned m = !q_ok ? 0u : (y2 >= 0 && y2 < H_) ? zx : all;
            if (!__any_sync(0xffffffffu, m != 0u)) continue;
            unsigned off = o_d + (unsigned)(a * P.Pw + bb0) * uplane;
            for (int u = 0; u < nb; ++u, off += uplane)
              if ((m >> u) & 1u) out_q[off] = 0.f;
          }
        }
      }
    }
  };

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(a_full(s), 1);
      ptx::mbar_init(a_empty(s), 1);
    }
    for (int a = 0; a < NACC; ++a) {
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
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_slot;
  const uint32_t a_slot_bytes = (uint32_t)(KB_MAX / 2) * A_KB_BYTES;

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    uint32_t stage = 0, phase = 0;
    int acount = 0;  // items with a query operand so far
    for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      const Item it = item_geometry(item);
      if (it.ntiles <= 0) continue;
      {
        const int slot = acount % na;
        const uint32_t aph = (uint32_t)(acount / na) & 1u;
        ++acount;
        ptx::mbar_wait(a_empty(slot), aph ^ 1);  // the MMAs of the item that used this slot before have completed
        ptx::mbar_expect_tx(a_full(slot), (uint32_t)(P.KB * A_KB_BYTES));
        for (int kb = 0; kb < P.KB; ++kb)
          ptx::tma_load_4d_hint(sA + slot * a_slot_bytes + kb * A_KB_BYTES, &tmA, a_full(slot), kb * BK, it.qx0, it.qy0, it.zq,
                                ptx::kEvictFirst);  // (oy, ox) space
      }
      int tyi = it.tt_begin / it.ntx, txi = it.tt_begin - tyi * it.ntx;  // tile row / column inside the bounding box
      for (int tt = 0; tt < it.ntiles; ++tt) {
        const int ty0 = it.y_lo + tyi * TTH, tx0 = it.x_lo + txi * TTW;
        if (++txi == it.ntx) { txi = 0; ++tyi; }
        if (!tile_needed(it, ty0, tx0)) continue;
        for (int kb = 0; kb < P.KB; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1);
          ptx::mbar_expect_tx(full_bar(stage), B_STAGE_BYTES);
          // in2 windows of neighbouring query tiles overlap heavily: keep them in L2
          ptx::tma_load_4d_hint(sB + stage * B_STAGE_BYTES, &tmB, full_bar(stage), kb * BK, it.cx + P.lat * tx0, it.cy + P.lat * ty0,
                                it.bg, ptx::kEvictLast);  // lattice -> pixel coordinates
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ptx::umma_idesc(/*f16*/ 0, BM, BN);
    uint32_t stage = 0, phase = 0;
    int acount = 0, tcount = 0;  // running counts of query operands / target tiles (accumulators rotate across items)
    for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      const Item it = item_geometry(item);
      if (it.ntiles <= 0) continue;
      const int slot = acount % na;
      const uint32_t aph = (uint32_t)(acount / na) & 1u;
      ++acount;
      ptx::mbar_wait(a_full(slot), aph);
      ptx::tc_fence_after();
      int tyi = it.tt_begin / it.ntx, txi = it.tt_begin - tyi * it.ntx;
      for (int tt = 0; tt < it.ntiles; ++tt) {
        const int ty0 = it.y_lo + tyi * TTH, tx0 = it.x_lo + txi * TTW;
        if (++txi == it.ntx) { txi = 0; ++tyi; }
        if (!tile_needed(it, ty0, tx0)) continue;
        const uint32_t acc = tcount % NACC, aphase = (tcount / NACC) & 1;
        ++tcount;
        ptx::mbar_wait(tmem_empty(acc), aphase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
        for (int kb = 0; kb < P.KB; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint64_t a_desc = ptx::umma_desc_kmajor_sw128(sA + slot * a_slot_bytes + kb * A_KB_BYTES);
          const uint64_t b_desc = ptx::umma_desc_kmajor_sw128(sB + stage * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) ptx::umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          ptx::umma_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(tmem_full(acc));
      }
      ptx::umma_commit(a_empty(slot));  // arrives once every MMA that read this query operand is done
    }
  } else if (warp == 2 || warp == 3) {
    // ===================== zero fill on the two otherwise idle warps (P.zero_idle) =====================
    // Concurrent with the sweep: the epilogue only writes in-image targets, so the two never touch the same element.
    if (P.zero_idle) {
      const int t = threadIdx.x - 64;
      for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
        const Item it = item_geometry(item);
        zero_fill(it, t, 64, it.split, P.nsplit);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: lane = query =====================
    // EPI_SPLIT warps share a TMEM lane quadrant (= 32 queries): each drains its share of the accumulator columns
    // into the queries' shared-memory rows, then they split the (pi, pj-chunk) work items between them.
    const int quad = warp & 3, part = (warp - 4) >> 2;
    const int q = quad * 32 + lane;
    const long long plane = (long long)P.out_h * P.out_w;
    float* Srow = S + q * S_STRIDE;
    float* tsq = TS + quad * BN;  // this quadrant's target-pixel scales of the current tile
    auto quad_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + quad), "n"(32 * EPI_SPLIT) : "memory"); };
    int tcount = 0;
    for (int item = blockIdx.x; item < P.n_items; item += gridDim.x) {
      const Item it = item_geometry(item);
      if (!it.valid) continue;
      const int qx0 = it.qx0, qy0 = it.qy0, H_ = it.H_, W_ = it.W_;
      const int oy = qy0 + q / QTW, ox = qx0 + q % QTW;
      const bool q_ok = oy < it.oh_ && ox < it.ow_;
      const int b = fdiv(it.bg, P.d_G), g = it.bg - b * P.G;
      // power-of-two factors of the operand rows: the query row's is folded into deq, the target rows' into the drain below
      const float deq = P.scale * (q_ok ? __ldg(P.qscale + ((long long)it.zq * P.oh + oy) * P.ow + ox) : 0.f);
      const float* tpix_scale = P.tscale + (long long)it.bg * P.Hf * P.Wf;
      // offsets inside one batch entry of the output fit 32 bits (checked by the launcher): one IMAD per store address
      float* out_q = P.out + (long long)b * P.out_bstride + (long long)(it.cy + P.lat * oy) * P.out_w + (it.cx + P.lat * ox);
      const unsigned uplane = (unsigned)plane;

      // this lane's query; window rows dealt over the EPI_SPLIT warps of the quadrant and the nsplit items of the query tile
      if (!P.zero_idle) zero_fill(it, q, BM, part + EPI_SPLIT * it.split, EPI_SPLIT * P.nsplit);

      // general case: this lane's undisplaced target and the query tile's undisplaced box (CTA-uniform), once per item
      const int yq = oy * P.sh - P.padh, xq = ox * P.sw - P.padw;
      const int qbox_y0 = qy0 * P.sh - P.padh, qbox_y1 = (qy0 + QTH - 1) * P.sh - P.padh;
      const int qbox_x0 = qx0 * P.sw - P.padw, qbox_x1 = (qx0 + QTW - 1) * P.sw - P.padw;

      int tyi = it.ntiles > 0 ? it.tt_begin / it.ntx : 0, txi = it.tt_begin - tyi * it.ntx;
      for (int tt = 0; tt < it.ntiles; ++tt) {
        const int ty0 = it.y_lo + tyi * TTH, tx0 = it.x_lo + txi * TTW;
        if (++txi == it.ntx) { txi = 0; ++tyi; }
        if (!tile_needed(it, ty0, tx0)) continue;
        const int ty1 = min(ty0 + TTH, H_) - 1, tx1 = min(tx0 + TTW, W_) - 1;  // last in-image row / column of the tile
        const uint32_t acc = tcount % NACC, aphase = (tcount / NACC) & 1;
        ++tcount;
        ptx::mbar_wait(tmem_full(acc), aphase);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + acc * ACC_STRIDE + ((uint32_t)(quad * 32) << 16) + part * CPW;
        {
          // CPW = 32 (+16) (+8) columns of this warp's share
          constexpr bool k32 = CPW >= 32, k16 = (CPW % 32) >= 16, k8 = (CPW % 16) >= 8;
          uint32_t r0[32], r1[16], r2[8];
          if constexpr (k32) ptx::tmem_ld_32x32b_x32(taddr, r0);
          if constexpr (k16) ptx::tmem_ld_32x32b_x16(taddr + (k32 ? 32 : 0), r1);
          if constexpr (k8) ptx::tmem_ld_32x32b_x8(taddr + (k32 ? 32 : 0) + (k16 ? 16 : 0), r2);
          // power-of-two scales of the target pixels behind this warp's CPW accumulator columns -> the quadrant's scale row
          // (applied when an element is extracted: far fewer elements are extracted than drained)
          {
            auto col_scale = [&](int col) {
              const int tyr = col / TTW, txr = col - tyr * TTW;
              const int yy = it.cy + P.lat * (ty0 + tyr), xx = it.cx + P.lat * (tx0 + txr);
              return (yy < P.Hf && xx < P.Wf) ? __ldg(tpix_scale + (long long)yy * P.Wf + xx) : 0.f;
            };
            if (lane < CPW) tsq[part * CPW + lane] = col_scale(part * CPW + lane);
            if constexpr (CPW > 32) {
              if (lane + 32 < CPW) tsq[part * CPW + 32 + lane] = col_scale(part * CPW + 32 + lane);
            }
          }
          ptx::tmem_ld_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(tmem_empty(acc));  // the accumulator is in registers: release it early
          float* dst = Srow + part * CPW;
          if constexpr (k32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[j] = __uint_as_float(r0[j]);
          }
          if constexpr (k16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) dst[(k32 ? 32 : 0) + j] = __uint_as_float(r1[j]);
          }
          if constexpr (k8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[(k32 ? 32 : 0) + (k16 ? 16 : 0) + j] = __uint_as_float(r2[j]);
          }
        }
        // regular case: a warp extracts exactly the tile rows it drained itself, so the warps of a quadrant never wait for
        // each other and drift apart over the accumulators (one tile's extraction overlaps the next tile's drain)
        if (P.regular == 1) __syncwarp(); else quad_sync();  // all column shares of the rows are in shared memory

        if (P.regular) {
          // Regular case (PWC-Net, CorrelationLayer, FlowNetC in lattice mode): a tile row tr holds, for every query of this
          // warp, the window row pi = ty0 + tr - (oy - md) as Pw consecutive accumulator columns starting at xoff.  The
          // warps of a quadrant take tile rows round-robin; a store instruction covers one (pi, pj): 2 x 64 contiguous bytes.
          const int mdh = (P.Ph - 1) >> 1, mdw = (P.Pw - 1) >> 1;
          const int xoff = ox - mdw - tx0;
          const int b_lo = max(0, mdw - ox), b_hi = min(P.Pw - 1, W_ - 1 - ox + mdw);  // window columns inside the image
          const bool x_inside = qx0 - mdw >= 0 && qx0 + QTW - 1 + mdw < W_ && qx0 + QTW <= it.ow_;  // CTA-uniform: no clipping in x
          // union of the warp's column ranges (clipped case)
          const int ub_lo = x_inside ? 0 : __reduce_min_sync(0xffffffffu, q_ok ? b_lo : INT_MAX);
          const int ub_hi = x_inside ? P.Pw - 1 : __reduce_max_sync(0xffffffffu, q_ok ? b_hi : INT_MIN);
          const unsigned o_d = (unsigned)(g * P.Ph * P.Pw) * uplane;
          auto rows = [&](auto act_tag) {
            constexpr bool ACT = decltype(act_tag)::value;
            constexpr int RPW = CPW / TTW;  // tile rows per epilogue warp: its own accumulator columns
            static_assert(CPW % TTW == 0, "an epilogue warp's columns are whole target-tile rows");
            for (int tr = part * RPW; tr < (part + 1) * RPW && ty0 + tr <= ty1; ++tr) {
              const int pi = ty0 + tr - oy + mdh;
              if (!(q_ok && pi >= 0 && pi < P.Ph)) continue;
              const float* srow = Srow + tr * TTW + xoff;
              const float* trow = tsq + tr * TTW + xoff;
              float* o = out_q + (o_d + (unsigned)(pi * P.Pw) * uplane);
              if (x_inside) {
                int pj = 0;
                for (; pj + 3 <= P.Pw; pj += 3) {  // three independent load -> store chains in flight
                  float w0 = srow[pj] * (trow[pj] * deq), w1 = srow[pj + 1] * (trow[pj + 1] * deq), w2 = srow[pj + 2] * (trow[pj + 2] * deq);
                  if (ACT) {
                    w0 = w0 >= 0.f ? w0 : w0 * P.slope;
                    w1 = w1 >= 0.f ? w1 : w1 * P.slope;
                    w2 = w2 >= 0.f ? w2 : w2 * P.slope;
                  }
                  o[0] = w0;
                  o[uplane] = w1;
                  o[2 * uplane] = w2;
                  o += 3 * uplane;
                }
                for (; pj < P.Pw; ++pj) {
                  float w = srow[pj] * (trow[pj] * deq);
                  if (ACT) w = w >= 0.f ? w : w * P.slope;
                  *o = w;
                  o += uplane;
                }
              } else if (!P.xclip_unroll) {
                for (int pj = b_lo; pj <= b_hi; ++pj) {
                  float w = srow[pj] * (trow[pj] * deq);
                  if (ACT) w = w >= 0.f ? w : w * P.slope;
                  o[(unsigned)pj * uplane] = w;
                }
              } else {
                // clipped in x (every tile of FlowNetC's 32-pixel-wide lattice): the warp walks the union of its lanes' column
                // ranges, three independent load -> store chains in flight, a lane keeps what lies inside its own range
                float* oo = o + (unsigned)ub_lo * uplane;
                int pj = ub_lo;
                for (; pj + 2 <= ub_hi; pj += 3) {
                  const bool k0 = pj >= b_lo && pj <= b_hi, k1 = pj + 1 >= b_lo && pj + 1 <= b_hi, k2 = pj + 2 >= b_lo && pj + 2 <= b_hi;
                  float w0 = k0 ? srow[pj] * (trow[pj] * deq) : 0.f;
                  float w1 = k1 ? srow[pj + 1] * (trow[pj + 1] * deq) : 0.f;
                  float w2 = k2 ? srow[pj + 2] * (trow[pj + 2] * deq) : 0.f;
                  if (ACT) {
                    w0 = w0 >= 0.f ? w0 : w0 * P.slope;
                    w1 = w1 >= 0.f ? w1 : w1 * P.slope;
                    w2 = w2 >= 0.f ? w2 : w2 * P.slope;
                  }
                  if (k0) oo[0] = w0;
                  if (k1) oo[uplane] = w1;
                  if (k2) oo[2 * uplane] = w2;
                  oo += 3 * uplane;
                }
                for (; pj <= ub_hi; ++pj) {
                  if (pj >= b_lo && pj <= b_hi) {
                    float w = srow[pj] * (trow[pj] * deq);
                    if (ACT) w = w >= 0.f ? w : w * P.slope;
                    *oo = w;
                  }
                  oo += uplane;
                }
              }
            }
          };
          if (P.slope != 1.f) rows(std::true_type{});
          else rows(std::false_type{});
        } else {
        // sparse sweep: the dilations with a displacement into this tile, straight from the tables (CTA-uniform)
        const unsigned dmask = P.sparse ? (unsigned)(P.rowtab[ty0 - qbox_y0 + mdh_max] & P.coltab[tx0 - qbox_x0 + mdw_max]) : 0xffu;
        // Many dilations (DCVNet: 6): a (tile, dilation) pair has few work items (1-27 warp iterations), so whole pairs are dealt
        // to the EPI_SPLIT warps of a quadrant instead of every warp evaluating every pair's ranges for a quarter of its items.
        const bool pair_mode = P.D >= 4;
        const int w0 = pair_mode ? 0 : part, wstep = pair_mode ? 1 : EPI_SPLIT;
        static_assert(EPI_SPLIT == 4 && MAX_DIL == 8, "the pair-mode mask below deals dilations d = k, k + 4 to warp k");
        unsigned dm = dmask & ((1u << P.D) - 1u);
        if (pair_mode) dm &= 0x11u << ((part - tcount) & (EPI_SPLIT - 1));  // this warp's dilations: (d + tcount) % 4 == part
        for (; dm; dm &= dm - 1u) {
          const int d = __ffs(dm) - 1;
          const int dph = P.dph[d], dpw = P.dpw[d], mh = P.mh[d], mw = P.mw[d];
          // CTA-uniform early out: this dilation's box around the query tile does not reach the target tile
          if (!P.sparse && P.D > 1 && (ty1 < qbox_y0 - mh || ty0 > qbox_y1 + mh || tx1 < qbox_x0 - mw || tx0 > qbox_x1 + mw)) continue;
          const int y2_0 = yq - mh, x2_0 = xq - mw;
          // this lane's displacement ranges whose target lies in the in-image part of the tile (the reciprocals are launch constants)
          const float inv_dph = P.inv_dph[d], inv_dpw = P.inv_dpw[d];
          int a_lo = max(0, ceil_div(ty0 - y2_0, inv_dph)), a_hi = min(P.Ph - 1, floor_div(ty1 - y2_0, inv_dph));
          if (!q_ok || a_lo > a_hi) { a_lo = INT_MAX; a_hi = INT_MIN; }
          // all lanes walk the union so that a store instruction covers one (pi, pj): contiguous output pixels
          const int ua_lo = __reduce_min_sync(0xffffffffu, a_lo), ua_hi = __reduce_max_sync(0xffffffffu, a_hi);
          if (ua_lo > ua_hi) continue;  // no lattice row of this dilation in the tile (half of the tiles at dilation 16)
          int b_lo = max(0, ceil_div(tx0 - x2_0, inv_dpw)), b_hi = min(P.Pw - 1, floor_div(tx1 - x2_0, inv_dpw));
          if (a_lo > a_hi || b_lo > b_hi) { a_lo = b_lo = INT_MAX; a_hi = b_hi = INT_MIN; }
          const int ub_lo = __reduce_min_sync(0xffffffffu, b_lo), ub_hi = __reduce_max_sync(0xffffffffu, b_hi);
          if (ub_lo > ub_hi) continue;
          const unsigned o_d = (unsigned)((d * P.G + g) * P.Ph * P.Pw) * uplane;
          // work items = (pi, chunk of CH pj), dealt round-robin to the warps of the quadrant.  CH (3 or 4) is whichever
          // wastes fewer slots on this row width (a 9- or 21-wide window splits into chunks of 3 without remainder);
          // ACT: LeakyReLU fused (slope != 1).
          auto extract = [&](auto ch_tag, auto act_tag) {
            constexpr int CH = decltype(ch_tag)::value;
            constexpr bool ACT = decltype(act_tag)::value;
            const int nchunk = (ub_hi - ub_lo) / CH + 1, nitems = (ua_hi - ua_lo + 1) * nchunk;
            int ai = 0, ci = w0;  // work item wi = ai * nchunk + ci, kept incrementally (no division per item)
            while (ci >= nchunk) { ci -= nchunk; ++ai; }
            for (int wi = w0; wi < nitems; wi += wstep) {
              const int a = ua_lo + ai, bb = ub_lo + ci * CH;
              ci += wstep;
              while (ci >= nchunk) { ci -= nchunk; ++ai; }
              const bool row_ok = a >= a_lo && a <= a_hi;
              const int col0 = (y2_0 + a * dph - ty0) * TTW + (x2_0 - tx0) + bb * dpw;
              const float* srow = Srow + col0;
              const float* trow = tsq + col0;
              const unsigned o_ab = o_d + (unsigned)(a * P.Pw + bb) * uplane;
              float v[CH];
              bool ok[CH];
#pragma unroll
              for (int u = 0; u < CH; ++u) {  // CH independent load -> store chains in flight
                ok[u] = row_ok && bb + u >= b_lo && bb + u <= b_hi;
                v[u] = ok[u] ? srow[u * dpw] * trow[u * dpw] : 0.f;
              }
#pragma unroll
              for (int u = 0; u < CH; ++u) {
                float w = v[u] * deq;
                if (ACT) w = w >= 0.f ? w : w * P.slope;
                if (ok[u]) out_q[o_ab + u * uplane] = w;
              }
            }
          };
          // chunks of 3 where ceil(nb / 3) * 3 < ceil(nb / 4) * 4 (nb = 1, 2, 3, 6, 9 for nb <= 31; wider rows: chunks of 4)
          const int nb = ub_hi - ub_lo + 1;
          const bool ch3 = nb < 32 && ((kCh3Mask >> nb) & 1u), act = P.slope != 1.f;
          if (ch3) {
            if (act) extract(std::integral_constant<int, 3>{}, std::true_type{});
            else extract(std::integral_constant<int, 3>{}, std::false_type{});
          } else {
            if (act) extract(std::integral_constant<int, 4>{}, std::true_type{});
            else extract(std::integral_constant<int, 4>{}, std::false_type{});
          }
        }
        }
        if (P.regular == 1) __syncwarp(); else quad_sync();  // every warp is done with the rows before the next tile overwrites them
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, TMEM_COLS);
}

struct Workspace {
  float* qscale;  // [BG * lat^2][oh * ow]
  float* tscale;  // [BG][H * W]
  __half* in1h;   // queries: (BG * lat^2, oh, ow, Cp)
  __half* in2h;
};

static size_t ws_layout(int BG, int C, int H, int W, int padh, int padw, void* base, Workspace* ws) {
  const int Cp = round_up(C, BK);
  uintptr_t p = reinterpret_cast<uintptr_t>(base);
  auto take = [&](size_t bytes, size_t align) {
    p = (p + align - 1) / align * align;
    uintptr_t r = p;
    p += bytes;
    return r;
  };
  // one scale per operand pixel (the query tensor is at most (H + 2 padh + 8) x (W + 2 padw + 8) pixels per image, see below)
  const uintptr_t qsc = take((size_t)BG * (H + 2 * padh + 8) * (W + 2 * padw + 8) * 4, 256);
  const uintptr_t tsc = take((size_t)BG * H * W * 4, 256);
  // queries: (oh, ow) <= (H + 2 padh, W + 2 padw) pixels per map (stride only shrinks it)
  // (+8: in lattice mode every residue class is rounded up to the size of class (0,0))
  const uintptr_t a = take((size_t)BG * (H + 2 * padh + 8) * (W + 2 * padw + 8) * Cp * 2, 1024);
  const uintptr_t b = take((size_t)BG * H * W * Cp * 2, 1024);
  if (ws) {
    ws->qscale = reinterpret_cast<float*>(qsc);
    ws->tscale = reinterpret_cast<float*>(tsc);
    ws->in1h = reinterpret_cast<__half*>(a);
    ws->in2h = reinterpret_cast<__half*>(b);
  }
  return p - reinterpret_cast<uintptr_t>(base);
}

}  // namespace k4tc

// true when the tensor-core kernel covers this problem (otherwise the exact fp32 SIMT kernel runs)
bool local_corr_tc_supported(int C, int sh, int sw, int D) {
  (void)sh; (void)sw;  // any stride: the query operand is gathered compactly
  return C <= k4tc::KB_MAX * k4tc::BK && D <= k4tc::MAX_DIL;
}

size_t local_corr_tc_workspace_bytes(int BG, int C, int H, int W, int padh, int padw) {
  return k4tc::ws_layout(BG, C, H, W, padh, padw, nullptr, nullptr) + 1024;
}

int local_corr_tc_launch(const float* in1, const float* in2, const float* flow2, int BG, int G, int C, int H, int W, int Ph, int Pw,
                         int sh, int sw, int padh, int padw, int D, const int* dph, const int* dpw, float scale, float slope,
                         float* out, long long out_bstride, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  using namespace k4tc;
  const int Cp = round_up(C, BK), N = H * W;
  uintptr_t wbase = (reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023);
  Workspace ws;
  const size_t need = ws_layout(BG, C, H, W, padh, padw, reinterpret_cast<void*>(wbase), &ws) + (wbase - reinterpret_cast<uintptr_t>(workspace));
  if (need > workspace_bytes) return fail(EZB_ERR_WORKSPACE, "workspace too small: need %zu, got %zu", need, workspace_bytes);
  EZB_REQUIRE(BG <= 65535, "batch too large for one launch (split it)");

  const int out_h = cdiv(H + 2 * padh, sh), out_w = cdiv(W + 2 * padw, sw);
  // Lattice mode: one dilation lat > 1 at stride 1 without padding splits into lat^2 independent dilation-1 problems on
  // sub-lattices (4x fewer wasted products for FlowNetC's dilation_patch = 2); classes must still fill a query tile.
  int lat = 1;
  if (D == 1 && sh == 1 && sw == 1 && padh == 0 && padw == 0 && dph[0] == dpw[0] && dph[0] >= 2 && dph[0] <= 8 &&
      cdiv(H, dph[0]) >= QTH && cdiv(W, dph[0]) >= QTW && getenv("EZB_LOCAL_CORR_NO_LATTICE") == nullptr)
    lat = dph[0];
  const int ncls = lat * lat;
  // sizes in the units the tiles are laid out in (lattice mode: the lattice of residue class (0,0), the largest one)
  const int Hu = lat > 1 ? cdiv(H, lat) : H, Wu = lat > 1 ? cdiv(W, lat) : W;
  const int oh = lat > 1 ? Hu : out_h, ow = lat > 1 ? Wu : out_w;
  const int one = 1;
  const int* dph_u = lat > 1 ? &one : dph;
  const int* dpw_u = lat > 1 ? &one : dpw;
  EZB_REQUIRE((long long)BG * ncls <= 65535, "batch too large for one launch (split it)");

  // one launch converts the targets (all of in2, optionally warped by flow2) and the queries (the query pixels of in1)
  {
    ConvJobs JJ;
    ConvJob& T = JJ.j[0];
    T.src = in2; T.flow = flow2; T.dst = ws.in2h; T.rowscale = ws.tscale;
    T.C = C; T.Cp = Cp; T.H = H; T.W = W; T.oh = H; T.ow = W; T.sh = T.sw = 1; T.padh = T.padw = 0; T.lat = 1; T.nimg = BG; T.G = G;
    ConvJob& Q = JJ.j[1];
    Q.src = in1; Q.flow = nullptr; Q.dst = ws.in1h; Q.rowscale = ws.qscale;
    Q.C = C; Q.Cp = Cp; Q.H = H; Q.W = W; Q.oh = oh; Q.ow = ow; Q.sh = sh; Q.sw = sw; Q.padh = padh; Q.padw = padw; Q.lat = lat;
    Q.nimg = BG * ncls; Q.G = G;
    EZB_REQUIRE((long long)BG * (1 + ncls) <= 65535, "batch too large for one launch (split it)");
    JJ.groups = Cp <= 64 ? 4 : 1;  // measured (ncu): 4 concurrent groups help at 64 channels (58 -> 47 us on PWC-Net's finest level), not at 128
    if (const char* e = getenv("EZB_LOCAL_CORR_CONV_GROUPS")) JJ.groups = std::max(1, std::min(CONV_MAX_PX / 32, atoi(e)));  // experiments
    const int conv_px = 32 * JJ.groups;
    const size_t conv_smem = (size_t)Cp * (conv_px + 1) * 4;
    EZB_CUDA_OK(cudaFuncSetAttribute(convert_pixels_kmajor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv_smem));
    convert_pixels_kmajor_kernel<<<dim3(cdiv(std::max(H * W, oh * ow), conv_px), 1, BG * (1 + ncls)), CONV_THREADS, conv_smem, st>>>(JJ);
    EZB_LAUNCH_OK();
  }

  CUtensorMap tmA, tmB;
  {
    cuuint64_t qdims[4] = {(cuuint64_t)Cp, (cuuint64_t)ow, (cuuint64_t)oh, (cuuint64_t)BG * ncls};
    cuuint64_t qstr[3] = {(cuuint64_t)Cp * 2, (cuuint64_t)ow * Cp * 2, (cuuint64_t)oh * ow * Cp * 2};
    cuuint32_t box[4] = {BK, QTW, QTH, 1};
    if (int r = make_tmap(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, ws.in1h, qdims, qstr, box, CU_TENSOR_MAP_SWIZZLE_128B)) return r;
  }
  int dmax_h = 1, dmax_w = 1;
  for (int d = 0; d < D; ++d) { dmax_h = std::max(dmax_h, dph_u[d]); dmax_w = std::max(dmax_w, dpw_u[d]); }
  // x-extent of a query tile's bounding box (layout units): if it fits, the target tile spans it (8x24 / 4x40)
  const int xext = (QTW - 1) * sw + (Pw - 1) * dmax_w + 1;
  const int TTW = xext <= 24 ? 24 : xext <= 40 ? 40 : 16, TTH = TTW == 40 ? 4 : 8;
  {
    // targets: the full-resolution in2; in lattice mode a box takes every lat-th pixel (TMA element strides)
    cuuint64_t dims[4] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)BG};
    cuuint64_t str[3] = {(cuuint64_t)Cp * 2, (cuuint64_t)W * Cp * 2, (cuuint64_t)H * W * Cp * 2};
    cuuint32_t box[4] = {BK, (cuuint32_t)(TTW * lat), (cuuint32_t)(TTH * lat), 1};
    cuuint32_t es[4] = {1, (cuuint32_t)lat, (cuuint32_t)lat, 1};
    if (int r = make_tmap_ex(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, ws.in2h, dims, str, box, es, CU_TENSOR_MAP_SWIZZLE_128B)) return r;
  }
  Params P;
  P.G = G; P.H = Hu; P.W = Wu;
  P.oh = oh; P.ow = ow;
  P.Ph = Ph; P.Pw = Pw; P.sh = sh; P.sw = sw; P.padh = padh; P.padw = padw;
  P.D = D; P.KB = Cp / BK;
  for (int d = 0; d < D; ++d) {
    P.dph[d] = dph_u[d]; P.dpw[d] = dpw_u[d];
    P.mh[d] = dph_u[d] * (Ph - 1) / 2; P.mw[d] = dpw_u[d] * (Pw - 1) / 2;
    P.inv_dph[d] = 1.f / (float)dph_u[d]; P.inv_dpw[d] = 1.f / (float)dpw_u[d];
  }
  P.regular = (D == 1 && dph_u[0] == 1 && dpw_u[0] == 1 && sh == 1 && sw == 1 && padh == 0 && padw == 0 && TTW != 16 &&
               getenv("EZB_LOCAL_CORR_GENERIC_EXTRACT") == nullptr) ? (getenv("EZB_LOCAL_CORR_QUADSYNC") ? 2 : 1) : 0;
  P.lat = lat; P.BG = BG; P.Hf = H; P.Wf = W; P.out_h = out_h; P.out_w = out_w;
  // Sparse sweep: which tiles of the bounding box hold a target at all.  A tile whose first row is `rel` rows below the query
  // box's first row holds a target of displacement `off` iff rel <= r + off <= rel + TTH - 1 for a query row offset
  // 0 <= r <= (QTH - 1) * sh (stride granularity ignored: a superset), i.e. rel in [off - (TTH - 1), off + (QTH - 1) * sh].
  P.sparse = 0;
  {
    int mh_max = 0, mw_max = 0;
    for (int d = 0; d < D; ++d) { mh_max = std::max(mh_max, P.mh[d]); mw_max = std::max(mw_max, P.mw[d]); }
    const int ny = 2 * mh_max + (QTH - 1) * sh + 1, nx = 2 * mw_max + (QTW - 1) * sw + 1;
    if (P.regular == 0 && lat == 1 && ny <= SWEEP_TAB && nx <= SWEEP_TAB && getenv("EZB_LOCAL_CORR_DENSE_SWEEP") == nullptr) {
      std::memset(P.rowtab, 0, sizeof(P.rowtab));
      std::memset(P.coltab, 0, sizeof(P.coltab));
      for (int d = 0; d < D; ++d) {
        for (int a = 0; a < Ph; ++a) {
          const int off = a * P.dph[d] - P.mh[d];
          for (int rel = off - (TTH - 1); rel <= off + (QTH - 1) * sh; ++rel)
            if (rel + mh_max >= 0 && rel + mh_max < ny) P.rowtab[rel + mh_max] |= (unsigned char)(1u << d);
        }
        for (int b = 0; b < Pw; ++b) {
          const int off = b * P.dpw[d] - P.mw[d];
          for (int rel = off - (TTW - 1); rel <= off + (QTW - 1) * sw; ++rel)
            if (rel + mw_max >= 0 && rel + mw_max < nx) P.coltab[rel + mw_max] |= (unsigned char)(1u << d);
        }
      }
      P.sparse = 1;
    }
  }
  // many dilations (DCVNet): few zeros, many (row, column) tests -> the idle warps; FlowNetC / PWC-Net: a third of the outputs
  // are zeros, which two warps cannot store fast enough (measured: 44 -> 64 us on config 2) -> the sixteen epilogue warps
  P.zero_idle = D >= 4 ? 1 : 0;
  if (const char* e = getenv("EZB_LOCAL_CORR_ZERO_IDLE")) P.zero_idle = atoi(e) != 0;
  P.xclip_unroll = 1;
  if (const char* e = getenv("EZB_LOCAL_CORR_XCLIP_UNROLL")) P.xclip_unroll = atoi(e) != 0;
  P.slope = slope;
  P.qscale = ws.qscale;
  P.tscale = ws.tscale;
  P.scale = scale;
  P.out = out;
  P.out_bstride = out_bstride;
  // split the target sweep of a query tile over several CTAs when there are too few query tiles to fill the GPU evenly
  const int n_qtiles = cdiv(ow, QTW) * cdiv(oh, QTH) * BG * ncls;
  const int max_tiles = std::min(cdiv(Hu, TTH) + 1, cdiv((Ph - 1) * dmax_h + QTH * sh, TTH) + 1) *
                        std::min(cdiv(Wu, TTW) + 1, cdiv((Pw - 1) * dmax_w + QTW * sw, TTW) + 1);
  // cost model: CTAs run in waves of one per SM; a CTA costs its share of the target tiles plus ~2 tiles of fixed work
  P.nsplit = 1;
  {
    const int sms = num_sms();
    long long best = (long long)cdiv(n_qtiles, sms) * (max_tiles + 2);
    for (int k = 2; k <= 8 && max_tiles / k >= 4; ++k) {
      const long long c = (long long)cdiv(n_qtiles * k, sms) * (cdiv(max_tiles, k) + 2);
      if (c < best) { best = c; P.nsplit = k; }
    }
  }
  if (const char* e = getenv("EZB_LOCAL_CORR_NSPLIT")) P.nsplit = std::max(1, std::min(8, atoi(e)));  // experiments
  P.qtx = cdiv(ow, QTW);
  P.qty = cdiv(oh, QTH);
  P.d_qtx = make_fastdiv(P.qtx); P.d_qty = make_fastdiv(P.qty); P.d_nsplit = make_fastdiv(P.nsplit);
  P.d_ncls = make_fastdiv(ncls); P.d_lat = make_fastdiv(lat); P.d_G = make_fastdiv(G);
  EZB_REQUIRE((long long)D * G * Ph * Pw * out_h * out_w < (1ll << 31), "output per batch entry too large for the tensor-core path");
  const long long n_items = (long long)P.qtx * P.qty * BG * ncls * P.nsplit;
  EZB_REQUIRE(n_items < (1ll << 31), "output too large for one launch");
  P.n_items = (int)n_items;
  // persistent CTAs (one per SM: the kernel uses most of the shared memory) walk the items round-robin, so that the query
  // load + MMAs of the next item overlap the extraction of the current one and the per-CTA set-up is paid once
  const int grid = getenv("EZB_LOCAL_CORR_NONPERSISTENT") ? P.n_items : std::min(P.n_items, num_sms());
  if (TTW == 24) {
    EZB_CUDA_OK(cudaFuncSetAttribute(local_corr_tc_kernel<24, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<24, 8>::SMEM_BYTES));
    local_corr_tc_kernel<24, 8><<<grid, NUM_THREADS, Geo<24, 8>::SMEM_BYTES, st>>>(tmA, tmB, P);
  } else if (TTW == 40) {
    EZB_CUDA_OK(cudaFuncSetAttribute(local_corr_tc_kernel<40, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<40, 4>::SMEM_BYTES));
    local_corr_tc_kernel<40, 4><<<grid, NUM_THREADS, Geo<40, 4>::SMEM_BYTES, st>>>(tmA, tmB, P);
  } else {
    EZB_CUDA_OK(cudaFuncSetAttribute(local_corr_tc_kernel<16, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<16, 8>::SMEM_BYTES));
    local_corr_tc_kernel<16, 8><<<grid, NUM_THREADS, Geo<16, 8>::SMEM_BYTES, st>>>(tmA, tmB, P);
  }
  EZB_LAUNCH_OK();
  return EZB_OK;
}

}  // namespace ezb
