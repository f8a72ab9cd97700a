// K4 / K4b / K6: local displacement-window correlation (FlowNetC, PWC-Net, CorrelationLayer) and the
// DCVNet Matryoshka multi-dilation cost volume built on it.
//
// Replaces iter_spatial_correlation_sample (reference ezflow/similarity/correlation/sampler.py:29-129)
// and MatryoshkaDilatedCostVolume.forward (ezflow/similarity/learnable_cost.py:306-325):
//   out[b,pi,pj,y,x] = sum_c in1p[b,c,y*sh,x*sw] * in2p[b,c, y*sh + pi*dph - mdh, x*sw + pj*dpw - mdw]
// where in*p are the inputs zero-padded by (padh,padw) and md = dp*(P-1)/2.
//
// Forward kernel: a block owns one output row segment (64 pixels) of one (batch*group, dilation) and
// walks the vertical displacements; for each it streams channel chunks of the in1 row segment and of
// the matching in2 row window through shared memory, every thread accumulating up to 6 horizontal
// displacements for its pixel in registers (fp32 exact).  Output rows are written coalesced, straight
// into the (possibly concatenated, K6) destination; scale and LeakyReLU are fused.
#include "ezb_common.cuh"

#include <cstdlib>

namespace ezb {
namespace k4 {

constexpr int TX = 64;       // output pixels per block (along x)
constexpr int NT = 256;      // threads: 64 pixels x 4 displacement groups
constexpr int CC = 16;       // channels per shared-memory chunk
constexpr int NJ = 6;        // horizontal displacements per thread per pass (covers Pw <= 24 in one pass)
constexpr int MAX_DIL = 8;

struct Params {
  const float* in1;
  const float* in2;
  float* out;
  int BG, G, C, H, W;  // inputs viewed (BG = B*G, C per group, H, W)
  int Ph, Pw, sh, sw, padh, padw;
  int D;               // number of dilations
  int dph[MAX_DIL], dpw[MAX_DIL];
  int oh, ow;
  int rx_max;          // shared-memory row length of the in2 window
  int pi_split;        // 1: a block walks all displacement rows; Ph: blockIdx.z also indexes the displacement row
  float scale, slope;
  long long out_bstride;
};

__global__ void __launch_bounds__(NT) local_corr_fwd_kernel(const Params P) {
  extern __shared__ float sm[];
  float* sA = sm;                 // [CC][TX]
  float* sB = sm + CC * TX;       // [CC][rx_max]
  const int x0 = blockIdx.x * TX, y = blockIdx.y;
  // displacement rows are independent: with pi in the grid a small map still fills the GPU (a block per output row
  // walking all P rows was latency-bound: 285 us for a 7x16 map)
  const int zz = blockIdx.z / P.pi_split, pi_only = blockIdx.z - zz * P.pi_split;
  const int bg = zz / P.D, d = zz - bg * P.D;
  const int b = bg / P.G, g = bg - b * P.G;
  const int dph = P.dph[d], dpw = P.dpw[d];
  const int mdh = dph * (P.Ph - 1) / 2, mdw = dpw * (P.Pw - 1) / 2;
  const int tid = threadIdx.x, xl = tid & (TX - 1), jg = tid / TX;
  const int RX = (TX - 1) * P.sw + (P.Pw - 1) * dpw + 1;
  const long long plane = (long long)P.H * P.W;
  const float* in1 = P.in1 + (long long)bg * P.C * plane;
  const float* in2 = P.in2 + (long long)bg * P.C * plane;
  const int y1 = y * P.sh - P.padh;             // in1 row (unpadded coordinates)
  const bool y1_ok = y1 >= 0 && y1 < P.H;
  const int xb = x0 * P.sw - mdw - P.padw;      // first in2 column of the window (unpadded coordinates)
  const long long ohw = (long long)P.oh * P.ow;
  float* out = P.out + (long long)b * P.out_bstride + ((long long)(d * P.G + g) * P.Ph * P.Pw) * ohw + (long long)y * P.ow;
  const bool x_ok = x0 + xl < P.ow;

  const int pi_begin = P.pi_split > 1 ? pi_only : 0, pi_end = P.pi_split > 1 ? pi_only + 1 : P.Ph;
  for (int pi = pi_begin; pi < pi_end; ++pi) {
    const int y2 = y * P.sh + pi * dph - mdh - P.padh;
    const bool rows_ok = y1_ok && y2 >= 0 && y2 < P.H;
    for (int jbase = 0; jbase < P.Pw; jbase += 4 * NJ) {
      float acc[NJ];
#pragma unroll
      for (int m = 0; m < NJ; ++m) acc[m] = 0.f;
      if (rows_ok) {
        for (int c0 = 0; c0 < P.C; c0 += CC) {
          for (int i = tid; i < CC * TX; i += NT) {
            const int cc = i / TX, xx = i - cc * TX;
            const int c = c0 + cc, x1 = (x0 + xx) * P.sw - P.padw;
            sA[i] = (c < P.C && x0 + xx < P.ow && x1 >= 0 && x1 < P.W) ? __ldg(in1 + c * plane + (long long)y1 * P.W + x1) : 0.f;
          }
          for (int i = tid; i < CC * RX; i += NT) {
            const int cc = i / RX, rx = i - cc * RX;
            const int c = c0 + cc, x2 = xb + rx;
            sB[cc * P.rx_max + rx] = (c < P.C && x2 >= 0 && x2 < P.W) ? __ldg(in2 + c * plane + (long long)y2 * P.W + x2) : 0.f;
          }
          __syncthreads();
#pragma unroll 4
          for (int cc = 0; cc < CC; ++cc) {
            const float a = sA[cc * TX + xl];
            const float* brow = sB + cc * P.rx_max + xl * P.sw;
#pragma unroll
            for (int m = 0; m < NJ; ++m) {
              const int pj = jbase + jg + 4 * m;
              if (pj < P.Pw) acc[m] = fmaf(a, brow[pj * dpw], acc[m]);
            }
          }
          __syncthreads();
        }
      }
      if (x_ok) {
#pragma unroll
        for (int m = 0; m < NJ; ++m) {
          const int pj = jbase + jg + 4 * m;
          if (pj < P.Pw) {
            float v = acc[m] * P.scale;
            v = v >= 0.f ? v : v * P.slope;
            out[(long long)(pi * P.Pw + pj) * ohw + x0 + xl] = v;
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------ K4b
struct BwdParams {
  const float* in1;
  const float* in2;
  const float* dout;
  float* din1;
  float* din2;
  int B, C, H, W, Ph, Pw, sh, sw, padh, padw, dph, dpw, oh, ow;
  float scale;
  long long dout_bstride;
};

// din1[b,c,y1,x1] = scale * sum_{pi,pj} dout[b,pi,pj,y,x] * in2[b,c,y2,x2]  (only strided (y1,x1) are hit)
__global__ void local_corr_bwd_in1_kernel(const BwdParams P) {
  const long long total = (long long)P.B * P.C * P.H * P.W;
  const int mdh = P.dph * (P.Ph - 1) / 2, mdw = P.dpw * (P.Pw - 1) / 2;
  const long long ohw = (long long)P.oh * P.ow;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x1 = (int)(i % P.W), y1 = (int)((i / P.W) % P.H);
    const int c = (int)((i / ((long long)P.W * P.H)) % P.C), b = (int)(i / ((long long)P.W * P.H * P.C));
    float acc = 0.f;
    const int yp = y1 + P.padh, xp = x1 + P.padw;  // padded coordinates
    if (yp % P.sh == 0 && xp % P.sw == 0) {
      const int y = yp / P.sh, x = xp / P.sw;
      const float* d = P.dout + (long long)b * P.dout_bstride + (long long)y * P.ow + x;
      const float* i2 = P.in2 + ((long long)b * P.C + c) * P.H * P.W;
      for (int pi = 0; pi < P.Ph; ++pi) {
        const int y2 = yp + pi * P.dph - mdh - P.padh;
        if (y2 < 0 || y2 >= P.H) continue;
        for (int pj = 0; pj < P.Pw; ++pj) {
          const int x2 = xp + pj * P.dpw - mdw - P.padw;
          if (x2 < 0 || x2 >= P.W) continue;
          acc = fmaf(__ldg(d + (long long)(pi * P.Pw + pj) * ohw), __ldg(i2 + (long long)y2 * P.W + x2), acc);
        }
      }
    }
    P.din1[i] = acc * P.scale;
  }
}

// din2[b,c,y2,x2] = scale * sum_{pi,pj} dout[b,pi,pj,y,x] * in1[b,c,y1,x1] with y*sh = y2p - pi*dph + mdh
__global__ void local_corr_bwd_in2_kernel(const BwdParams P) {
  const long long total = (long long)P.B * P.C * P.H * P.W;
  const int mdh = P.dph * (P.Ph - 1) / 2, mdw = P.dpw * (P.Pw - 1) / 2;
  const long long ohw = (long long)P.oh * P.ow;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x2 = (int)(i % P.W), y2 = (int)((i / P.W) % P.H);
    const int c = (int)((i / ((long long)P.W * P.H)) % P.C), b = (int)(i / ((long long)P.W * P.H * P.C));
    const float* d = P.dout + (long long)b * P.dout_bstride;
    const float* i1 = P.in1 + ((long long)b * P.C + c) * P.H * P.W;
    float acc = 0.f;
    for (int pi = 0; pi < P.Ph; ++pi) {
      const int yp = y2 + P.padh - pi * P.dph + mdh;  // padded in1 row = y*sh
      if (yp < 0 || yp % P.sh) continue;
      const int y = yp / P.sh, y1 = yp - P.padh;
      if (y >= P.oh || y1 < 0 || y1 >= P.H) continue;
      for (int pj = 0; pj < P.Pw; ++pj) {
        const int xp = x2 + P.padw - pj * P.dpw + mdw;
        if (xp < 0 || xp % P.sw) continue;
        const int x = xp / P.sw, x1 = xp - P.padw;
        if (x >= P.ow || x1 < 0 || x1 >= P.W) continue;
        acc = fmaf(__ldg(d + (long long)(pi * P.Pw + pj) * ohw + (long long)y * P.ow + x), __ldg(i1 + (long long)y1 * P.W + x1), acc);
      }
    }
    P.din2[i] = acc * P.scale;
  }
}

// Stride-1 adjoints, tiled (FlowNetC, PWC-Net, CorrelationLayer, DCVNet's stride-1 scales).  A block owns an 8 x 32 tile of
// gradient pixels and a chunk of 16 channels; thread = pixel (lane = x: every shared-memory access below is
// conflict-free and every global access coalesced), 16 accumulators per thread.  Per displacement row pi the block stages
//   osh[ty][j][c]    the other operand's row for that pi with a halo of (Pw - 1) * dpw columns (zero outside the image),
//                    channel-innermost with a pitch of 20 floats: the 16 channels of a pixel are four conflict-free LDS.128
// and every thread accumulates  acc[c] += d[pj] * osh[x + offset(pj)][c], d[pj] = the upstream gradient of displacement (pi, pj) at
// the pixel that pairs with (ty, x) — private to the thread, read from global memory a few displacements ahead.
//   SECOND = false: din1[c,y1,x1] = sum dout[pi,pj,y1+padh,x1+padw] * in2[c, y1 + pi*dph - mdh, x1 + pj*dpw - mdw]
//   SECOND = true : din2[c,y2,x2] = sum dout[pi,pj,y1+padh,x1+padw] * in1[c,y1,x1],  (y1,x1) = (y2 - pi*dph + mdh, x2 - pj*dpw + mdw)
// The gather kernels above cost P^2 uncached-pattern loads per output element (9.7 ms for FlowNetC's 21 x 21 window at
// batch 8); here the P^2 products come out of shared memory.
constexpr int BT_Y = 8, BT_X = 32, BT_C = 16, BT_CP = 20;  // BT_CP: channel pitch of a staged pixel (floats)
template <bool SECOND>
__global__ void __launch_bounds__(BT_Y * BT_X) local_corr_bwd_tiled_kernel(const BwdParams P) {
  extern __shared__ float smem_bt[];
  const int halo = (P.Pw - 1) * P.dpw, ow_s = BT_X + halo;  // staged row width
  float* osh = smem_bt;                           // [BT_Y][ow_s][BT_CP]
  const int lane = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int xtiles = (P.W + BT_X - 1) / BT_X;
  const int x0 = (blockIdx.x % xtiles) * BT_X, y0 = (blockIdx.x / xtiles) * BT_Y;
  const int c0 = blockIdx.y * BT_C, b = blockIdx.z;
  const int mdh = P.dph * (P.Ph - 1) / 2, mdw = P.dpw * (P.Pw - 1) / 2;
  const long long HW = (long long)P.H * P.W, ohw = (long long)P.oh * P.ow;
  const float* other = (SECOND ? P.in1 : P.in2) + ((long long)b * P.C + c0) * HW;
  const float* dout = P.dout + (long long)b * P.dout_bstride;
  float acc[BT_C];
#pragma unroll
  for (int c = 0; c < BT_C; ++c) acc[c] = 0.f;
  const int nch = min(BT_C, P.C - c0);
  // first staged column of the other operand: din1 reads x1 + pj*dpw - mdw, din2 reads x2 - pj*dpw + mdw
  const int xb = SECOND ? x0 + mdw - halo : x0 - mdw;

  // Displacement rows are walked so that the other operand's rows only ever move down: step `it` needs the image rows
  // ybase + a for a in [it*dph, it*dph + BT_Y) (din2 walks pi backwards for that), kept in a ring of BT_Y row slots
  // (slot = a mod BT_Y), so a step stages only the min(dph, BT_Y) rows its predecessor did not have.
  const int ybase = SECOND ? y0 + mdh - (P.Ph - 1) * P.dph : y0 - mdh;
  const int rows_new = min(P.dph, BT_Y);
  for (int it = 0; it < P.Ph; ++it) {
    const int pi = SECOND ? P.Ph - 1 - it : it;
    const int dy = SECOND ? mdh - pi * P.dph : pi * P.dph - mdh;  // row of the other operand = tile row + dy
    const int a0 = it * P.dph;
    const int a_first = it == 0 ? a0 : a0 + BT_Y - rows_new, n_rows = it == 0 ? BT_Y : rows_new;
    __syncthreads();  // the previous iteration's reads are done
    // ---- stage the other operand's new rows (zero outside the image / beyond C)
    // (four independent loads per thread are issued before the first store: the staging is latency-bound otherwise)
    for (int r = ty; r < (BT_C / 4) * n_rows; r += BT_Y) {  // warp = (group of 4 channels, new row), lanes sweep the columns
      const int cg = r / n_rows, a = a_first + (r - cg * n_rows);
      const int ys = ybase + a, slot = a & (BT_Y - 1);
      const bool row_in = ys >= 0 && ys < P.H;
      const float* src = other + (long long)(cg * 4) * HW + (long long)ys * P.W;
      for (int j = lane; j < ow_s; j += 32) {
        const int xs = xb + j;
        const bool ok = row_in && xs >= 0 && xs < P.W;
        float4 v;
        v.x = (ok && cg * 4 + 0 < nch) ? __ldg(src + xs) : 0.f;
        v.y = (ok && cg * 4 + 1 < nch) ? __ldg(src + HW + xs) : 0.f;
        v.z = (ok && cg * 4 + 2 < nch) ? __ldg(src + 2 * HW + xs) : 0.f;
        v.w = (ok && cg * 4 + 3 < nch) ? __ldg(src + 3 * HW + xs) : 0.f;
        *reinterpret_cast<float4*>(osh + (slot * ow_s + j) * BT_CP + cg * 4) = v;
      }
    }
    // ---- the upstream gradient dout[pi, pj] at the pixel that pairs with (ty, lane) is private to this thread: it comes
    // straight from global memory (L2) into registers, four displacements ahead of their use
    const int y1 = SECOND ? y0 + ty + dy : y0 + ty;  // the in1 pixel (= query) row this gradient element belongs to
    const bool y_ok = y1 >= 0 && y1 < P.H;
    const float* drow_g = dout + (long long)pi * P.Pw * ohw + (long long)(y1 + P.padh) * P.ow + P.padw;
    const int xq = SECOND ? x0 + lane + mdw : x0 + lane;
    auto load_d = [&](int pj) {
      const int x1 = SECOND ? xq - pj * P.dpw : xq;
      return (pj < P.Pw && y_ok && x1 >= 0 && x1 < P.W) ? __ldg(drow_g + (long long)pj * ohw + x1) : 0.f;
    };
    float dn[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) dn[u] = load_d(u);
    __syncthreads();
    const float4* orow = reinterpret_cast<const float4*>(osh + (((a0 + ty) & (BT_Y - 1)) * ow_s + lane + (SECOND ? halo : 0)) * BT_CP);
    const int ostep = (SECOND ? -P.dpw : P.dpw) * (BT_CP / 4);
    for (int pj0 = 0; pj0 < P.Pw; pj0 += 4) {
      float dc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        dc[u] = dn[u];
        dn[u] = load_d(pj0 + 4 + u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (pj0 + u < P.Pw) {
          const float d = dc[u];
          const float4* o = orow + (pj0 + u) * ostep;
#pragma unroll
          for (int g = 0; g < BT_C / 4; ++g) {
            const float4 v = o[g];
            acc[4 * g + 0] = fmaf(d, v.x, acc[4 * g + 0]);
            acc[4 * g + 1] = fmaf(d, v.y, acc[4 * g + 1]);
            acc[4 * g + 2] = fmaf(d, v.z, acc[4 * g + 2]);
            acc[4 * g + 3] = fmaf(d, v.w, acc[4 * g + 3]);
          }
        }
      }
    }
  }
  const int y = y0 + ty, x = x0 + lane;
  if (y < P.H && x < P.W) {
    float* dst = (SECOND ? P.din2 : P.din1) + ((long long)b * P.C + c0) * HW + (long long)y * P.W + x;
#pragma unroll
    for (int c = 0; c < BT_C; ++c)
      if (c < nch) dst[(long long)c * HW] = acc[c] * P.scale;
  }
}

// x / (||x||_2 over channels + eps)        learnable_cost.py:426-428
__global__ void l2_normalize_kernel(const float* __restrict__ x, float* __restrict__ out, int C, long long HW, long long total_px,
                                    float eps) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_px; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, p = i - b * HW;
    const float* src = x + b * C * HW + p;
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
      const float v = __ldg(src + c * HW);
      s = fmaf(v, v, s);
    }
    const float inv = 1.0f / (sqrtf(s) + eps);
    float* dst = out + b * C * HW + p;
    for (int c = 0; c < C; ++c) dst[c * HW] = __ldg(src + c * HW) * inv;
  }
}

static int launch_fwd(Params& P, cudaStream_t st) {
  int rx = 0;
  for (int d = 0; d < P.D; ++d) rx = std::max(rx, (TX - 1) * P.sw + (P.Pw - 1) * P.dpw[d] + 1);
  P.rx_max = rx | 1;
  const size_t smem = (size_t)(CC * TX + CC * P.rx_max) * sizeof(float);
  if (smem > 200 * 1024)
    return fail(EZB_ERR_UNSUPPORTED, "displacement window too wide for the shared-memory tile (patch %d, dilation_patch up to %d, stride %d)",
                P.Pw, P.dpw[0], P.sw);
  EZB_CUDA_OK(cudaFuncSetAttribute(local_corr_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // small maps only: with enough row blocks to fill the GPU, splitting just re-stages the in1 row P times more often
  const long long row_blocks = (long long)cdiv(P.ow, TX) * P.oh * P.BG * P.D;
  P.pi_split = (row_blocks < 2000 && (long long)P.BG * P.D * P.Ph <= 65535) ? P.Ph : 1;
  dim3 grid(cdiv(P.ow, TX), P.oh, P.BG * P.D * P.pi_split);
  EZB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "output too large for one launch");
  local_corr_fwd_kernel<<<grid, NT, smem, st>>>(P);
  EZB_LAUNCH_OK();
  return EZB_OK;
}

}  // namespace k4
}  // namespace ezb

namespace ezb {
// local_corr_tc.cu
bool local_corr_tc_supported(int C, int sh, int sw, int D);
size_t local_corr_tc_workspace_bytes(int BG, int C, int H, int W, int padh, int padw);
int local_corr_tc_launch(const float* in1, const float* in2, const float* flow2, int BG, int G, int C, int H, int W, int Ph, int Pw, int sh, int sw,
                         int padh, int padw, int D, const int* dph, const int* dpw, float scale, float slope, float* out,
                         long long out_bstride, void* workspace, size_t workspace_bytes, cudaStream_t st);
// local_corr_rt.cu
bool local_corr_rt_supported(int H, int W, int BG, int Ph, int Pw, int sh, int sw, int padh, int padw, int dph, int dpw);
bool local_corr_rt_preferred(int BG, int C, int H, int W, bool tc_available);
int local_corr_rt_launch(const float* in1, const float* in2, int BG, int G, int C, int H, int W, int P, float scale,
                         float slope, float* out, long long out_bstride, cudaStream_t st);
bool local_corr_bwd_rt_supported(const float* in1, const float* in2, const float* dout, const float* din1, const float* din2, int B, int C,
                                 int H, int W, int Ph, int Pw, int sh, int sw, int padh, int padw, int dph, int dpw, long long dout_bstride);
int local_corr_bwd_rt_launch(const float* in1, const float* in2, const float* dout, int B, int C, int H, int W, int P, float scale,
                             long long dout_bstride, float* din1, float* din2, cudaStream_t st);
bool local_corr_rt_strided_supported(const float* in1, const float* in2, const float* out, int BG, int C, int H, int W, int Ph, int Pw,
                                     int sh, int sw, int padh, int padw, int dph, int dpw, long long out_bstride);
int local_corr_rt_strided_launch(const float* in1, const float* in2, int BG, int G, int C, int H, int W, int P, int S, float scale, float slope,
                                 float* out, long long out_bstride, cudaStream_t st);
bool local_corr_bwd_strided_supported(const float* in1, const float* in2, const float* dout, const float* din1, const float* din2, int BG,
                                      int C, int H, int W, int Ph, int Pw, int sh, int sw, int padh, int padw, int dph, int dpw);
int local_corr_bwd_strided_launch(const float* in1, const float* in2, const float* dout, int BG, int G, int C, int H, int W, int P, float scale,
                                  long long dout_bstride, float* din1, float* din2, cudaStream_t st);
// local_corr_bwd_tc.cu
bool local_corr_bwd_tc_supported(int C, int H, int W, int Ph, int Pw, int sh, int sw, int padh, int padw, int dph, int dpw);
size_t local_corr_bwd_tc_workspace_bytes(int BG, int C, int H, int W);
int local_corr_bwd_tc_launch(const float* in1, const float* in2, const float* dout, long long dout_bstride, int BG, int G, int C, int H, int W,
                             int Pn, int dil, float scale, int accumulate, float* din1, float* din2, void* workspace, size_t workspace_bytes,
                             cudaStream_t st);
}  // namespace ezb

using namespace ezb;

extern "C" int ezb_warp_fwd(const float* x, const float* flow, int B, int C, int H, int W, float* out, void* stream);  // warp.cu

static int check_local_args(int B, int C, int H, int W, int Ph, int Pw, int sh, int sw, int padh, int padw, int dph, int dpw) {
  EZB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "bad input shape B=%d C=%d H=%d W=%d", B, C, H, W);
  EZB_REQUIRE(Ph > 0 && Pw > 0 && sh > 0 && sw > 0 && padh >= 0 && padw >= 0 && dph > 0 && dpw > 0, "bad correlation parameters");
  if ((Ph % 2) == 0 || (Pw % 2) == 0) return fail(EZB_ERR_UNSUPPORTED, "Only odd patch sizes are supperted.");  // sampler.py:93-94
  return EZB_OK;
}

extern "C" int ezb_local_corr_workspace_bytes(int BG, int C, int H, int W, int padh, int padw, size_t* bytes) {
  EZB_REQUIRE(BG > 0 && C > 0 && H > 0 && W > 0 && padh >= 0 && padw >= 0 && bytes, "bad arguments");
  *bytes = local_corr_tc_workspace_bytes(BG, C, H, W, padh, padw);
  return EZB_OK;
}

extern "C" int ezb_local_corr_fwd(const float* in1, const float* in2, int B, int C, int H, int W, int Ph, int Pw, int sh,
                                  int sw, int padh, int padw, int dph, int dpw, float scale, float slope, float* out,
                                  int64_t out_batch_stride, void* workspace, size_t workspace_bytes, void* stream) {
  EZB_REQUIRE(in1 && in2 && out, "null pointer");
  if (int r = check_local_args(B, C, H, W, Ph, Pw, sh, sw, padh, padw, dph, dpw)) return r;
  const bool tc_ok = workspace != nullptr && local_corr_tc_supported(C, sh, sw, 1);
  // PWC-Net's shape (9x9 window, stride 1) on large maps with few channels: register-tiled fp32 kernel, no operand conversion
  if (local_corr_rt_supported(H, W, B, Ph, Pw, sh, sw, padh, padw, dph, dpw) && local_corr_rt_preferred(B, C, H, W, tc_ok))
    return local_corr_rt_launch(in1, in2, B, 1, C, H, W, Ph, scale, slope, out,
                                out_batch_stride > 0 ? out_batch_stride : (int64_t)Ph * Pw * H * W, static_cast<cudaStream_t>(stream));
  {  // stride 2 / 4 (16 or fewer window products per input pixel): strided register-tiled fp32 kernel, bound by reading in2 once
    const int64_t obs = out_batch_stride > 0 ? out_batch_stride : (int64_t)Ph * Pw * cdiv(H, sh) * cdiv(W, sw);
    if (local_corr_rt_strided_supported(in1, in2, out, B, C, H, W, Ph, Pw, sh, sw, padh, padw, dph, dpw, obs))
      return local_corr_rt_strided_launch(in1, in2, B, 1, C, H, W, Ph, sh, scale, slope, out, obs, static_cast<cudaStream_t>(stream));
  }
  if (tc_ok) {
    const int oh = cdiv(H + 2 * padh, sh), ow = cdiv(W + 2 * padw, sw);
    return local_corr_tc_launch(in1, in2, nullptr, B, 1, C, H, W, Ph, Pw, sh, sw, padh, padw, 1, &dph, &dpw, scale, slope, out,
                                out_batch_stride > 0 ? out_batch_stride : (int64_t)Ph * Pw * oh * ow, workspace, workspace_bytes,
                                static_cast<cudaStream_t>(stream));
  }
  k4::Params P;
  P.in1 = in1; P.in2 = in2; P.out = out;
  P.BG = B; P.G = 1; P.C = C; P.H = H; P.W = W;
  P.Ph = Ph; P.Pw = Pw; P.sh = sh; P.sw = sw; P.padh = padh; P.padw = padw;
  P.D = 1; P.dph[0] = dph; P.dpw[0] = dpw;
  P.oh = cdiv(H + 2 * padh, sh); P.ow = cdiv(W + 2 * padw, sw);
  P.scale = scale; P.slope = slope;
  P.out_bstride = out_batch_stride > 0 ? out_batch_stride : (int64_t)Ph * Pw * P.oh * P.ow;
  return k4::launch_fwd(P, static_cast<cudaStream_t>(stream));
}

extern "C" int ezb_matryoshka_fwd(const float* x1, const float* x2, int B, int G, int Cg, int H, int W, int P_, int stride,
                                  const int* dil, int D, float slope, float* out, int64_t out_batch_stride, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  EZB_REQUIRE(x1 && x2 && out && dil, "null pointer");
  EZB_REQUIRE(G > 0 && D > 0 && D <= k4::MAX_DIL, "bad group/dilation count G=%d D=%d (max %d dilations)", G, D, k4::MAX_DIL);
  if (int r = check_local_args(B, Cg, H, W, P_, P_, stride, stride, 0, 0, 1, 1)) return r;
  for (int d = 0; d < D; ++d) EZB_REQUIRE(dil[d] > 0, "dilation must be positive");
  if (D == 1) {  // DCVNet's fine scale (stride-2 features sampled every 4 pixels, one dilation of 1): local_corr_rt.cu
    const int64_t obs = out_batch_stride > 0 ? out_batch_stride : (int64_t)G * P_ * P_ * cdiv(H, stride) * cdiv(W, stride);
    if (local_corr_rt_strided_supported(x1, x2, out, B * G, Cg, H, W, P_, P_, stride, stride, 0, 0, dil[0], dil[0], obs))
      return local_corr_rt_strided_launch(x1, x2, B * G, G, Cg, H, W, P_, stride, 1.f, slope, out, obs, static_cast<cudaStream_t>(stream));
  }
  if (workspace != nullptr && local_corr_tc_supported(Cg, stride, stride, D)) {
    const int oh = cdiv(H, stride), ow = cdiv(W, stride);
    return local_corr_tc_launch(x1, x2, nullptr, B * G, G, Cg, H, W, P_, P_, stride, stride, 0, 0, D, dil, dil, 1.f, slope, out,
                                out_batch_stride > 0 ? out_batch_stride : (int64_t)D * G * P_ * P_ * oh * ow, workspace,
                                workspace_bytes, static_cast<cudaStream_t>(stream));
  }
  k4::Params P;
  P.in1 = x1; P.in2 = x2; P.out = out;
  P.BG = B * G; P.G = G; P.C = Cg; P.H = H; P.W = W;
  P.Ph = P_; P.Pw = P_; P.sh = stride; P.sw = stride; P.padh = 0; P.padw = 0;
  P.D = D;
  for (int d = 0; d < D; ++d) {
    EZB_REQUIRE(dil[d] > 0, "dilation must be positive");
    P.dph[d] = dil[d];
    P.dpw[d] = dil[d];
  }
  P.oh = cdiv(H, stride); P.ow = cdiv(W, stride);
  P.scale = 1.f; P.slope = slope;
  P.out_bstride = out_batch_stride > 0 ? out_batch_stride : (int64_t)D * G * P_ * P_ * P.oh * P.ow;
  return k4::launch_fwd(P, static_cast<cudaStream_t>(stream));
}

extern "C" int ezb_warp_local_corr_fwd(const float* in1, const float* in2, const float* flow, int B, int C, int H, int W, int Ph,
                                       int Pw, int sh, int sw, int padh, int padw, int dph, int dpw, float scale, float slope,
                                       float* out, int64_t out_batch_stride, void* workspace, size_t workspace_bytes,
                                       void* stream) {
  EZB_REQUIRE(in1 && in2 && flow && out, "null pointer");
  if (int r = check_local_args(B, C, H, W, Ph, Pw, sh, sw, padh, padw, dph, dpw)) return r;
  const bool tc_ok = workspace != nullptr && local_corr_tc_supported(C, sh, sw, 1);
  // Where the register-tiled kernel beats the tensor-core path (PWC-Net's finest levels), warping into the workspace first
  // (K5, one read + one write of in2) and correlating from there is faster than warping inside the tensor-core path's
  // conversion pass: 71 vs 162 us on the (8,32,112,256) level.  The workspace is always large enough (>= 4 bytes per element).
  if (tc_ok && local_corr_rt_supported(H, W, B, Ph, Pw, sh, sw, padh, padw, dph, dpw) && local_corr_rt_preferred(B, C, H, W, true) &&
      workspace_bytes >= (size_t)B * C * H * W * sizeof(float)) {
    float* warped = static_cast<float*>(workspace);
    if (int r = ezb_warp_fwd(in2, flow, B, C, H, W, warped, stream)) return r;
    return local_corr_rt_launch(in1, warped, B, 1, C, H, W, Ph, scale, slope, out,
                                out_batch_stride > 0 ? out_batch_stride : (int64_t)Ph * Pw * H * W, static_cast<cudaStream_t>(stream));
  }
  if (!tc_ok)
    return fail(EZB_ERR_UNSUPPORTED, "the warp-fused correlation needs a workspace and C <= 256; call ezb_warp_fwd and ezb_local_corr_fwd instead");
  const int oh = cdiv(H + 2 * padh, sh), ow = cdiv(W + 2 * padw, sw);
  return local_corr_tc_launch(in1, in2, flow, B, 1, C, H, W, Ph, Pw, sh, sw, padh, padw, 1, &dph, &dpw, scale, slope, out,
                              out_batch_stride > 0 ? out_batch_stride : (int64_t)Ph * Pw * oh * ow, workspace, workspace_bytes,
                              static_cast<cudaStream_t>(stream));
}

static int local_corr_bwd_simt(const float* in1, const float* in2, const float* dout, int B, int C, int H, int W, int Ph,
                               int Pw, int sh, int sw, int padh, int padw, int dph, int dpw, float scale,
                               int64_t dout_batch_stride, float* din1, float* din2, void* stream) {
  EZB_REQUIRE(in1 && in2 && dout && (din1 || din2), "null pointer");
  if (int r = check_local_args(B, C, H, W, Ph, Pw, sh, sw, padh, padw, dph, dpw)) return r;
  k4::BwdParams P;
  P.in1 = in1; P.in2 = in2; P.dout = dout; P.din1 = din1; P.din2 = din2;
  P.B = B; P.C = C; P.H = H; P.W = W; P.Ph = Ph; P.Pw = Pw; P.sh = sh; P.sw = sw;
  P.padh = padh; P.padw = padw; P.dph = dph; P.dpw = dpw;
  P.oh = cdiv(H + 2 * padh, sh); P.ow = cdiv(W + 2 * padw, sw);
  P.scale = scale;
  P.dout_bstride = dout_batch_stride > 0 ? dout_batch_stride : (int64_t)Ph * Pw * P.oh * P.ow;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // PWC-Net's shape (odd window <= 9, stride 1, no dilation / padding): register-tiled kernels (local_corr_rt.cu)
  if (local_corr_bwd_rt_supported(in1, in2, dout, din1, din2, B, C, H, W, Ph, Pw, sh, sw, padh, padw, dph, dpw, P.dout_bstride))
    return local_corr_bwd_rt_launch(in1, in2, dout, B, C, H, W, Ph, scale, P.dout_bstride, din1, din2, st);
  // DCVNet's fine scale (stride 4): strided gather kernels with the window products in registers (local_corr_rt.cu)
  if (local_corr_bwd_strided_supported(in1, in2, dout, din1, din2, B, C, H, W, Ph, Pw, sh, sw, padh, padw, dph, dpw))
    return local_corr_bwd_strided_launch(in1, in2, dout, B, 1, C, H, W, Ph, scale, P.dout_bstride, din1, din2, st);
  const long long total = (long long)B * C * H * W;
  const int blocks = (int)std::min<long long>(cdiv64(total, 256), 148 * 32);
  // stride 1: tiled kernels (products out of shared memory); otherwise the gather kernels
  const size_t tiled_smem = (size_t)k4::BT_CP * k4::BT_Y * (k4::BT_X + (Pw - 1) * dpw) * sizeof(float);
  const long long tiles = (long long)cdiv(W, k4::BT_X) * cdiv(H, k4::BT_Y);
  if (sh == 1 && sw == 1 && tiled_smem <= 200 * 1024 && tiles < (1ll << 31) && cdiv(C, k4::BT_C) <= 65535 && B <= 65535 &&
      getenv("EZB_LOCAL_CORR_BWD_GATHER") == nullptr) {
    const dim3 grid((unsigned)tiles, cdiv(C, k4::BT_C), B);
    if (din1) {
      EZB_CUDA_OK(cudaFuncSetAttribute(k4::local_corr_bwd_tiled_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled_smem));
      k4::local_corr_bwd_tiled_kernel<false><<<grid, k4::BT_Y * k4::BT_X, tiled_smem, st>>>(P);
      EZB_LAUNCH_OK();
    }
    if (din2) {
      EZB_CUDA_OK(cudaFuncSetAttribute(k4::local_corr_bwd_tiled_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tiled_smem));
      k4::local_corr_bwd_tiled_kernel<true><<<grid, k4::BT_Y * k4::BT_X, tiled_smem, st>>>(P);
      EZB_LAUNCH_OK();
    }
    return EZB_OK;
  }
  if (din1) { k4::local_corr_bwd_in1_kernel<<<blocks, 256, 0, st>>>(P); EZB_LAUNCH_OK(); }
  if (din2) { k4::local_corr_bwd_in2_kernel<<<blocks, 256, 0, st>>>(P); EZB_LAUNCH_OK(); }
  return EZB_OK;
}

extern "C" int ezb_local_corr_bwd(const float* in1, const float* in2, const float* dout, int B, int C, int H, int W, int Ph,
                                  int Pw, int sh, int sw, int padh, int padw, int dph, int dpw, float scale,
                                  int64_t dout_batch_stride, float* din1, float* din2, void* stream) {
  return local_corr_bwd_simt(in1, in2, dout, B, C, H, W, Ph, Pw, sh, sw, padh, padw, dph, dpw, scale, dout_batch_stride, din1, din2, stream);
}

extern "C" int ezb_local_corr_bwd_workspace_bytes(int BG, int C, int H, int W, size_t* bytes) {
  EZB_REQUIRE(BG > 0 && C > 0 && H > 0 && W > 0 && bytes, "bad arguments");
  // tensor-core operands + (K6) two gradient-sized scratch tensors for dilations the tensor-core path does not cover
  *bytes = local_corr_bwd_tc_workspace_bytes(BG, C, H, W) + 2 * ((size_t)BG * C * H * W * 4 + 256) + 1024;  // + alignment slack
  return EZB_OK;
}

// With a workspace the regular case (stride 1, no padding, square odd window <= 25, one isotropic dilation <= 8, C <= 256) runs
// on tcgen05 (fp16 operands, <= 1e-3 relative); everything else, or workspace == NULL, takes the exact fp32 CUDA-core kernels.
extern "C" int ezb_local_corr_bwd_ex(const float* in1, const float* in2, const float* dout, int B, int C, int H, int W, int Ph,
                                     int Pw, int sh, int sw, int padh, int padw, int dph, int dpw, float scale,
                                     int64_t dout_batch_stride, float* din1, float* din2, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  EZB_REQUIRE(in1 && in2 && dout && (din1 || din2), "null pointer");
  if (int r = check_local_args(B, C, H, W, Ph, Pw, sh, sw, padh, padw, dph, dpw)) return r;
  if (workspace != nullptr && local_corr_bwd_tc_supported(C, H, W, Ph, Pw, sh, sw, padh, padw, dph, dpw))
    return local_corr_bwd_tc_launch(in1, in2, dout, dout_batch_stride > 0 ? dout_batch_stride : (int64_t)Ph * Pw * H * W, B, 1, C, H, W, Ph, dph,
                                    scale, 0, din1, din2, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
  return local_corr_bwd_simt(in1, in2, dout, B, C, H, W, Ph, Pw, sh, sw, padh, padw, dph, dpw, scale, dout_batch_stride, din1, din2, stream);
}

namespace ezb {
namespace k4 {
__global__ void __launch_bounds__(256) add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] += src[i];
}
}  // namespace k4
}  // namespace ezb

// Adjoint of ezb_matryoshka_fwd (autograd of MatryoshkaDilatedCostVolume.forward, ezflow/similarity/learnable_cost.py:306-325):
// dx1, dx2 (B, G*Cg, H, W) from dout (B, D*G, P, P, oh, ow) (batch stride dout_batch_stride), summed over the D dilations in ONE
// call: dilations the tensor-core adjoint covers accumulate straight into the outputs, the others go through the fp32 kernels
// into scratch and are added.  workspace: ezb_local_corr_bwd_workspace_bytes(B*G, Cg, H, W) (NULL: fp32 kernels only, then
// the scratch is still required).
extern "C" int ezb_matryoshka_bwd(const float* x1, const float* x2, const float* dout, int B, int G, int Cg, int H, int W, int P_,
                                  int stride, const int* dil, int D, int64_t dout_batch_stride, float* dx1, float* dx2,
                                  void* workspace, size_t workspace_bytes, int use_tensor_cores, void* stream) {
  EZB_REQUIRE(x1 && x2 && dout && dx1 && dx2 && dil && workspace, "null pointer");
  EZB_REQUIRE(G > 0 && D > 0 && D <= k4::MAX_DIL, "bad group/dilation count G=%d D=%d (max %d dilations)", G, D, k4::MAX_DIL);
  if (int r = check_local_args(B, Cg, H, W, P_, P_, stride, stride, 0, 0, 1, 1)) return r;
  const int BG = B * G, oh = cdiv(H, stride), ow = cdiv(W, stride);
  const size_t tc_bytes = local_corr_bwd_tc_workspace_bytes(BG, Cg, H, W);
  const size_t grad_bytes = (size_t)BG * Cg * H * W * 4;
  size_t need = 0;
  ezb_local_corr_bwd_workspace_bytes(BG, Cg, H, W, &need);
  if (need > workspace_bytes) return fail(EZB_ERR_WORKSPACE, "workspace too small: need %zu, got %zu", need, workspace_bytes);
  uintptr_t p = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
  void* tc_ws = reinterpret_cast<void*>(p);
  p = (p + tc_bytes + 255) & ~uintptr_t(255);
  float* t1 = reinterpret_cast<float*>(p);
  p = (p + grad_bytes + 255) & ~uintptr_t(255);
  float* t2 = reinterpret_cast<float*>(p);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = (long long)BG * Cg * H * W;
  const int64_t bstride = dout_batch_stride > 0 ? dout_batch_stride : (int64_t)D * G * P_ * P_ * oh * ow;
  if (D == 1 && local_corr_bwd_strided_supported(x1, x2, dout, dx1, dx2, BG, Cg, H, W, P_, P_, stride, stride, 0, 0, dil[0], dil[0]))
    return local_corr_bwd_strided_launch(x1, x2, dout, BG, G, Cg, H, W, P_, 1.f, bstride, dx1, dx2, st);  // any group count
  bool first = true;
  for (int d = 0; d < D; ++d) {
    EZB_REQUIRE(dil[d] > 0, "dilation must be positive");
    const float* dd = dout + (long long)d * G * P_ * P_ * oh * ow;  // this dilation's (G, P, P, oh, ow) block of batch entry 0
    if (use_tensor_cores && local_corr_bwd_tc_supported(Cg, H, W, P_, P_, stride, stride, 0, 0, dil[d], dil[d])) {
      if (int r = local_corr_bwd_tc_launch(x1, x2, dd, bstride, BG, G, Cg, H, W, P_, dil[d], 1.f, first ? 0 : 1, dx1, dx2, tc_ws, tc_bytes, st)) return r;
    } else {
      // the fp32 kernels index dout per (batch*group) entry: present the block as BG entries when G == 1, else loop the groups
      for (int g = 0; g < G; ++g) {
        // entry (b, g): x views (B*G, Cg, H, W) are strided by G in the batch dimension -> one launch per group with B entries
        // is not expressible with a single batch stride for x; the fp32 path therefore requires G == 1 or falls back per (b, g)
        if (G == 1) break;
        return fail(EZB_ERR_UNSUPPORTED, "ezb_matryoshka_bwd: dilation %d with %d groups needs the per-dilation entry points", dil[d], G);
      }
      float* o1 = first ? dx1 : t1;
      float* o2 = first ? dx2 : t2;
      if (int r = local_corr_bwd_simt(x1, x2, dd, BG, Cg, H, W, P_, P_, stride, stride, 0, 0, dil[d], dil[d], 1.f, bstride, o1, o2, stream)) return r;
      if (!first) {
        const int blocks = (int)std::min<long long>(cdiv64(n, 256 * 4), 148 * 16);
        k4::add_inplace_kernel<<<blocks, 256, 0, st>>>(dx1, t1, n);
        k4::add_inplace_kernel<<<blocks, 256, 0, st>>>(dx2, t2, n);
        EZB_LAUNCH_OK();
      }
    }
    first = false;
  }
  return EZB_OK;
}

extern "C" int ezb_l2_normalize_fwd(const float* x, int B, int C, int HW, float eps, float* out, void* stream) {
  EZB_REQUIRE(x && out && B > 0 && C > 0 && HW > 0, "bad arguments");
  const long long total = (long long)B * HW;
  const int blocks = (int)std::min<long long>(cdiv64(total, 256), 148 * 32);
  k4::l2_normalize_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, C, HW, total, eps);
  EZB_LAUNCH_OK();
  return EZB_OK;
}
