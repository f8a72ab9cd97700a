// K7: convex-combination flow upsampling, the consumer of every RAFT / DCVNet iteration's coarse flow.
//
// Replaces convex_upsample_flow (reference ezflow/utils/resampling.py:112-141; called 12x per RAFT forward at
// models/raft.py:127 and by DCVNet at decoder/dilated_flow_stack_filter.py:219):
//   mask = softmax over k of mask_logits.view(N, 1, 9, s, s, H, W)            (k = ky*3 + kx)
//   out[n, c, s*h + i, s*w + j] = sum_k mask[n, k, i, j, h, w] * flow[n, c, h + ky - 1, w + kx - 1]   (zero padded)
// HBM-bound: 9*s*s*4 B of logits read and C*s*s*4 B written per coarse pixel (2.6 KB at s = 8, C = 2).
//
// Tiling: one block = one coarse row segment of 32 pixels.  lane = coarse x (so every logit plane is read in 128-byte
// lines), warp g sweeps the sub-positions p = i*s + j = g, g+8, ...; the s*s x 32 results of a channel are transposed
// through shared memory so that the fine rows leave as contiguous 32*s-float runs.
#include <algorithm>

#include "ezb_common.cuh"

namespace ezb {
namespace k7 {

constexpr int TW = 32;       // coarse pixels per block (one per lane)
constexpr int NWARP = 8;
constexpr int NTHREADS = NWARP * 32;

__host__ __device__ inline int col_stride(int s) {
  // [j][w] staging rows: the stride makes both the lane = w writes and the lane = fine-x reads conflict-free for s = 2^k
  return (s <= 32 && (s & (s - 1)) == 0) ? TW + TW / s : TW + 1;
}

template <int C>
__device__ __forceinline__ void load_neighbours(const float* fl_s, int lane, float (&nb)[C][9]) {
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) nb[c][ky * 3 + kx] = fl_s[(c * 3 + ky) * (TW + 2) + lane + kx];
}

// flow rows h-1..h+1, columns w0-1..w0+TW of every channel, zero outside the image
template <int C>
__device__ __forceinline__ void stage_flow(const float* __restrict__ flow, float* fl_s, int n, int h, int w0, int H, int W) {
  for (int idx = threadIdx.x; idx < C * 3 * (TW + 2); idx += NTHREADS) {
    const int x = idx % (TW + 2), r = idx / (TW + 2);
    const int ky = r % 3, c = r / 3;
    const int yy = h + ky - 1, xx = w0 + x - 1;
    float v = 0.f;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(flow + (((long long)n * C + c) * H + yy) * W + xx);
    fl_s[idx] = v;
  }
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// softmax over the 9 taps.  exp(l - m) = 2^((l - m) log2 e) through MUFU.EX2 (2 ulp) and an approximate reciprocal of the
// sum (which lies in [1, 9]): relative error ~1e-6 for |logit| up to ~16, far inside the 2e-5 parity gate, and a quarter of
// the instructions of expf + IEEE division, which this kernel would otherwise be issue-bound on.
__device__ __forceinline__ void softmax9(float (&l)[9]) {
  constexpr float LOG2E = 1.4426950408889634f;
  float m = l[0];
#pragma unroll
  for (int k = 1; k < 9; ++k) m = fmaxf(m, l[k]);
  const float ml = m * LOG2E;
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    l[k] = ex2_approx(fmaf(l[k], LOG2E, -ml));
    sum += l[k];
  }
  const float inv = __fdividef(1.0f, sum);
#pragma unroll
  for (int k = 0; k < 9; ++k) l[k] *= inv;
}

// S > 0: out_stride known at compile time (8 for RAFT and DCVNet) so that the index arithmetic folds into shifts; S == 0: any
template <int C, int S>
__global__ void __launch_bounds__(NTHREADS, C <= 2 ? 4 : 2) convex_upsample_fwd_kernel(const float* __restrict__ flow,
                                                                       const float* __restrict__ logits,
                                                                       float* __restrict__ out, int N, int H, int W, int s_rt,
                                                                       int wtiles) {
  extern __shared__ float smem[];
  const int s = S > 0 ? S : s_rt;
  const int ss = s * s, cs = col_stride(s);
  float* fl_s = smem;                        // [C][3][TW+2]
  float* out_s = smem + C * 3 * (TW + 2);    // [C][s (i)][s (j)][cs]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long bid = blockIdx.x;
  const int wt = (int)(bid % wtiles);
  bid /= wtiles;
  const int h = (int)(bid % H), n = (int)(bid / H);
  const int w0 = wt * TW, w = w0 + lane;
  const long long HW = (long long)H * W;

  stage_flow<C>(flow, fl_s, n, h, w0, H, W);
  __syncthreads();
  float nb[C][9];
  load_neighbours<C>(fl_s, lane, nb);

  // lanes beyond the image edge read the last valid pixel (their results are never stored): no predicated loads
  // (offsets inside one image fit 32 bits — checked by the launcher — which keeps the address arithmetic to one IMAD per load)
  const float* lg = logits + (long long)n * 9 * ss * HW + (long long)h * W + min(w, W - 1);
  const unsigned hw = (unsigned)HW, ps = (unsigned)ss * hw;  // stride between sub-positions / taps
  // two sub-positions per trip: 18 independent loads in flight per thread
  for (int p = warp; p < ss; p += 2 * NWARP) {
    const bool has2 = p + NWARP < ss;
    const int p2 = has2 ? p + NWARP : p;
    const unsigned oa = (unsigned)p * hw, ob = (unsigned)p2 * hw;
    float a[9], b[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      a[k] = __ldcs(lg + (oa + k * ps));
      b[k] = __ldcs(lg + (ob + k * ps));
    }
    softmax9(a);
    softmax9(b);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float ra = 0.f, rb = 0.f;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        ra = fmaf(a[k], nb[c][k], ra);
        rb = fmaf(b[k], nb[c][k], rb);
      }
      out_s[(c * ss + p) * cs + lane] = ra;
      if (has2) out_s[(c * ss + p2) * cs + lane] = rb;
    }
  }
  __syncthreads();
  // fine rows: (c, i) -> 32*s contiguous floats; thread = fine column
  const int roww = TW * s, xmax = min(roww, (W - w0) * s);
  const unsigned Ws = (unsigned)W * s, plane = (unsigned)H * s * Ws;
  float* on = out + (long long)n * C * plane;
  for (int x = threadIdx.x; x < xmax; x += NTHREADS) {
    const float* src = out_s + (x % s) * cs + x / s;
    const unsigned o0 = (unsigned)h * s * Ws + (unsigned)w0 * s + x;
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll(S > 0 ? S : 1)
      for (int i = 0; i < s; ++i) __stcs(on + (o0 + c * plane + i * Ws), src[(c * s + i) * s * cs]);
    }
  }
}

// Backward.  With p = softmax(logits), g_k = sum_c dout_c * nb_c,k:
//   dlogit_k = p_k * (g_k - sum_m p_m g_m)
//   dnb[c][k] (h, w) = sum_{i,j} p_k * dout_c           -> part[n][c][k][h][w]   (deterministic block-level reduction)
// and a second pass gathers dflow[c][y][x] = sum_k part[c][k][y - ky + 1][x - kx + 1].
template <int C, int S>
__global__ void __launch_bounds__(NTHREADS, C <= 2 ? 3 : 1) convex_upsample_bwd_kernel(const float* __restrict__ flow,
                                                                       const float* __restrict__ logits,
                                                                       const float* __restrict__ dout, float* __restrict__ dlogits,
                                                                       float* __restrict__ part, int N, int H, int W, int s_rt,
                                                                       int wtiles) {
  extern __shared__ float smem[];
  const int s = S > 0 ? S : s_rt;
  const int ss = s * s, cs = col_stride(s);
  float* fl_s = smem;                        // [C][3][TW+2]
  float* do_s = smem + C * 3 * (TW + 2);    // [C][s][s][cs]; reused as the [NWARP][C*9][32] reduction buffer
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long bid = blockIdx.x;
  const int wt = (int)(bid % wtiles);
  bid /= wtiles;
  const int h = (int)(bid % H), n = (int)(bid / H);
  const int w0 = wt * TW, w = w0 + lane;
  const long long HW = (long long)H * W;

  stage_flow<C>(flow, fl_s, n, h, w0, H, W);
  {
    const int roww = TW * s, xmax = min(roww, (W - w0) * s);
    const unsigned Ws = (unsigned)W * s, plane = (unsigned)H * s * Ws;
    const float* dn = dout + (long long)n * C * plane;
    for (int x = threadIdx.x; x < roww; x += NTHREADS) {
      float* dstp = do_s + (x % s) * cs + x / s;
      const unsigned o0 = (unsigned)h * s * Ws + (unsigned)w0 * s + x;
#pragma unroll
      for (int c = 0; c < C; ++c) {
#pragma unroll(S > 0 ? S : 1)
        for (int i = 0; i < s; ++i) dstp[(c * s + i) * s * cs] = x < xmax ? __ldcs(dn + (o0 + c * plane + i * Ws)) : 0.f;
      }
    }
  }
  __syncthreads();
  float nb[C][9];
  load_neighbours<C>(fl_s, lane, nb);
  float acc[C][9];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[c][k] = 0.f;

  const bool live = w < W;
  const long long base = (long long)n * 9 * ss * HW + (long long)h * W + min(w, W - 1);
  const float* lg = logits + base;
  const unsigned hw = (unsigned)HW, ps = (unsigned)ss * hw;
  for (int p = warp; p < ss; p += NWARP) {
    float a[9];
    const unsigned oa = (unsigned)p * hw;
#pragma unroll
    for (int k = 0; k < 9; ++k) a[k] = __ldcs(lg + (oa + k * ps));
    softmax9(a);
    float d[C];
#pragma unroll
    for (int c = 0; c < C; ++c) d[c] = do_s[(c * ss + p) * cs + lane];
    float g[9], dot = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      g[k] = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        g[k] = fmaf(d[c], nb[c][k], g[k]);
        acc[c][k] = fmaf(a[k], d[c], acc[c][k]);
      }
      dot = fmaf(a[k], g[k], dot);
    }
    if (live && dlogits) {
      float* pd = dlogits + base;
#pragma unroll
      for (int k = 0; k < 9; ++k) __stcs(pd + (oa + k * ps), a[k] * (g[k] - dot));
    }
  }
  if (!part) return;
  __syncthreads();  // do_s is dead; reuse it
  float* red = do_s;  // [NWARP][C*9][32]
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int k = 0; k < 9; ++k) red[(warp * C * 9 + c * 9 + k) * 32 + lane] = acc[c][k];
  __syncthreads();
  for (int idx = threadIdx.x; idx < C * 9 * 32; idx += NTHREADS) {
    const int l = idx & 31, ck = idx >> 5;
    if (w0 + l >= W) continue;
    float sum = 0.f;
#pragma unroll
    for (int g = 0; g < NWARP; ++g) sum += red[(g * C * 9 + ck) * 32 + l];
    part[(((long long)n * C * 9 + ck) * H + h) * W + w0 + l] = sum;
  }
}

__global__ void __launch_bounds__(256) convex_upsample_gather_kernel(const float* __restrict__ part, float* __restrict__ dflow,
                                                                     long long total, int H, int W) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % W);
  const int y = (int)((i / W) % H);
  const long long nc = i / ((long long)W * H);
  float sum = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yy = y - ky + 1, xx = x - kx + 1;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) sum += __ldg(part + ((nc * 9 + ky * 3 + kx) * H + yy) * W + xx);
    }
  dflow[i] = sum;
}

inline size_t smem_bytes(int C, int s, bool bwd) {
  size_t stage = (size_t)C * s * s * col_stride(s);
  if (bwd) stage = std::max(stage, (size_t)NWARP * C * 9 * 32);
  return ((size_t)C * 3 * (TW + 2) + stage) * sizeof(float);
}

template <int C, int S>
int launch_fwd_s(const float* flow, const float* logits, float* out, int N, int H, int W, int s, cudaStream_t st) {
  const size_t sm = smem_bytes(C, s, false);
  EZB_REQUIRE(sm <= 200 * 1024, "out_stride %d too large for the staging tile", s);
  if (sm > 48 * 1024)
    EZB_CUDA_OK(cudaFuncSetAttribute(convex_upsample_fwd_kernel<C, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  const int wtiles = cdiv(W, TW);
  const long long blocks = (long long)N * H * wtiles;
  EZB_REQUIRE(blocks < (1ll << 31) && 9ll * s * s * H * W < (1ll << 31), "tensor too large for one launch");
  convex_upsample_fwd_kernel<C, S><<<(unsigned)blocks, NTHREADS, sm, st>>>(flow, logits, out, N, H, W, s, wtiles);
  EZB_LAUNCH_OK();
  return EZB_OK;
}

template <int C>
int launch_fwd(const float* flow, const float* logits, float* out, int N, int H, int W, int s, cudaStream_t st) {
  if (s == 8) return launch_fwd_s<C, 8>(flow, logits, out, N, H, W, s, st);
  if (s == 4) return launch_fwd_s<C, 4>(flow, logits, out, N, H, W, s, st);
  return launch_fwd_s<C, 0>(flow, logits, out, N, H, W, s, st);
}

template <int C, int S>
int launch_bwd_main(const float* flow, const float* logits, const float* dout, float* dlogits, float* part, int N, int H, int W, int s,
                    cudaStream_t st) {
  const size_t sm = smem_bytes(C, s, true);
  EZB_REQUIRE(sm <= 200 * 1024, "out_stride %d too large for the staging tile", s);
  if (sm > 48 * 1024)
    EZB_CUDA_OK(cudaFuncSetAttribute(convex_upsample_bwd_kernel<C, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  const int wtiles = cdiv(W, TW);
  const long long blocks = (long long)N * H * wtiles;
  EZB_REQUIRE(blocks < (1ll << 31) && 9ll * s * s * H * W < (1ll << 31), "tensor too large for one launch");
  convex_upsample_bwd_kernel<C, S><<<(unsigned)blocks, NTHREADS, sm, st>>>(flow, logits, dout, dlogits, part, N, H, W, s, wtiles);
  EZB_LAUNCH_OK();
  return EZB_OK;
}

template <int C>
int launch_bwd(const float* flow, const float* logits, const float* dout, float* dflow, float* dlogits, float* part, int N, int H,
               int W, int s, cudaStream_t st) {
  float* p = dflow ? part : nullptr;
  int rc;
  if (s == 8) rc = launch_bwd_main<C, 8>(flow, logits, dout, dlogits, p, N, H, W, s, st);
  else if (s == 4) rc = launch_bwd_main<C, 4>(flow, logits, dout, dlogits, p, N, H, W, s, st);
  else rc = launch_bwd_main<C, 0>(flow, logits, dout, dlogits, p, N, H, W, s, st);
  if (rc) return rc;
  if (dflow) {
    const long long total = (long long)N * C * H * W;
    convex_upsample_gather_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, st>>>(part, dflow, total, H, W);
    EZB_LAUNCH_OK();
  }
  return EZB_OK;
}

}  // namespace k7
}  // namespace ezb

using namespace ezb;

extern "C" int ezb_convex_upsample_fwd(const float* flow, const float* mask_logits, int N, int C, int H, int W, int out_stride,
                                       float* out, void* stream) {
  EZB_REQUIRE(flow && mask_logits && out, "null pointer");
  EZB_REQUIRE(N > 0 && H > 0 && W > 0 && out_stride > 0, "bad shape N=%d H=%d W=%d out_stride=%d", N, H, W, out_stride);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (C) {
    case 1: return k7::launch_fwd<1>(flow, mask_logits, out, N, H, W, out_stride, st);
    case 2: return k7::launch_fwd<2>(flow, mask_logits, out, N, H, W, out_stride, st);
    case 3: return k7::launch_fwd<3>(flow, mask_logits, out, N, H, W, out_stride, st);
    case 4: return k7::launch_fwd<4>(flow, mask_logits, out, N, H, W, out_stride, st);
    default: return fail(EZB_ERR_UNSUPPORTED, "convex upsampling supports 1..4 flow channels, got %d", C);
  }
}

extern "C" int ezb_convex_upsample_bwd_workspace_bytes(int N, int C, int H, int W, size_t* bytes) {
  EZB_REQUIRE(bytes, "null pointer");
  EZB_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, "bad shape N=%d C=%d H=%d W=%d", N, C, H, W);
  *bytes = (size_t)N * C * 9 * H * W * sizeof(float);
  return EZB_OK;
}

extern "C" int ezb_convex_upsample_bwd(const float* flow, const float* mask_logits, const float* dout, int N, int C, int H, int W,
                                       int out_stride, float* dflow, float* dmask_logits, void* workspace, size_t workspace_bytes,
                                       void* stream) {
  EZB_REQUIRE(flow && mask_logits && dout && (dflow || dmask_logits), "null pointer");
  EZB_REQUIRE(N > 0 && H > 0 && W > 0 && out_stride > 0, "bad shape N=%d H=%d W=%d out_stride=%d", N, H, W, out_stride);
  if (dflow) {
    size_t need = 0;
    if (int rc = ezb_convex_upsample_bwd_workspace_bytes(N, C, H, W, &need)) return rc;
    EZB_REQUIRE(workspace && workspace_bytes >= need, "workspace too small: %zu < %zu bytes", workspace_bytes, need);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* part = static_cast<float*>(workspace);
  switch (C) {
    case 1: return k7::launch_bwd<1>(flow, mask_logits, dout, dflow, dmask_logits, part, N, H, W, out_stride, st);
    case 2: return k7::launch_bwd<2>(flow, mask_logits, dout, dflow, dmask_logits, part, N, H, W, out_stride, st);
    case 3: return k7::launch_bwd<3>(flow, mask_logits, dout, dflow, dmask_logits, part, N, H, W, out_stride, st);
    case 4: return k7::launch_bwd<4>(flow, mask_logits, dout, dflow, dmask_logits, part, N, H, W, out_stride, st);
    default: return fail(EZB_ERR_UNSUPPORTED, "convex upsampling supports 1..4 flow channels, got %d", C);
  }
}
