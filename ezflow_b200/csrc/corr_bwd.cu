// K2b + K1b: adjoint of the correlation-pyramid build.
//
// Autograd of MutliScalePairwise4DCorr.__init__/.corr (reference ezflow/similarity/correlation/
// pairwise.py:26-41, :78-88; closed forms SURVEY.md App. A7):
//   fold   : dL_{l-1}[2y+a, 2x+b] += dL_l[y,x] / 4      (floor-pool adjoint; dropped rows/cols get 0)
//   gemms  : df1[b,c,q] = sum_t g[q,t] f2[b,c,t] / sqrt(C),  df2[b,c,t] = sum_q g[q,t] f1[b,c,q] / sqrt(C)
// With a workspace, C <= 256 and H*W a multiple of 4 everything runs on tcgen05 (kind::tf32, corr_bwd_tc.cu) without
// folding; the fold kernel and the fp32 shared-memory-tiled SIMT GEMMs below cover the remaining shapes.
#include "ezb_common.cuh"

#include <climits>
#include <cstdlib>

namespace ezb {
namespace k1b {

__global__ void fold_level_kernel(const float* __restrict__ up, float* __restrict__ down, long long rows, int ho, int wo,
                                  int pitch_up, long long qs_up, int pitch_down, long long qs_down) {
  const long long total = rows * ho * wo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % wo), y = (int)((i / wo) % ho);
    const long long q = i / ((long long)wo * ho);
    const float g = 0.25f * up[q * qs_up + (long long)y * pitch_up + x];
    float* d = down + q * qs_down + (long long)(2 * y) * pitch_down + 2 * x;
    d[0] += g;
    d[1] += g;
    d[pitch_down] += g;
    d[pitch_down + 1] += g;
  }
}

// ---- support of the gradient pyramid -------------------------------------------------------------------------------
// A query's gradient maps are non-zero only inside the lookup windows of the step's lookups (12 for RAFT), i.e. ~1 % of
// a 1080p pyramid.  ranges[b][T][l] = [t0, t1): the 256-target tiles of level l that any window of any query of the
// 256-query tile T can touch (same window arithmetic as corr_lookup_bwd_kernel).  The adjoint GEMMs skip everything
// outside (corr_bwd_tc.cu), and after the GEMMs only these ranges are zeroed again, so a gradient pyramid that starts
// all-zero is all-zero again for the next step without a 5.6 GB memset per pair.
constexpr int RG_MAX_LOOKUPS = 32, RG_LEVELS = 4;
struct RangeParams {
  const float* coords[RG_MAX_LOOKUPS];  // (B,2,N) each
  int n, B, N, L, r, nT;
  int h[RG_LEVELS], w[RG_LEVELS];
  int2* ranges;  // [B][nT][RG_LEVELS], pre-set to (INT_MAX, 0)
};
__device__ __forceinline__ int floor_clamped(float v) {
  // floorf with the same saturation as the lookup kernels (NaN / huge coordinates fall outside every level)
  if (!(v > -16777216.f)) return -(1 << 24);
  if (!(v < 16777216.f)) return 1 << 24;
  return (int)floorf(v);
}
__global__ void __launch_bounds__(256) grad_ranges_kernel(const RangeParams P) {
  const int b = blockIdx.y, T = blockIdx.x, q = T * 256 + threadIdx.x;
  int lo[RG_LEVELS], hi[RG_LEVELS];
#pragma unroll
  for (int l = 0; l < RG_LEVELS; ++l) { lo[l] = INT_MAX; hi[l] = -1; }
  if (q < P.N) {
    const int D = 2 * P.r + 2;
    for (int i = 0; i < P.n; ++i) {
      const float cx0 = __ldg(P.coords[i] + (long long)b * 2 * P.N + q), cy0 = __ldg(P.coords[i] + (long long)b * 2 * P.N + P.N + q);
#pragma unroll
      for (int l = 0; l < RG_LEVELS; ++l) {
        if (l >= P.L) break;
        const float inv = 1.0f / (float)(1 << l);
        const int hl = P.h[l], wl = P.w[l];
        const int xs = max(-D, min(wl, floor_clamped(cx0 * inv) - P.r)), ys = max(-D, min(hl, floor_clamped(cy0 * inv) - P.r));
        const int x0 = max(xs, 0), x1 = min(xs + D, wl) - 1, y0 = max(ys, 0), y1 = min(ys + D, hl) - 1;  // window clipped to the level
        if (x0 > x1 || y0 > y1) continue;
        lo[l] = min(lo[l], y0 * wl + x0);
        hi[l] = max(hi[l], y1 * wl + x1);
      }
    }
  }
  __shared__ int s_lo[RG_LEVELS], s_hi[RG_LEVELS];
  if (threadIdx.x < RG_LEVELS) { s_lo[threadIdx.x] = INT_MAX; s_hi[threadIdx.x] = -1; }
  __syncthreads();
#pragma unroll
  for (int l = 0; l < RG_LEVELS; ++l) {
    const int a = __reduce_min_sync(0xffffffffu, lo[l]), c = __reduce_max_sync(0xffffffffu, hi[l]);
    if ((threadIdx.x & 31) == 0) { atomicMin(&s_lo[l], a); atomicMax(&s_hi[l], c); }
  }
  __syncthreads();
  if (threadIdx.x < RG_LEVELS && s_hi[threadIdx.x] >= 0) {
    int2* r = P.ranges + ((long long)b * P.nT + T) * RG_LEVELS + threadIdx.x;
    atomicMin(&r->x, s_lo[threadIdx.x] >> 8);
    atomicMax(&r->y, (s_hi[threadIdx.x] >> 8) + 1);
  }
}
__global__ void __launch_bounds__(256) grad_ranges_init_kernel(int2* ranges, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) ranges[i] = make_int2(INT_MAX, 0);
}
// zero rows q of tile T over the columns of its range, level by level (blockIdx.x = T, blockIdx.y = b, blockIdx.z = row slice)
struct RezeroParams {
  float* dlvl[RG_LEVELS];
  long long qstride[RG_LEVELS];
  int nl[RG_LEVELS];
  const int2* ranges;
  int N, L, nT;
};
__global__ void __launch_bounds__(256) grad_rezero_kernel(const RezeroParams P) {
  const int b = blockIdx.y, T = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int l = 0; l < P.L; ++l) {
    const int2 r = __ldg(P.ranges + ((long long)b * P.nT + T) * RG_LEVELS + l);
    const int c0 = r.x == INT_MAX ? 0 : r.x * 256, c1 = min(P.nl[l], r.y * 256);  // columns, multiples of 4 except the level's end
    if (c1 <= c0) continue;
    const int c1v = c0 + ((c1 - c0) & ~3);
    // a warp per row (qstride and c0 are multiples of 4 floats: 16-byte stores), rows dealt over blockIdx.z
    for (int row = blockIdx.z * 8 + warp; row < 256; row += gridDim.z * 8) {
      const int q = T * 256 + row;
      if (q >= P.N) break;
      float* d = P.dlvl[l] + ((long long)b * P.N + q) * P.qstride[l];
      for (int c = c0 + 4 * lane; c < c1v; c += 128) *reinterpret_cast<float4*>(d + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      if (lane < c1 - c1v) d[c1v + lane] = 0.f;
    }
  }
}

constexpr int TM = 64, TN = 64, TK = 16;

struct GView {  // level-0 gradient volume g[b][q][t], t = y*W + x stored with a row pitch
  const float* p;
  int W, pitch;
  long long qstride;
  __device__ __forceinline__ float at(long long row, int t) const {
    const int y = t / W, x = t - y * W;
    return __ldg(p + row * qstride + (long long)y * pitch + x);
  }
};

// out[b][c][n] = alpha * sum_k F[b][c][k] * Gm(k, n)   with  Gm(k,n) = TRANSPOSED ? g[n][k] : g[k][n]
//   TRANSPOSED = true  -> df1 (n = q, k = t)        TRANSPOSED = false -> df2 (n = t, k = q)
template <bool TRANSPOSED>
__global__ void __launch_bounds__(256) corr_bwd_gemm_kernel(const float* __restrict__ F, GView G, float* __restrict__ out, int C,
                                                            int N, float alpha) {
  __shared__ float sF[TK][TM + 1];
  __shared__ float sG[TK][TN + 1];
  const int b = blockIdx.z, c0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const float* Fb = F + (long long)b * C * N;
  const long long grow0 = (long long)b * N;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < N; k0 += TK) {
    for (int i = tid; i < TM * TK; i += 256) {
      const int kk = i % TK, m = i / TK;
      const int c = c0 + m, k = k0 + kk;
      sF[kk][m] = (c < C && k < N) ? __ldg(Fb + (long long)c * N + k) : 0.f;
    }
    for (int i = tid; i < TN * TK; i += 256) {
      int kk, n;
      if (TRANSPOSED) { kk = i % TK; n = i / TK; } else { n = i % TN; kk = i / TN; }
      const int k = k0 + kk, nn = n0 + n;
      float v = 0.f;
      if (k < N && nn < N) v = TRANSPOSED ? G.at(grow0 + nn, k) : G.at(grow0 + k, nn);
      sG[kk][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], g[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sF[kk][ty * 4 + i]; g[i] = sG[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], g[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty * 4 + i;
    if (c >= C) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) out[((long long)b * C + c) * N + n] = acc[i][j] * alpha;
    }
  }
}

}  // namespace k1b
}  // namespace ezb

namespace ezb {
// corr_bwd_tc.cu
bool corr_bwd_tc_supported(int C, int N);
size_t corr_bwd_tc_workspace_bytes(int B, int C, int H, int W, int L);
int corr_bwd_tc_launch(float* const* dlvl, const long long* qstride, const float* fmap1, const float* fmap2, int B, int C, int H,
                       int W, int L, float alpha, float* dfmap1, float* dfmap2, void* workspace, size_t workspace_bytes,
                       const int32_t* ranges, cudaStream_t st);
}  // namespace ezb

using namespace ezb;

extern "C" int ezb_corr_pyramid_bwd_workspace_bytes(int B, int C, int H, int W, int L, size_t* bytes) {
  EZB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && L >= 1 && L <= EZB_MAX_LEVELS && bytes, "bad arguments");
  *bytes = (L <= 4 && corr_bwd_tc_supported(C, H * W)) ? corr_bwd_tc_workspace_bytes(B, C, H, W, L) : 0;
  return EZB_OK;
}

extern "C" int ezb_corr_pyramid_bwd_ex(float* const* dlvl_ptrs, const float* fmap1, const float* fmap2, int B, int C, int H,
                                       int W, int L, float* dfmap1, float* dfmap2, void* workspace, size_t workspace_bytes,
                                       const int32_t* ranges, void* stream) {
  EZB_REQUIRE(dlvl_ptrs && fmap1 && fmap2 && (dfmap1 || dfmap2), "null pointer");
  EZB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && L >= 1 && L <= EZB_MAX_LEVELS, "bad shape B=%d C=%d H=%d W=%d L=%d", B, C,
              H, W, L);
  LevelGeom g[EZB_MAX_LEVELS];
  EZB_REQUIRE(pyramid_geom(H, W, L, EZB_F32, g), "feature map %dx%d too small for %d pyramid levels", H, W, L);
  for (int l = 0; l < L; ++l) EZB_REQUIRE(dlvl_ptrs[l], "level %d gradient pointer is null", l);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int N = H * W;
  const long long rows = (long long)B * N;
  const float alpha = 1.0f / sqrtf((float)C);
  if (workspace != nullptr && L <= 4 && corr_bwd_tc_supported(C, N) && getenv("EZB_BWD_SIMT") == nullptr) {
    // tensor cores, no fold: the K loop of the df1 GEMM runs over the targets of all levels (corr_bwd_tc.cu)
    long long qs[EZB_MAX_LEVELS];
    for (int l = 0; l < L; ++l) qs[l] = g[l].qstride;
    return corr_bwd_tc_launch(dlvl_ptrs, qs, fmap1, fmap2, B, C, H, W, L, alpha, dfmap1, dfmap2, workspace, workspace_bytes, ranges, st);
  }
  if (ranges != nullptr)
    return fail(EZB_ERR_UNSUPPORTED, "gradient ranges need the tensor-core adjoint (ezb_corr_grad_ranges_supported)");
  for (int l = L - 1; l >= 1; --l) {
    const long long total = rows * g[l].h * g[l].w;
    const int blocks = (int)std::min<long long>(cdiv64(total, 256), 148 * 32);
    k1b::fold_level_kernel<<<blocks, 256, 0, st>>>(dlvl_ptrs[l], dlvl_ptrs[l - 1], rows, g[l].h, g[l].w, g[l].pitch, g[l].qstride,
                                                   g[l - 1].pitch, g[l - 1].qstride);
    EZB_LAUNCH_OK();
  }
  k1b::GView G{dlvl_ptrs[0], W, g[0].pitch, g[0].qstride};
  dim3 grid(cdiv(N, k1b::TN), cdiv(C, k1b::TM), B);
  if (dfmap1) {
    k1b::corr_bwd_gemm_kernel<true><<<grid, 256, 0, st>>>(fmap2, G, dfmap1, C, N, alpha);
    EZB_LAUNCH_OK();
  }
  if (dfmap2) {
    k1b::corr_bwd_gemm_kernel<false><<<grid, 256, 0, st>>>(fmap1, G, dfmap2, C, N, alpha);
    EZB_LAUNCH_OK();
  }
  return EZB_OK;
}

extern "C" int ezb_corr_pyramid_bwd(float* const* dlvl_ptrs, const float* fmap1, const float* fmap2, int B, int C, int H,
                                    int W, int L, float* dfmap1, float* dfmap2, void* workspace, size_t workspace_bytes,
                                    void* stream) {
  return ezb_corr_pyramid_bwd_ex(dlvl_ptrs, fmap1, fmap2, B, C, H, W, L, dfmap1, dfmap2, workspace, workspace_bytes, nullptr, stream);
}

extern "C" int ezb_corr_grad_ranges_supported(int C, int H, int W, int L) {
  return (L >= 1 && L <= 4 && C > 0 && H > 0 && W > 0 && corr_bwd_tc_supported(C, H * W) && getenv("EZB_BWD_SIMT") == nullptr &&
          getenv("EZB_BWD_DENSE") == nullptr) ? 1 : 0;
}

extern "C" int ezb_corr_grad_ranges_count(int B, int H, int W, int64_t* n_int32) {
  EZB_REQUIRE(B > 0 && H > 0 && W > 0 && n_int32, "bad arguments");
  *n_int32 = (int64_t)B * cdiv(H * W, 256) * k1b::RG_LEVELS * 2;
  return EZB_OK;
}

extern "C" int ezb_corr_grad_ranges(const float* const* coords_ptrs, int n_lookups, int B, int H, int W, int L, int radius,
                                    int32_t* ranges, void* stream) {
  EZB_REQUIRE(coords_ptrs && ranges && n_lookups >= 0, "bad arguments");
  EZB_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && L >= 1 && L <= k1b::RG_LEVELS && radius >= 0, "bad shape B=%d H=%d W=%d L=%d r=%d", B, H, W,
              L, radius);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int N = H * W, nT = cdiv(N, 256);
  const long long n = (long long)B * nT * k1b::RG_LEVELS;
  k1b::grad_ranges_init_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, st>>>(reinterpret_cast<int2*>(ranges), n);
  EZB_LAUNCH_OK();
  for (int i0 = 0; i0 < n_lookups; i0 += k1b::RG_MAX_LOOKUPS) {
    k1b::RangeParams P;
    P.n = std::min(k1b::RG_MAX_LOOKUPS, n_lookups - i0);
    for (int i = 0; i < P.n; ++i) {
      EZB_REQUIRE(coords_ptrs[i0 + i], "coords %d is null", i0 + i);
      P.coords[i] = coords_ptrs[i0 + i];
    }
    P.B = B; P.N = N; P.L = L; P.r = radius; P.nT = nT;
    for (int l = 0; l < k1b::RG_LEVELS; ++l) { P.h[l] = l < L ? H >> l : 0; P.w[l] = l < L ? W >> l : 0; }
    P.ranges = reinterpret_cast<int2*>(ranges);
    k1b::grad_ranges_kernel<<<dim3(nT, B), 256, 0, st>>>(P);
    EZB_LAUNCH_OK();
  }
  return EZB_OK;
}

extern "C" int ezb_corr_grad_rezero(float* const* dlvl_ptrs, const int32_t* ranges, int B, int H, int W, int L, void* stream) {
  EZB_REQUIRE(dlvl_ptrs && ranges, "null pointer");
  EZB_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && L >= 1 && L <= k1b::RG_LEVELS, "bad shape B=%d H=%d W=%d L=%d", B, H, W, L);
  LevelGeom g[EZB_MAX_LEVELS];
  EZB_REQUIRE(pyramid_geom(H, W, L, EZB_F32, g), "feature map %dx%d too small for %d pyramid levels", H, W, L);
  k1b::RezeroParams P;
  for (int l = 0; l < k1b::RG_LEVELS; ++l) {
    P.dlvl[l] = l < L ? dlvl_ptrs[l] : nullptr;
    P.qstride[l] = l < L ? g[l].qstride : 0;
    P.nl[l] = l < L ? g[l].h * g[l].w : 0;
    if (l < L) EZB_REQUIRE(dlvl_ptrs[l] && (reinterpret_cast<uintptr_t>(dlvl_ptrs[l]) & 15) == 0, "level %d gradient pointer is null or misaligned", l);
  }
  P.ranges = reinterpret_cast<const int2*>(ranges);
  P.N = H * W; P.L = L; P.nT = cdiv(P.N, 256);
  k1b::grad_rezero_kernel<<<dim3(P.nT, B, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(P);
  EZB_LAUNCH_OK();
  return EZB_OK;
}
