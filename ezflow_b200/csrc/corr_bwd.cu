// K2b + K1b: adjoint of the correlation-pyramid build.
//
// Autograd of MutliScalePairwise4DCorr.__init__/.corr (reference ezflow/similarity/correlation/
// pairwise.py:26-41, :78-88; closed forms SURVEY.md App. A7):
//   fold   : dL_{l-1}[2y+a, 2x+b] += dL_l[y,x] / 4      (floor-pool adjoint; dropped rows/cols get 0)
//   gemms  : df1[b,c,q] = sum_t g[q,t] f2[b,c,t] / sqrt(C),  df2[b,c,t] = sum_q g[q,t] f1[b,c,q] / sqrt(C)
// With a workspace, C <= 256 and H*W a multiple of 4 everything runs on tcgen05 (kind::tf32, corr_bwd_tc.cu) without
// folding; the fold kernel and the fp32 shared-memory-tiled SIMT GEMMs below cover the remaining shapes.
#include "ezb_common.cuh"

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
                       cudaStream_t st);
}  // namespace ezb

using namespace ezb;

extern "C" int ezb_corr_pyramid_bwd_workspace_bytes(int B, int C, int H, int W, int L, size_t* bytes) {
  EZB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && L >= 1 && L <= EZB_MAX_LEVELS && bytes, "bad arguments");
  *bytes = (L <= 4 && corr_bwd_tc_supported(C, H * W)) ? corr_bwd_tc_workspace_bytes(B, C, H, W, L) : 0;
  return EZB_OK;
}

extern "C" int ezb_corr_pyramid_bwd(float* const* dlvl_ptrs, const float* fmap1, const float* fmap2, int B, int C, int H,
                                    int W, int L, float* dfmap1, float* dfmap2, void* workspace, size_t workspace_bytes,
                                    void* stream) {
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
    return corr_bwd_tc_launch(dlvl_ptrs, qs, fmap1, fmap2, B, C, H, W, L, alpha, dfmap1, dfmap2, workspace, workspace_bytes, st);
  }
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
