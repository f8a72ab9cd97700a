// K3 / K3b: radius-r bilinear lookup into the correlation pyramid, all levels in one launch.
//
// Replaces MutliScalePairwise4DCorr.__call__ + bilinear_sampler (reference
// ezflow/similarity/correlation/pairwise.py:43-69, ezflow/utils/resampling.py:56-85) and the autograd
// adjoint of that op w.r.t. the pyramid.
//
// Every query (b,h,w) owns a private correlation map per level, so the lookup is a gather of one
// (2r+2)x(2r+2) patch per (query, level): the (2r+1)^2 window taps sit at integer offsets from the
// centre and therefore share one fractional part.  A block serves 32 consecutive queries:
//   phase 1  warps fetch whole patches from the warp-strip-major pyramid with 16-byte vector loads (a
//            patch row is 20 B in fp16: <= 3 chunks; 30 lanes fetch a full patch in one instruction),
//            dequantise and drop them into shared memory (fp32, odd per-query stride => conflict-free
//            in phase 2);
//   phase 2  lane = query, warps stride over the window taps; every store instruction writes 128
//            contiguous bytes of the (B, L*(2r+1)^2, H, W) output.
// HBM-bound: algorithmic bytes per query = L*[(2r+2)^2*es + (2r+1)^2*4] + 8  (SURVEY.md §8d).
#include "ezb_common.cuh"

namespace ezb {
namespace k3 {

constexpr int QB = 32;        // queries per block
constexpr int NTHREADS = 256;
constexpr int MAX_RADIUS = 8;

struct LookupParams {
  const char* pyr;      // warp-strip-major pyramid (ezb_common.cuh)
  int h[PYR_MAX_LEVELS], w[PYR_MAX_LEVELS];
  int n_pw, n_patches, n_groups;
  long long group_bytes;
  const float* deq;     // [B]
  const float* coords;  // (B,2,N)
  float* out;           // (B, L*K*K, N)
  int B, N, L, r;
};

__device__ __forceinline__ int float_floor_to_int_clamped(float v, int lo, int hi) {
  // NaN -> lo; +-inf and huge values saturate
  float f = floorf(v);
  f = fminf(fmaxf(f, (float)lo), (float)hi);
  return (f == f) ? (int)f : lo;
}

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg_nc_v2(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

template <typename T>
__device__ __forceinline__ float word_elem(const uint32_t (&wd)[4], int j);
template <>
__device__ __forceinline__ float word_elem<__half>(const uint32_t (&wd)[4], int j) {
  const uint32_t u = wd[j >> 1];
  const __half_raw hr = {(unsigned short)((j & 1) ? (u >> 16) : (u & 0xffff))};
  return __half2float(__half(hr));
}
template <>
__device__ __forceinline__ float word_elem<float>(const uint32_t (&wd)[4], int j) {
  return __uint_as_float(wd[j]);
}

template <typename T>
__global__ void __launch_bounds__(NTHREADS) corr_lookup_fwd_kernel(const LookupParams P) {
  extern __shared__ float patches[];  // [QB][pstride]
  constexpr int ES = (int)sizeof(T);
  constexpr int EPC = 16 / ES;  // elements per 16-byte chunk
  const int r = P.r, D = 2 * r + 2, DD = D * D, K = 2 * r + 1, KK = K * K;
  const int pstride = (P.L * DD) | 1;
  const int b = blockIdx.y, grp = blockIdx.x, q0 = grp * QB;  // QB == PYR_QG: a block serves one query group
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float deq = __ldg(P.deq + b);
  const float* cx_ptr = P.coords + (long long)b * 2 * P.N;
  const float* cy_ptr = cx_ptr + P.N;
  const char* pair_base = P.pyr + (long long)b * P.n_patches * P.n_groups * P.group_bytes + (long long)grp * P.group_bytes;
  const long long patch_stride = (long long)P.n_groups * P.group_bytes;

  // ---------------- phase 1: gather patches ----------------
  // item = (level, query).  A patch row is fetched as aligned groups of GW consecutive x (one 16-byte
  // chunk, or the 4-wide level-3 tile row); a group never straddles a tile because tile widths are
  // multiples of GW.  task = (row, group); up to 30 lanes fetch a whole (2r+2)^2 patch per instruction.
  for (int item = warp; item < QB * P.L; item += NTHREADS / 32) {
    const int l = item / QB, ql = item - l * QB, q = q0 + ql;
    if (q >= P.N) continue;
    const int tw_log = 5 - l, th_log = 3 - l;              // tile (8>>l) x (32>>l)
    const int GW = min(EPC, 1 << tw_log);
    const int gw_log = (GW == 8) ? 3 : 2;
    const int ng = (GW - 1 + D + GW - 1) >> gw_log;        // groups covering D elements at any misalignment
    const int tasks = D * ng;
    const float inv = 1.0f / (float)(1 << l);
    const int hl = P.h[l], wl = P.w[l];
    const int x0 = float_floor_to_int_clamped(__ldg(cx_ptr + q) * inv, -(1 << 24), 1 << 24);
    const int y0 = float_floor_to_int_clamped(__ldg(cy_ptr + q) * inv, -(1 << 24), 1 << 24);
    const int xs = max(-D, min(wl, x0 - r));  // first patch column / row (clamped: fully-outside stays fully outside)
    const int ys = max(-D, min(hl, y0 - r));
    const int g0 = (xs + 64) >> gw_log;       // floor(xs / GW) + 64/GW   (xs >= -D >= -18)
    float* dst = patches + ql * pstride + l * DD;
    for (int t = lane; t < tasks; t += 32) {
      const int row = t / ng, gi = t - row * ng;
      const int yy_abs = ys + row;
      const int xg = ((g0 + gi) << gw_log) - 64;  // first x of this group
      const bool ok = (yy_abs >= 0) && (yy_abs < hl) && (xg + GW > 0) && (xg < wl);
      uint32_t wd[4] = {0u, 0u, 0u, 0u};
      if (ok && xg >= 0) {
        const int ph = yy_abs >> th_log, yy = yy_abs & ((1 << th_log) - 1);
        const int pw = xg >> tw_log, xx = xg & ((1 << tw_log) - 1);
        const char* src = pair_base + (long long)(ph * P.n_pw + pw) * patch_stride + pyr_elem_offset(ES, l, ql, yy, xx);
        if (GW * ES == 16) {
          const uint4 v = ldg_nc_v4(src);
          wd[0] = v.x; wd[1] = v.y; wd[2] = v.z; wd[3] = v.w;
        } else {  // fp16 level 3: 4 halfs = 8 bytes
          const uint2 v = ldg_nc_v2(src);
          wd[0] = v.x; wd[1] = v.y;
        }
      }
#pragma unroll
      for (int j = 0; j < EPC; ++j) {
        if (j < GW) {
          const int xa = xg + j, xr = xa - xs;
          if (xr >= 0 && xr < D) dst[row * D + xr] = (ok && xa >= 0 && xa < wl) ? word_elem<T>(wd, j) * deq : 0.f;
        }
      }
    }
  }
  __syncthreads();

  // ---------------- phase 2: bilinear taps, lane = query ----------------
  const int q = q0 + lane;
  if (q >= P.N) return;
  const float x = __ldg(cx_ptr + q), y = __ldg(cy_ptr + q);
  const float* mine = patches + lane * pstride;
  float* out = P.out + (long long)b * P.L * KK * P.N + q;
  for (int l = 0; l < P.L; ++l) {
    const float inv = 1.0f / (float)(1 << l);
    const float cx = x * inv, cy = y * inv;
    const float fx = cx - floorf(cx), fy = cy - floorf(cy);
    // a clamped patch origin means the window is entirely outside: the patch is all zeros, weights moot
    const float w00 = (1.f - fx) * (1.f - fy), w01 = fx * (1.f - fy), w10 = (1.f - fx) * fy, w11 = fx * fy;
    const float* pl = mine + l * DD;
    for (int k = warp; k < KK; k += NTHREADS / 32) {
      const int i = k / K, j = k - i * K;  // i offsets x (slow index!), j offsets y   pairwise.py:53-61
      const float* p = pl + j * D + i;
      const float v = w00 * p[0] + w01 * p[1] + w10 * p[D] + w11 * p[D + 1];
      out[(long long)(l * KK + k) * P.N] = v;
    }
  }
}

// ------------------------------------------------------------------------------------ K3b
struct LookupBwdParams {
  float* dlvl[EZB_MAX_LEVELS];
  long long qstride[EZB_MAX_LEVELS];
  int h[EZB_MAX_LEVELS], w[EZB_MAX_LEVELS], pitch[EZB_MAX_LEVELS];
  const float* dout;    // (B, L*K*K, N)
  const float* coords;  // (B,2,N)
  int B, N, L, r;
};

// dpyr[l][b,q,y,x] += sum over window taps of weight * dout.  A query's gradient patch is private to
// that query, so the block first builds the (2r+2)^2 patch in shared memory (each element gathers <= 4
// taps) and then adds it to HBM with coalesced row-segment reductions (no inter-query conflicts).
__global__ void __launch_bounds__(NTHREADS) corr_lookup_bwd_kernel(const LookupBwdParams P) {
  extern __shared__ float sm[];
  const int r = P.r, D = 2 * r + 2, DD = D * D, K = 2 * r + 1, KK = K * K;
  const int dstride = (P.L * KK) | 1;  // dout values per query (odd stride)
  float* s_d = sm;                     // [QB][dstride]
  float* s_g = sm + QB * dstride;      // [QB][L*DD]
  const int b = blockIdx.y, q0 = blockIdx.x * QB;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* cx_ptr = P.coords + (long long)b * 2 * P.N;
  const float* cy_ptr = cx_ptr + P.N;

  // load dout: lane = query (coalesced), warps stride over channels
  {
    const int q = q0 + lane;
    const float* src = P.dout + (long long)b * P.L * KK * P.N + q;
    for (int c = warp; c < P.L * KK; c += NTHREADS / 32) s_d[lane * dstride + c] = (q < P.N) ? __ldg(src + (long long)c * P.N) : 0.f;
  }
  __syncthreads();
  // gradient patches: thread -> (query, level, Y, X)
  for (int idx = threadIdx.x; idx < QB * P.L * DD; idx += NTHREADS) {
    const int ql = idx / (P.L * DD);
    int rem = idx - ql * (P.L * DD);
    const int l = rem / DD;
    rem -= l * DD;
    const int Y = rem / D, X = rem - Y * D;
    const int q = min(q0 + ql, P.N - 1);
    const float inv = 1.0f / (float)(1 << l);
    const float cx = __ldg(cx_ptr + q) * inv, cy = __ldg(cy_ptr + q) * inv;
    const float fx = cx - floorf(cx), fy = cy - floorf(cy);
    const float* d = s_d + ql * dstride + l * KK;
    // out(i,j) touches patch (row j+a, col i+c) with weight wy[a]*wx[c]; invert: i = X-c, j = Y-a
    float g = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int i = X - c, j = Y - a;
        if (i >= 0 && i < K && j >= 0 && j < K) g += (a ? fy : 1.f - fy) * (c ? fx : 1.f - fx) * d[i * K + j];
      }
    s_g[ql * (P.L * DD) + l * DD + Y * D + X] = g;
  }
  __syncthreads();
  // add to HBM: warp per (query, level), lanes sweep the patch row-major (row segments are contiguous)
  for (int item = warp; item < QB * P.L; item += NTHREADS / 32) {
    const int l = item / QB, ql = item - l * QB, q = q0 + ql;
    if (q >= P.N) continue;
    const float inv = 1.0f / (float)(1 << l);
    const int hl = P.h[l], wl = P.w[l];
    const int x0 = float_floor_to_int_clamped(__ldg(cx_ptr + q) * inv, -(1 << 24), 1 << 24);
    const int y0 = float_floor_to_int_clamped(__ldg(cy_ptr + q) * inv, -(1 << 24), 1 << 24);
    const int xs = max(-D, min(wl, x0 - r)), ys = max(-D, min(hl, y0 - r));
    float* dst = P.dlvl[l] + ((long long)b * P.N + q) * P.qstride[l];
    const float* g = s_g + ql * (P.L * DD) + l * DD;
    for (int e = lane; e < DD; e += 32) {
      const int Y = e / D, X = e - Y * D;
      const int yy = ys + Y, xx = xs + X;
      if (yy >= 0 && yy < hl && xx >= 0 && xx < wl) atomicAdd(dst + (long long)yy * P.pitch[l] + xx, g[e]);
    }
  }
}

}  // namespace k3
}  // namespace ezb

using namespace ezb;

extern "C" int ezb_corr_lookup_fwd(const void* pyramid, const float* deq_scale, const float* coords, int B, int H, int W, int L,
                                   int radius, int dtype, float* out, void* stream) {
  EZB_REQUIRE(pyramid && deq_scale && coords && out, "null pointer");
  EZB_REQUIRE(B > 0 && H > 0 && W > 0 && L >= 1, "bad shape B=%d H=%d W=%d L=%d", B, H, W, L);
  EZB_REQUIRE(radius >= 0 && radius <= k3::MAX_RADIUS, "corr_radius %d outside [0,%d]", radius, k3::MAX_RADIUS);
  EZB_REQUIRE(dtype == EZB_F32 || dtype == EZB_F16, "bad dtype %d", dtype);
  if (L > PYR_MAX_LEVELS) return fail(EZB_ERR_UNSUPPORTED, "num_levels %d > %d is not supported by the fused pyramid", L, PYR_MAX_LEVELS);
  PyrLayout lay;
  EZB_REQUIRE(pyr_layout(H, W, L, dtype, &lay), "feature map %dx%d too small for %d pyramid levels", H, W, L);
  EZB_REQUIRE((reinterpret_cast<uintptr_t>(pyramid) & 127) == 0, "pyramid buffer must be 128-byte aligned");
  k3::LookupParams P;
  const int N = H * W;
  for (int l = 0; l < L; ++l) {
    P.h[l] = H >> l;
    P.w[l] = W >> l;
  }
  P.pyr = static_cast<const char*>(pyramid);
  P.n_pw = lay.n_pw;
  P.n_patches = lay.n_patches;
  P.n_groups = lay.n_groups;
  P.group_bytes = lay.group_bytes;
  P.deq = deq_scale;
  P.coords = coords;
  P.out = out;
  P.B = B;
  P.N = N;
  P.L = L;
  P.r = radius;
  const int D = 2 * radius + 2;
  const size_t smem = (size_t)k3::QB * ((L * D * D) | 1) * sizeof(float);
  if (smem > 200 * 1024) return fail(EZB_ERR_UNSUPPORTED, "lookup window too large (L=%d, r=%d)", L, radius);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static_assert(k3::QB == PYR_QG, "one lookup block per pyramid query group");
  dim3 grid(lay.n_groups, B);
  if (dtype == EZB_F16) {
    EZB_CUDA_OK(cudaFuncSetAttribute(k3::corr_lookup_fwd_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k3::corr_lookup_fwd_kernel<__half><<<grid, k3::NTHREADS, smem, st>>>(P);
  } else {
    EZB_CUDA_OK(cudaFuncSetAttribute(k3::corr_lookup_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k3::corr_lookup_fwd_kernel<float><<<grid, k3::NTHREADS, smem, st>>>(P);
  }
  EZB_LAUNCH_OK();
  return EZB_OK;
}

extern "C" int ezb_corr_grad_layout(int H, int W, int L, int* h, int* w, int* pitch, int64_t* qstride) {
  EZB_REQUIRE(H > 0 && W > 0 && L >= 1 && L <= EZB_MAX_LEVELS, "bad pyramid shape H=%d W=%d L=%d", H, W, L);
  LevelGeom g[EZB_MAX_LEVELS];
  EZB_REQUIRE(pyramid_geom(H, W, L, EZB_F32, g), "feature map %dx%d too small for %d pyramid levels", H, W, L);
  for (int l = 0; l < L; ++l) {
    if (h) h[l] = g[l].h;
    if (w) w[l] = g[l].w;
    if (pitch) pitch[l] = g[l].pitch;
    if (qstride) qstride[l] = g[l].qstride;
  }
  return EZB_OK;
}

extern "C" int ezb_corr_lookup_bwd(const float* dout, const float* coords, int B, int H, int W, int L, int radius,
                                   float* const* dlvl_ptrs, void* stream) {
  EZB_REQUIRE(dout && coords && dlvl_ptrs, "null pointer");
  EZB_REQUIRE(B > 0 && H > 0 && W > 0 && L >= 1 && L <= EZB_MAX_LEVELS, "bad shape B=%d H=%d W=%d L=%d", B, H, W, L);
  EZB_REQUIRE(radius >= 0 && radius <= k3::MAX_RADIUS, "corr_radius %d outside [0,%d]", radius, k3::MAX_RADIUS);
  LevelGeom g[EZB_MAX_LEVELS];
  EZB_REQUIRE(pyramid_geom(H, W, L, EZB_F32, g), "feature map %dx%d too small for %d pyramid levels", H, W, L);
  k3::LookupBwdParams P;
  const int N = H * W;
  for (int l = 0; l < L; ++l) {
    EZB_REQUIRE(dlvl_ptrs[l], "level %d gradient pointer is null", l);
    P.dlvl[l] = dlvl_ptrs[l];
    P.h[l] = g[l].h;
    P.w[l] = g[l].w;
    P.pitch[l] = g[l].pitch;
    P.qstride[l] = g[l].qstride;
  }
  P.dout = dout;
  P.coords = coords;
  P.B = B;
  P.N = N;
  P.L = L;
  P.r = radius;
  const int D = 2 * radius + 2, K = 2 * radius + 1;
  const size_t smem = (size_t)k3::QB * (((L * K * K) | 1) + L * D * D) * sizeof(float);
  if (smem > 200 * 1024) return fail(EZB_ERR_UNSUPPORTED, "lookup window too large (L=%d, r=%d)", L, radius);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EZB_CUDA_OK(cudaFuncSetAttribute(k3::corr_lookup_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k3::corr_lookup_bwd_kernel<<<dim3(cdiv(N, k3::QB), B), k3::NTHREADS, smem, st>>>(P);
  EZB_LAUNCH_OK();
  return EZB_OK;
}
