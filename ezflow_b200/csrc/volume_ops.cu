// the cost-volume assemblies of the two remaining models, and the memory-free RAFT lookup.
//
//  * VCN   `_corr_fn`  (reference ezflow/models/vcn.py:162-206): per-CHANNEL shifted products,
//        out[b,c,i,j,y,x] = leaky_relu( f1[b,c,y,x] * f2[b,c,y + j - mv, x + i - mu], 0.1 ),  0 where the shifted pixel is outside
//    (i shifts x over 2*max_disp+1 values, j shifts y over 2*(max_disp//factorization)+1 values).  The reference fills the
//    volume with (2mu+1)(2mv+1) slice-multiply-assign ops and a separate activation pass.
//  * DICL  `LearnableMatchingCost.forward` volume assembly (ezflow/similarity/learnable_cost.py:157-231): for every
//    displacement (i, j) the 2C-channel image [x ; y shifted by (j - mv, i - mu)] restricted to the overlap, optionally
//    zeroed where the shifted y is all-zero ("warp holes"), delivered as (B*U*V, 2C, H, W) for the matching network.  The
//    reference zero-fills a (B,2C,U,V,H,W) tensor, assigns 2*U*V slices, multiplies by the mask and permute-copies.
//  * On-demand correlation lookup ("alt_cuda_corr"): MutliScalePairwise4DCorr's output (pairwise.py:43-69) computed from the
//    feature maps directly — the (2r+2)^2 dot products each query needs per level against the l-times pooled fmap2 —
//    without the O(N^2) volume, for inputs whose pyramid does not fit in memory.  Exact fp32.
//
// All three are HBM/L2-bound data-movement kernels on CUDA cores: one pass, every output element written once.
#include "ezb_common.cuh"

#include <cstdlib>

#include <climits>

namespace ezb {
namespace k8 {

// --------------------------------------------------------------------------------------------- VCN
struct VcnParams {
  const float* f1;
  const float* f2;
  const float* dout;  // backward only
  float* out;         // forward: (BC, U, V, H, W); backward: df1 / df2 (BC, H, W)
  float* out2;
  long long BC;
  int H, W, U, V, mu, mv;
  float slope;
};

__global__ void __launch_bounds__(256) vcn_corr_fwd_kernel(const VcnParams P) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= P.W || y >= P.H) return;
  const long long HW = (long long)P.H * P.W;
  for (long long bc = blockIdx.z; bc < P.BC; bc += gridDim.z) {
    const float a = __ldg(P.f1 + bc * HW + (long long)y * P.W + x);
    const float* f2 = P.f2 + bc * HW;
    float* o = P.out + bc * P.U * P.V * HW + (long long)y * P.W + x;
    for (int i = 0; i < P.U; ++i) {
      const int x2 = x + i - P.mu;
      for (int j = 0; j < P.V; ++j) {
        const int y2 = y + j - P.mv;
        float v = 0.f;
        if (x2 >= 0 && x2 < P.W && y2 >= 0 && y2 < P.H) {
          v = a * __ldg(f2 + (long long)y2 * P.W + x2);
          v = v > 0.f ? v : v * P.slope;
        }
        o[(long long)(i * P.V + j) * HW] = v;
      }
    }
  }
}

// df1[bc,y,x] = sum_ij g'(i,j,y,x) f2[y+dj, x+di];  df2[bc,y',x'] = sum_ij g'(i,j,y'-dj,x'-di) f1[y'-dj, x'-di]
// g' = dout * (prod > 0 ? 1 : slope)   (F.leaky_relu's backward)
__global__ void __launch_bounds__(256) vcn_corr_bwd_kernel(const VcnParams P) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= P.W || y >= P.H) return;
  const long long HW = (long long)P.H * P.W;
  for (long long bc = blockIdx.z; bc < P.BC; bc += gridDim.z) {
    const float* f1 = P.f1 + bc * HW;
    const float* f2 = P.f2 + bc * HW;
    const float* g = P.dout + bc * P.U * P.V * HW;
    const float a = __ldg(f1 + (long long)y * P.W + x), b = __ldg(f2 + (long long)y * P.W + x);
    float d1 = 0.f, d2 = 0.f;
    for (int i = 0; i < P.U; ++i) {
      const int di = i - P.mu;
      for (int j = 0; j < P.V; ++j) {
        const int dj = j - P.mv;
        const long long plane = (long long)(i * P.V + j) * HW;
        if (x + di >= 0 && x + di < P.W && y + dj >= 0 && y + dj < P.H) {  // this pixel as the f1 pixel
          const float t = __ldg(f2 + (long long)(y + dj) * P.W + x + di), pr = a * t;
          d1 += __ldg(g + plane + (long long)y * P.W + x) * (pr > 0.f ? 1.f : P.slope) * t;
        }
        if (x - di >= 0 && x - di < P.W && y - dj >= 0 && y - dj < P.H) {  // this pixel as the shifted f2 pixel
          const float s = __ldg(f1 + (long long)(y - dj) * P.W + x - di), pr = s * b;
          d2 += __ldg(g + plane + (long long)(y - dj) * P.W + x - di) * (pr > 0.f ? 1.f : P.slope) * s;
        }
      }
    }
    if (P.out) P.out[bc * HW + (long long)y * P.W + x] = d1;
    if (P.out2) P.out2[bc * HW + (long long)y * P.W + x] = d2;
  }
}

// -------------------------------------------------------------------------------------------- DICL
struct DiclParams {
  const float* x;
  const float* y;
  const float* dout;    // backward: (B*U*V, 2C, H, W)
  float* out;           // forward: (B*U*V, 2C, H, W); backward: dx (B,C,H,W)
  float* out2;          // backward: dy
  unsigned char* mask;  // (B*U*V, H, W) 1 = kept; written by the forward (optional), read by the backward
  int B, C, H, W, U, V, mu, mv, remove_holes;
};

__global__ void __launch_bounds__(256) dicl_volume_fwd_kernel(const DiclParams P) {
  const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (px >= P.W || py >= P.H) return;
  const long long HW = (long long)P.H * P.W, pix = (long long)py * P.W + px;
  const int n_img = P.B * P.U * P.V;
  for (int img = blockIdx.z; img < n_img; img += gridDim.z) {
    const int b = img / (P.U * P.V), ij = img - b * (P.U * P.V), i = ij / P.V, j = ij - i * P.V;
    const int x2 = px + i - P.mu, y2 = py + j - P.mv;
    const bool inside = x2 >= 0 && x2 < P.W && y2 >= 0 && y2 < P.H;
    const float* xs = P.x + (long long)b * P.C * HW + pix;
    const float* ys = P.y + (long long)b * P.C * HW + (long long)y2 * P.W + x2;
    bool keep = inside;
    if (inside && P.remove_holes) {  // valid_mask = (sum over the shifted-y channels) != 0   learnable_cost.py:218-222
      float s = 0.f;
      for (int c = 0; c < P.C; ++c) s += __ldg(ys + (long long)c * HW);
      keep = s != 0.f;
    }
    float* o = P.out + (long long)img * 2 * P.C * HW + pix;
    for (int c = 0; c < P.C; ++c) o[(long long)c * HW] = keep ? __ldg(xs + (long long)c * HW) : 0.f;
    for (int c = 0; c < P.C; ++c) o[(long long)(P.C + c) * HW] = keep ? __ldg(ys + (long long)c * HW) : 0.f;
    if (P.mask) P.mask[(long long)img * HW + pix] = keep ? 1 : 0;
  }
}

// dx[b,c,p] = sum_ij mask(b,ij,p) g[(b,ij), c, p];   dy[b,c,p'] = sum_ij mask(b,ij,p'-d) g[(b,ij), C+c, p'-d]
// grid.z strides over (b, c)
__global__ void __launch_bounds__(256) dicl_volume_bwd_kernel(const DiclParams P) {
  const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (px >= P.W || py >= P.H) return;
  const long long HW = (long long)P.H * P.W, pix = (long long)py * P.W + px;
  for (int bc = blockIdx.z; bc < P.B * P.C; bc += gridDim.z) {
    const int b = bc / P.C, c = bc - b * P.C;
    float d1 = 0.f, d2 = 0.f;
    for (int i = 0; i < P.U; ++i)
      for (int j = 0; j < P.V; ++j) {
        const long long img = ((long long)b * P.U + i) * P.V + j;
        const float* g = P.dout + img * 2 * P.C * HW;
        if (P.mask[img * HW + pix]) d1 += __ldg(g + (long long)c * HW + pix);
        const int xq = px - (i - P.mu), yq = py - (j - P.mv);  // the unshifted pixel this y pixel was paired with
        if (xq >= 0 && xq < P.W && yq >= 0 && yq < P.H) {
          const long long q = (long long)yq * P.W + xq;
          if (P.mask[img * HW + q]) d2 += __ldg(g + (long long)(P.C + c) * HW + q);
        }
      }
    if (P.out) P.out[(long long)bc * HW + pix] = d1;
    if (P.out2) P.out2[(long long)bc * HW + pix] = d2;
  }
}

// --------------------------------------------------------------------------- on-demand correlation
// NHWC copies: dst[b, n, c] = src[b, c, n]; with pool > 1 the destination pixel is the floor 2x2-average-pooled one of
// the PREVIOUS level (src is then NHWC of the previous level)
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int N) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = warp; i < 32; i += 8) {
    const int c = c0 + i, n = n0 + lane;
    tile[i][lane] = (c < C && n < N) ? __ldg(src + ((long long)b * C + c) * N + n) : 0.f;
  }
  __syncthreads();
  for (int i = warp; i < 32; i += 8) {
    const int n = n0 + i, c = c0 + lane;
    if (n < N && c < C) dst[((long long)b * N + n) * C + c] = tile[lane][i];
  }
}
__global__ void __launch_bounds__(256) pool_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int hs, int ws,
                                                        int hd, int wd, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int x = (int)(r % wd);
    r /= wd;
    const int y = (int)(r % hd);
    const long long b = r / hd;
    const float* s = src + ((b * hs + 2 * y) * ws + 2 * x) * C + c;
    dst[i] = 0.25f * ((__ldg(s) + __ldg(s + C)) + (__ldg(s + (long long)ws * C) + __ldg(s + (long long)ws * C + C)));
  }
}

struct OnDemandParams {
  const float* f1;              // (B, N, C) NHWC
  const float* f2[PYR_MAX_LEVELS];  // level l: (B, h_l, w_l, C) NHWC
  const float* coords;          // (B, 2, N)
  float* out;                   // (B, L*K*K, N)
  int B, C, H, W, N, L, r;
  float scale;                  // 1 / sqrt(C)
};

// One warp per (query, level): the lanes split the (2r+2)^2 window pixels, each forming the full C-long dot product with
// the query's feature vector (shared memory, broadcast reads; the fmap2 pixel is a contiguous C-float row: float4 loads),
// then the (2r+1)^2 bilinear taps are formed from the window of dot products.  Block = 8 warps = 2 queries x 4 levels.
constexpr int OD_WARPS = 8;
__global__ void __launch_bounds__(OD_WARPS * 32) corr_ondemand_lookup_kernel(const OnDemandParams P) {
  extern __shared__ float sm[];
  const int r = P.r, D = 2 * r + 2, K = 2 * r + 1, KK = K * K;
  const int qpb = OD_WARPS / P.L > 0 ? OD_WARPS / P.L : 1;  // queries per block
  float* s_f1 = sm;                                // [qpb][C]
  float* s_win = sm + qpb * P.C;                   // [OD_WARPS][D*D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const long long q0 = (long long)blockIdx.x * qpb;
  for (int i = threadIdx.x; i < qpb * P.C; i += blockDim.x) {
    const long long q = q0 + i / P.C;
    s_f1[i] = q < P.N ? __ldg(P.f1 + ((long long)b * P.N + q) * P.C + (i % P.C)) : 0.f;
  }
  __syncthreads();
  const int ql = warp / P.L, l = warp - ql * P.L;
  const long long q = q0 + ql;
  if (ql >= qpb || q >= P.N) return;
  const int hl = P.H >> l, wl = P.W >> l;
  const float inv = 1.0f / (float)(1 << l);
  const float cx = __ldg(P.coords + (long long)b * 2 * P.N + q) * inv, cy = __ldg(P.coords + (long long)b * 2 * P.N + P.N + q) * inv;
  float fxf = floorf(cx), fyf = floorf(cy);
  fxf = fminf(fmaxf(fxf, -16777216.f), 16777216.f);
  fyf = fminf(fmaxf(fyf, -16777216.f), 16777216.f);
  const int x0 = (fxf == fxf) ? (int)fxf : -(1 << 24), y0 = (fyf == fyf) ? (int)fyf : -(1 << 24);
  const float fx = cx - floorf(cx), fy = cy - floorf(cy);
  const int xs = max(-D, min(wl, x0 - r)), ys = max(-D, min(hl, y0 - r));  // a window entirely outside stays outside
  const float* f1v = s_f1 + ql * P.C;
  const float* lvl = P.f2[l] + (long long)b * hl * wl * P.C;
  float* win = s_win + warp * D * D;
  const bool vec = (P.C & 3) == 0;
  for (int t = lane; t < D * D; t += 32) {
    const int wy = t / D, wx = t - wy * D, y = ys + wy, x = xs + wx;
    float acc = 0.f;
    if (y >= 0 && y < hl && x >= 0 && x < wl) {
      const float* p = lvl + ((long long)y * wl + x) * P.C;
      if (vec) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int c = 0; c < P.C; c += 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(p + c));
          a0 = fmaf(v.x, f1v[c], a0);
          a1 = fmaf(v.y, f1v[c + 1], a1);
          a2 = fmaf(v.z, f1v[c + 2], a2);
          a3 = fmaf(v.w, f1v[c + 3], a3);
        }
        acc = (a0 + a1) + (a2 + a3);
      } else {
        for (int c = 0; c < P.C; ++c) acc = fmaf(__ldg(p + c), f1v[c], acc);
      }
    }
    win[t] = acc * P.scale;
  }
  __syncwarp();
  float* o = P.out + ((long long)b * P.L * KK + (long long)l * KK) * P.N + q;
  for (int t = lane; t < KK; t += 32) {
    const int i = t / K, j = t - i * K;  // i offsets x (slow output index), j offsets y   pairwise.py:53-61
    const float* w = win + j * D + i;
    const float top = (1.f - fx) * w[0] + fx * w[1], bot = (1.f - fx) * w[D] + fx * w[D + 1];
    o[(long long)t * P.N] = top + fy * (bot - top);
  }
}


// The same lookup with the fmap2 pixels shared between neighbouring queries (round 2).  The kernel above fetches every
// window pixel's C-float row from L2 for every query: 100 KB per (query, level), 13 GB of L2 traffic per 1080p pair and
// lookup (3 ms).  Here a block owns 8 x 8 queries of one level; when the bounding box of their windows fits a 24 x 28
// region (coherent flow: always at the coarse levels) the region is staged in shared memory 32 channels at a time and shared
// by the 64 queries: thread = (query, every fifth window row), <= 20 running dot products in registers, per chunk the
// query's 32 f1 values in registers and eight LDS.128 per window pixel (pixel pitch 36 floats and 24 pixels per region row:
// the eight x-adjacent queries of a quarter-warp hit eight different bank groups whatever rows their windows start in).  A
// bounding box of up to four regions is swept region by region; blocks whose windows scatter wider fall back to per-pixel
// global loads inside the same kernel.
constexpr int ODT_QW = 8, ODT_QH = 8, ODT_Q = ODT_QW * ODT_QH, ODT_RH = 28, ODT_RW = 24, ODT_CK = 32, ODT_PITCH = ODT_CK + 4;
constexpr int ODT_PARTS = 5, ODT_THREADS = ODT_Q * ODT_PARTS, ODT_MAXD = 10, ODT_NR = (ODT_MAXD + ODT_PARTS - 1) / ODT_PARTS;
constexpr size_t ODT_SMEM = ((size_t)ODT_RH * ODT_RW + ODT_Q) * ODT_PITCH * sizeof(float);
__global__ void __launch_bounds__(ODT_THREADS, 2) corr_ondemand_tiled_kernel(const OnDemandParams P) {
  extern __shared__ __align__(16) float sm[];
  float* s_reg = sm;                                 // [ODT_RH * ODT_RW][ODT_PITCH]; afterwards the windows of dot products
  float* s_f1 = sm + ODT_RH * ODT_RW * ODT_PITCH;    // [ODT_Q][ODT_PITCH]
  __shared__ int s_box[4];
  const int r = P.r, D = 2 * r + 2, K = 2 * r + 1, KK = K * K, DD = D * D;
  const int tid = threadIdx.x, ql = tid % ODT_Q, part = tid / ODT_Q;
  const int b = blockIdx.z / P.L, l = blockIdx.z - b * P.L;
  const int qx = blockIdx.x * ODT_QW + (ql % ODT_QW), qy = blockIdx.y * ODT_QH + ql / ODT_QW;
  const bool qok = qx < P.W && qy < P.H;
  const long long q = (long long)qy * P.W + qx;
  const int hl = P.H >> l, wl = P.W >> l;
  int xs = -D, ys = -D;
  float fx = 0.f, fy = 0.f;
  if (qok) {
    const float inv = 1.0f / (float)(1 << l);
    const float cx = __ldg(P.coords + (long long)b * 2 * P.N + q) * inv, cy = __ldg(P.coords + (long long)b * 2 * P.N + P.N + q) * inv;
    float fxf = floorf(cx), fyf = floorf(cy);
    fxf = fminf(fmaxf(fxf, -16777216.f), 16777216.f);
    fyf = fminf(fmaxf(fyf, -16777216.f), 16777216.f);
    const int x0 = (fxf == fxf) ? (int)fxf : -(1 << 24), y0 = (fyf == fyf) ? (int)fyf : -(1 << 24);
    fx = cx - floorf(cx);
    fy = cy - floorf(cy);
    xs = max(-D, min(wl, x0 - r));
    ys = max(-D, min(hl, y0 - r));  // a window entirely outside stays outside
  }
  const bool any = qok && xs + D > 0 && xs < wl && ys + D > 0 && ys < hl;
  if (tid == 0) { s_box[0] = INT_MAX; s_box[1] = INT_MAX; s_box[2] = INT_MIN; s_box[3] = INT_MIN; }
  __syncthreads();
  if (part == 0 && any) {
    atomicMin(&s_box[0], xs); atomicMin(&s_box[1], ys);
    atomicMax(&s_box[2], xs); atomicMax(&s_box[3], ys);
  }
  __syncthreads();
  const bool none = s_box[0] == INT_MAX;  // no window of this block touches the level: all zeros
  const int rx0 = none ? 0 : max(s_box[0], 0), ry0 = none ? 0 : max(s_box[1], 0);
  const int rw = none ? 0 : min(s_box[2] + D, wl) - rx0, rh = none ? 0 : min(s_box[3] + D, hl) - ry0;
  // the bounding box is swept in sub-rectangles of at most ODT_RH x ODT_RW pixels (one when the flow is coherent); more than
  // four: the windows are scattered and sharing does not pay -> per-pixel loads below
  const int nsy = (rh + ODT_RH - 1) / ODT_RH, nsx = (rw + ODT_RW - 1) / ODT_RW;
  const bool tiled = nsy * nsx <= 4;
  const float* lvl = P.f2[l] + (long long)b * hl * wl * P.C;
  const float* f1row = P.f1 + ((long long)b * P.N + (qok ? q : 0)) * P.C;

  // this thread's window rows wy = part, part + ODT_PARTS (a 10 x 10 window: two rows each) x all D columns
  float acc[ODT_NR][ODT_MAXD];
#pragma unroll
  for (int k = 0; k < ODT_NR; ++k)
#pragma unroll
    for (int wx = 0; wx < ODT_MAXD; ++wx) acc[k][wx] = 0.f;

  if (tiled && !none) {
    for (int sub = 0; sub < nsy * nsx; ++sub) {
      const int sy0 = ry0 + (sub / nsx) * ODT_RH, sx0 = rx0 + (sub % nsx) * ODT_RW;    // this pass's sub-rectangle of the level
      const int sh = min(ODT_RH, ry0 + rh - sy0), sw = min(ODT_RW, rx0 + rw - sx0);
      bool rowok[ODT_NR];
      unsigned colmask = 0;
#pragma unroll
      for (int wx = 0; wx < ODT_MAXD; ++wx)
        if (wx < D && xs + wx >= sx0 && xs + wx < sx0 + sw) colmask |= 1u << wx;
#pragma unroll
      for (int k = 0; k < ODT_NR; ++k) {
        const int wy = part + k * ODT_PARTS, y = ys + wy;
        rowok[k] = any && wy < D && y >= sy0 && y < sy0 + sh;
      }
      for (int c0 = 0; c0 < P.C; c0 += ODT_CK) {
        __syncthreads();  // the previous chunk's readers are done
        for (int i = tid; i < sh * sw * (ODT_CK / 4); i += ODT_THREADS) {
          const int px = i / (ODT_CK / 4), k = i - px * (ODT_CK / 4), ry = px / sw, rx = px - ry * sw;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c0 + 4 * k < P.C) v = __ldg(reinterpret_cast<const float4*>(lvl + ((long long)(sy0 + ry) * wl + sx0 + rx) * P.C + c0 + 4 * k));
          *reinterpret_cast<float4*>(s_reg + (ry * ODT_RW + rx) * ODT_PITCH + 4 * k) = v;
        }
        for (int i = tid; i < ODT_Q * (ODT_CK / 4); i += ODT_THREADS) {
          const int qq = i / (ODT_CK / 4), k = i - qq * (ODT_CK / 4);
          const int x = blockIdx.x * ODT_QW + (qq % ODT_QW), y = blockIdx.y * ODT_QH + qq / ODT_QW;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (x < P.W && y < P.H && c0 + 4 * k < P.C)
            v = __ldg(reinterpret_cast<const float4*>(P.f1 + ((long long)b * P.N + (long long)y * P.W + x) * P.C + c0 + 4 * k));
          *reinterpret_cast<float4*>(s_f1 + qq * ODT_PITCH + 4 * k) = v;
        }
        __syncthreads();
        float4 a[ODT_CK / 4];
#pragma unroll
        for (int k = 0; k < ODT_CK / 4; ++k) a[k] = *reinterpret_cast<const float4*>(s_f1 + ql * ODT_PITCH + 4 * k);
#pragma unroll
        for (int k = 0; k < ODT_NR; ++k) {
          if (!rowok[k]) continue;
          const float* rowp = s_reg + ((ys + part + k * ODT_PARTS - sy0) * ODT_RW + (xs - sx0)) * ODT_PITCH;
#pragma unroll
          for (int wx = 0; wx < ODT_MAXD; ++wx) {
            if (!((colmask >> wx) & 1u)) continue;
            const float4* pv = reinterpret_cast<const float4*>(rowp + wx * ODT_PITCH);
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int j = 0; j < ODT_CK / 4; j += 2) {
              const float4 v = pv[j], w = pv[j + 1];
              s0 = fmaf(v.x, a[j].x, s0); s0 = fmaf(v.y, a[j].y, s0); s0 = fmaf(v.z, a[j].z, s0); s0 = fmaf(v.w, a[j].w, s0);
              s1 = fmaf(w.x, a[j + 1].x, s1); s1 = fmaf(w.y, a[j + 1].y, s1); s1 = fmaf(w.z, a[j + 1].z, s1); s1 = fmaf(w.w, a[j + 1].w, s1);
            }
            acc[k][wx] += s0 + s1;
          }
        }
      }
    }
  } else if (!none) {
    unsigned colmask = 0;
#pragma unroll
    for (int wx = 0; wx < ODT_MAXD; ++wx)
      if (wx < D && xs + wx >= 0 && xs + wx < wl) colmask |= 1u << wx;
#pragma unroll
    for (int k = 0; k < ODT_NR; ++k) {
      const int wy0 = part + k * ODT_PARTS;
      if (!(any && wy0 < D && ys + wy0 >= 0 && ys + wy0 < hl)) continue;
      const float* rowp = lvl + ((long long)(ys + part + k * ODT_PARTS) * wl + xs) * P.C;
#pragma unroll
      for (int wx = 0; wx < ODT_MAXD; ++wx) {
        if (!((colmask >> wx) & 1u)) continue;
        const float* p = rowp + (long long)wx * P.C;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int c = 0; c < P.C; c += 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(p + c)), u = __ldg(reinterpret_cast<const float4*>(f1row + c));
          a0 = fmaf(v.x, u.x, a0); a1 = fmaf(v.y, u.y, a1); a2 = fmaf(v.z, u.z, a2); a3 = fmaf(v.w, u.w, a3);
        }
        acc[k][wx] = (a0 + a1) + (a2 + a3);
      }
    }
  }
  __syncthreads();  // the region is dead: its memory now holds every query's window of dot products
  const int WP = DD | 1;
  float* win = s_reg + ql * WP;
#pragma unroll
  for (int k = 0; k < ODT_NR; ++k) {
    const int wy = part + k * ODT_PARTS;
    if (wy >= D) continue;
#pragma unroll
    for (int wx = 0; wx < ODT_MAXD; ++wx)
      if (wx < D) win[wy * D + wx] = acc[k][wx] * P.scale;
  }
  __syncthreads();
  if (!qok) return;
  float* o = P.out + ((long long)b * P.L * KK + (long long)l * KK) * P.N + q;
  for (int t = part; t < KK; t += ODT_PARTS) {
    const int i = t / K, j = t - i * K;  // i offsets x (slow output index), j offsets y   pairwise.py:53-61
    const float* w = win + j * D + i;
    const float top = (1.f - fx) * w[0] + fx * w[1], bot = (1.f - fx) * w[D] + fx * w[D + 1];
    o[(long long)t * P.N] = top + fy * (bot - top);
  }
}

}  // namespace k8
}  // namespace ezb

using namespace ezb;

static dim3 pixel_grid(int H, int W, long long nz) { return dim3(cdiv(W, 32), cdiv(H, 8), (unsigned)std::min<long long>(nz, 16384)); }

extern "C" int ezb_vcn_corr_fwd(const float* f1, const float* f2, int B, int C, int H, int W, int max_disp_x, int max_disp_y,
                                float slope, float* out, void* stream) {
  EZB_REQUIRE(f1 && f2 && out, "null pointer");
  EZB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && max_disp_x >= 0 && max_disp_y >= 0, "bad arguments");
  k8::VcnParams P;
  P.f1 = f1; P.f2 = f2; P.dout = nullptr; P.out = out; P.out2 = nullptr;
  P.BC = (long long)B * C; P.H = H; P.W = W; P.U = 2 * max_disp_x + 1; P.V = 2 * max_disp_y + 1; P.mu = max_disp_x; P.mv = max_disp_y;
  P.slope = slope;
  k8::vcn_corr_fwd_kernel<<<pixel_grid(H, W, P.BC), 256, 0, static_cast<cudaStream_t>(stream)>>>(P);
  EZB_LAUNCH_OK();
  return EZB_OK;
}

extern "C" int ezb_vcn_corr_bwd(const float* f1, const float* f2, const float* dout, int B, int C, int H, int W, int max_disp_x,
                                int max_disp_y, float slope, float* df1, float* df2, void* stream) {
  EZB_REQUIRE(f1 && f2 && dout && (df1 || df2), "null pointer");
  EZB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && max_disp_x >= 0 && max_disp_y >= 0, "bad arguments");
  k8::VcnParams P;
  P.f1 = f1; P.f2 = f2; P.dout = dout; P.out = df1; P.out2 = df2;
  P.BC = (long long)B * C; P.H = H; P.W = W; P.U = 2 * max_disp_x + 1; P.V = 2 * max_disp_y + 1; P.mu = max_disp_x; P.mv = max_disp_y;
  P.slope = slope;
  k8::vcn_corr_bwd_kernel<<<pixel_grid(H, W, P.BC), 256, 0, static_cast<cudaStream_t>(stream)>>>(P);
  EZB_LAUNCH_OK();
  return EZB_OK;
}

extern "C" int ezb_dicl_volume_fwd(const float* x, const float* y, int B, int C, int H, int W, int max_u, int max_v,
                                   int remove_warp_hole, float* out, unsigned char* mask, void* stream) {
  EZB_REQUIRE(x && y && out, "null pointer");
  EZB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && max_u >= 0 && max_v >= 0, "bad arguments");
  k8::DiclParams P;
  P.x = x; P.y = y; P.dout = nullptr; P.out = out; P.out2 = nullptr; P.mask = mask;
  P.B = B; P.C = C; P.H = H; P.W = W; P.U = 2 * max_u + 1; P.V = 2 * max_v + 1; P.mu = max_u; P.mv = max_v;
  P.remove_holes = remove_warp_hole;
  k8::dicl_volume_fwd_kernel<<<pixel_grid(H, W, (long long)B * P.U * P.V), 256, 0, static_cast<cudaStream_t>(stream)>>>(P);
  EZB_LAUNCH_OK();
  return EZB_OK;
}

extern "C" int ezb_dicl_volume_bwd(const float* dout, const unsigned char* mask, int B, int C, int H, int W, int max_u, int max_v,
                                   float* dx, float* dy, void* stream) {
  EZB_REQUIRE(dout && mask && (dx || dy), "null pointer");
  EZB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && max_u >= 0 && max_v >= 0, "bad arguments");
  k8::DiclParams P;
  P.x = nullptr; P.y = nullptr; P.dout = dout; P.out = dx; P.out2 = dy; P.mask = const_cast<unsigned char*>(mask);
  P.B = B; P.C = C; P.H = H; P.W = W; P.U = 2 * max_u + 1; P.V = 2 * max_v + 1; P.mu = max_u; P.mv = max_v;
  P.remove_holes = 0;
  k8::dicl_volume_bwd_kernel<<<pixel_grid(H, W, (long long)B * C), 256, 0, static_cast<cudaStream_t>(stream)>>>(P);
  EZB_LAUNCH_OK();
  return EZB_OK;
}

extern "C" int ezb_corr_ondemand_workspace_bytes(int B, int C, int H, int W, int L, size_t* bytes) {
  EZB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && L >= 1 && bytes, "bad arguments");
  if (L > PYR_MAX_LEVELS) return fail(EZB_ERR_UNSUPPORTED, "num_levels %d > %d is not supported", L, PYR_MAX_LEVELS);
  size_t n = 0;
  for (int l = 0; l < L; ++l) n += (size_t)B * (H >> l) * (W >> l) * C * 4 + 256;
  *bytes = n + (size_t)B * H * W * C * 4 + 512;
  return EZB_OK;
}

// prepare: fmap1 -> NHWC, fmap2 -> NHWC levels 0..L-1 (floor 2x2 average pools), all inside `workspace`
extern "C" int ezb_corr_ondemand_prepare(const float* fmap1, const float* fmap2, int B, int C, int H, int W, int L, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  EZB_REQUIRE(fmap1 && fmap2 && workspace, "null pointer");
  EZB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && L >= 1, "bad shape");
  if (L > PYR_MAX_LEVELS) return fail(EZB_ERR_UNSUPPORTED, "num_levels %d > %d is not supported", L, PYR_MAX_LEVELS);
  EZB_REQUIRE((H >> (L - 1)) >= 1 && (W >> (L - 1)) >= 1, "feature map %dx%d too small for %d pyramid levels", H, W, L);
  EZB_REQUIRE(B <= 65535, "batch too large for one launch (split it)");
  size_t need = 0;
  ezb_corr_ondemand_workspace_bytes(B, C, H, W, L, &need);
  if (need > workspace_bytes) return fail(EZB_ERR_WORKSPACE, "workspace too small: need %zu, got %zu", need, workspace_bytes);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int N = H * W;
  uintptr_t p = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
  float* f1n = reinterpret_cast<float*>(p);
  p += ((size_t)B * N * C * 4 + 255) & ~size_t(255);
  k8::nchw_to_nhwc_kernel<<<dim3(cdiv(N, 32), cdiv(C, 32), B), 256, 0, st>>>(fmap1, f1n, C, N);
  EZB_LAUNCH_OK();
  float* prev = nullptr;
  for (int l = 0; l < L; ++l) {
    const int hl = H >> l, wl = W >> l;
    float* cur = reinterpret_cast<float*>(p);
    p += ((size_t)B * hl * wl * C * 4 + 255) & ~size_t(255);
    if (l == 0) {
      k8::nchw_to_nhwc_kernel<<<dim3(cdiv(N, 32), cdiv(C, 32), B), 256, 0, st>>>(fmap2, cur, C, N);
    } else {
      const long long total = (long long)B * hl * wl * C;
      k8::pool_nhwc_kernel<<<(unsigned)std::min<long long>(cdiv64(total, 256), 148 * 32), 256, 0, st>>>(prev, cur, C, H >> (l - 1), W >> (l - 1), hl, wl,
                                                                                                      total);
    }
    EZB_LAUNCH_OK();
    prev = cur;
  }
  return EZB_OK;
}

extern "C" int ezb_corr_ondemand_lookup(const void* workspace, const float* coords, int B, int C, int H, int W, int L, int radius,
                                        float* out, void* stream) {
  EZB_REQUIRE(workspace && coords && out, "null pointer");
  EZB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && L >= 1 && radius >= 0 && radius <= 8, "bad arguments");
  if (L > PYR_MAX_LEVELS) return fail(EZB_ERR_UNSUPPORTED, "num_levels %d > %d is not supported", L, PYR_MAX_LEVELS);
  EZB_REQUIRE(B <= 65535, "batch too large for one launch (split it)");
  const int N = H * W;
  k8::OnDemandParams P;
  uintptr_t p = (reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255);
  P.f1 = reinterpret_cast<const float*>(p);
  p += ((size_t)B * N * C * 4 + 255) & ~size_t(255);
  for (int l = 0; l < PYR_MAX_LEVELS; ++l) {
    P.f2[l] = reinterpret_cast<const float*>(p);
    if (l < L) p += ((size_t)B * (H >> l) * (W >> l) * C * 4 + 255) & ~size_t(255);
  }
  P.coords = coords; P.out = out;
  P.B = B; P.C = C; P.H = H; P.W = W; P.N = N; P.L = L; P.r = radius;
  P.scale = 1.0f / sqrtf((float)C);
  cudaStream_t st0 = static_cast<cudaStream_t>(stream);
  if (2 * radius + 2 <= k8::ODT_MAXD && C % 4 == 0 && (long long)B * L <= 65535 && cdiv(H, k8::ODT_QH) <= 65535 &&
      getenv("EZB_ONDEMAND_UNTILED") == nullptr) {
    // shared-memory tiled kernel: neighbouring queries share the fmap2 pixels of their windows
    EZB_CUDA_OK(cudaFuncSetAttribute(k8::corr_ondemand_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k8::ODT_SMEM));
    k8::corr_ondemand_tiled_kernel<<<dim3(cdiv(W, k8::ODT_QW), cdiv(H, k8::ODT_QH), B * L), k8::ODT_THREADS, k8::ODT_SMEM, st0>>>(P);
    EZB_LAUNCH_OK();
    return EZB_OK;
  }
  const int qpb = std::max(1, k8::OD_WARPS / L), D = 2 * radius + 2;
  const size_t smem = ((size_t)qpb * C + (size_t)k8::OD_WARPS * D * D) * 4;
  if (smem > 200 * 1024) return fail(EZB_ERR_UNSUPPORTED, "feature dimension too large for the on-demand lookup (C=%d)", C);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EZB_CUDA_OK(cudaFuncSetAttribute(k8::corr_ondemand_lookup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k8::corr_ondemand_lookup_kernel<<<dim3(cdiv(N, qpb), B), k8::OD_WARPS * 32, smem, st>>>(P);
  EZB_LAUNCH_OK();
  return EZB_OK;
}
