// K5: backward feature warping used by PWC-Net before each local correlation.
//
// Replaces warp (reference ezflow/utils/warp.py:5-42):
//   out = grid_sample(x, pixel_grid + flow, bilinear, zeros, align_corners=True) * mask,
//   mask = grid_sample(ones, same grid) binarised (mask < 0.9999 -> 0, else 1).
// One thread per output pixel computes the four taps once and sweeps the channels (reads and writes
// are coalesced along x).  HBM-bound: C*(4 B read + 4 B write) + 8 B of flow per pixel.
#include "ezb_common.cuh"
#include "ezb_warp.cuh"

namespace ezb {
namespace k5 {

// thread = pixel, blockIdx.y = chunk of CCH channels: the taps (weights folded with validity and the mask, addresses
// clamped into the image) are computed once per thread, then the channel loop is 4 loads + 4 FMAs + 1 store, unrolled so
// that several channels' loads are in flight.  Reads and writes are coalesced along x.
constexpr int CCH = 16;
__global__ void __launch_bounds__(128) warp_fwd_kernel(const float* __restrict__ x, const float* __restrict__ flow,
                                                       float* __restrict__ out, int B, int C, int H, int W) {
  const long long HW = (long long)H * W, total = (long long)B * HW;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = (int)(i / HW);
  const long long p = i - (long long)b * HW;
  const int py = (int)(p / W), px = (int)(p - (long long)py * W);
  const Taps t = make_taps(__ldg(flow + (long long)b * 2 * HW + p), __ldg(flow + (long long)b * 2 * HW + HW + p), px, py, H, W);
  const int c0 = blockIdx.y * CCH, c1 = min(C, c0 + CCH);
  float* dst = out + ((long long)b * C + c0) * HW + p;
  if (t.mask == 0.f) {
    for (int c = c0; c < c1; ++c, dst += HW) *dst = 0.f;
    return;
  }
  const Gather g = make_gather(t, W);
  const float* s = x + ((long long)b * C + c0) * HW;
#pragma unroll 4
  for (int c = c0; c < c1; ++c, s += HW, dst += HW)
    *dst = (g.w00 * __ldg(s + g.oaa) + g.w01 * __ldg(s + g.oab)) + (g.w10 * __ldg(s + g.oba) + g.w11 * __ldg(s + g.obb));
}

// dx (pre-zeroed) += scatter of dout*mask through the bilinear taps; dflow = d(sample)/d(position).
// The binarised mask carries no gradient (it is overwritten by constants in the reference).
__global__ void warp_bwd_kernel(const float* __restrict__ x, const float* __restrict__ flow, const float* __restrict__ dout,
                                float* __restrict__ dx, float* __restrict__ dflow, int B, int C, int H, int W) {
  const long long HW = (long long)H * W, total = (long long)B * HW;
  const float sx = (W > 1) ? 1.f : 0.f, sy = (H > 1) ? 1.f : 0.f;  // (2/max(W-1,1)) * ((W-1)/2)
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / HW);
    const long long p = i - (long long)b * HW;
    const int py = (int)(p / W), px = (int)(p - (long long)py * W);
    const Taps t = make_taps(__ldg(flow + (long long)b * 2 * HW + p), __ldg(flow + (long long)b * 2 * HW + HW + p), px, py, H, W);
    float gx = 0.f, gy = 0.f;
    if (t.mask != 0.f) {
      const long long o00 = (long long)t.y0 * W + t.x0;
      for (int c = 0; c < C; ++c) {
        const float g = __ldg(dout + ((long long)b * C + c) * HW + p);
        const long long cb = ((long long)b * C + c) * HW + o00;
        const float v00 = (t.vy0 && t.vx0) ? __ldg(x + cb) : 0.f;
        const float v01 = (t.vy0 && t.vx1) ? __ldg(x + cb + 1) : 0.f;
        const float v10 = (t.vy1 && t.vx0) ? __ldg(x + cb + W) : 0.f;
        const float v11 = (t.vy1 && t.vx1) ? __ldg(x + cb + W + 1) : 0.f;
        gx += g * (t.wy0 * (v01 - v00) + t.wy1 * (v11 - v10));
        gy += g * (t.wx0 * (v10 - v00) + t.wx1 * (v11 - v01));
        if (dx) {
          if (t.vy0 && t.vx0) atomicAdd(dx + cb, g * t.wy0 * t.wx0);
          if (t.vy0 && t.vx1) atomicAdd(dx + cb + 1, g * t.wy0 * t.wx1);
          if (t.vy1 && t.vx0) atomicAdd(dx + cb + W, g * t.wy1 * t.wx0);
          if (t.vy1 && t.vx1) atomicAdd(dx + cb + W + 1, g * t.wy1 * t.wx1);
        }
      }
    }
    if (dflow) {
      dflow[(long long)b * 2 * HW + p] = gx * sx;
      dflow[(long long)b * 2 * HW + HW + p] = gy * sy;
    }
  }
}

}  // namespace k5
}  // namespace ezb

using namespace ezb;

extern "C" int ezb_warp_fwd(const float* x, const float* flow, int B, int C, int H, int W, float* out, void* stream) {
  EZB_REQUIRE(x && flow && out, "null pointer");
  EZB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  const long long total = (long long)B * H * W;
  EZB_REQUIRE(cdiv64(total, 128) < (1ll << 31) && cdiv(C, k5::CCH) <= 65535, "tensor too large for one launch");
  k5::warp_fwd_kernel<<<dim3((unsigned)cdiv64(total, 128), cdiv(C, k5::CCH)), 128, 0, static_cast<cudaStream_t>(stream)>>>(x, flow, out, B, C,
                                                                                                                         H, W);
  EZB_LAUNCH_OK();
  return EZB_OK;
}

extern "C" int ezb_warp_bwd(const float* x, const float* flow, const float* dout, int B, int C, int H, int W, float* dx,
                            float* dflow, void* stream) {
  EZB_REQUIRE(x && flow && dout && (dx || dflow), "null pointer");
  EZB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dx) EZB_CUDA_OK(cudaMemsetAsync(dx, 0, (size_t)B * C * H * W * sizeof(float), st));
  const long long total = (long long)B * H * W;
  const int blocks = (int)std::min<long long>(cdiv64(total, 128), 148 * 32);
  k5::warp_bwd_kernel<<<blocks, 128, 0, st>>>(x, flow, dout, dx, dflow, B, C, H, W);
  EZB_LAUNCH_OK();
  return EZB_OK;
}
