// Bilinear taps of the PWC-Net feature warp (reference ezflow/utils/warp.py:5-42), shared by the K5 kernels (warp.cu) and
// the warp-fused operand conversion of the local correlation (local_corr_tc.cu).
#pragma once

#include "ezb_common.cuh"

namespace ezb {
namespace k5 {

struct Taps {
  int x0, y0;
  float wx0, wx1, wy0, wy1;
  bool vx0, vx1, vy0, vy1;
  float mask;
};

__device__ __forceinline__ Taps make_taps(float fx, float fy, int x, int y, int H, int W) {
  // follow the reference's normalise -> grid_sample un-normalise round trip (warp.py:32-33)
  const float gx = 2.0f * ((float)x + fx) / (float)max(W - 1, 1) - 1.0f;
  const float gy = 2.0f * ((float)y + fy) / (float)max(H - 1, 1) - 1.0f;
  const float ix = ((gx + 1.f) / 2.f) * (float)(W - 1);
  const float iy = ((gy + 1.f) / 2.f) * (float)(H - 1);
  Taps t;
  const float x0f = floorf(ix), y0f = floorf(iy);
  t.wx1 = ix - x0f;
  t.wx0 = 1.f - t.wx1;
  t.wy1 = iy - y0f;
  t.wy0 = 1.f - t.wy1;
  const float xc = fminf(fmaxf(x0f, -2.f), (float)W), yc = fminf(fmaxf(y0f, -2.f), (float)H);
  t.x0 = (xc == xc) ? (int)xc : -2;
  t.y0 = (yc == yc) ? (int)yc : -2;
  t.vx0 = t.x0 >= 0 && t.x0 < W;
  t.vx1 = t.x0 + 1 >= 0 && t.x0 + 1 < W;
  t.vy0 = t.y0 >= 0 && t.y0 < H;
  t.vy1 = t.y0 + 1 >= 0 && t.y0 + 1 < H;
  float m = 0.f;
  if (t.vy0 && t.vx0) m += t.wy0 * t.wx0;
  if (t.vy0 && t.vx1) m += t.wy0 * t.wx1;
  if (t.vy1 && t.vx0) m += t.wy1 * t.wx0;
  if (t.vy1 && t.vx1) m += t.wy1 * t.wx1;
  t.mask = (m < 0.9999f) ? 0.f : 1.f;  // mask<0.9999 -> 0 ; mask>0 -> 1   (warp.py:39-40)
  return t;
}

// The four taps with validity and the binarised mask folded into the weights and the addresses clamped into the image
// (an invalid tap reads a valid neighbour with weight 0): value = (w00*s[oaa] + w01*s[oab]) + (w10*s[oba] + w11*s[obb]).
struct Gather {
  float w00, w01, w10, w11;
  int oaa, oab, oba, obb;
  bool zero;  // mask == 0: the warped pixel is 0 in every channel
};

__device__ __forceinline__ Gather make_gather(const Taps& t, int W) {
  Gather g;
  g.zero = t.mask == 0.f;
  g.w00 = (t.vy0 && t.vx0) ? t.wy0 * t.wx0 : 0.f;
  g.w01 = (t.vy0 && t.vx1) ? t.wy0 * t.wx1 : 0.f;
  g.w10 = (t.vy1 && t.vx0) ? t.wy1 * t.wx0 : 0.f;
  g.w11 = (t.vy1 && t.vx1) ? t.wy1 * t.wx1 : 0.f;
  const int ya = t.vy0 ? t.y0 : t.y0 + 1, yb = t.vy1 ? t.y0 + 1 : t.y0;
  const int xa = t.vx0 ? t.x0 : t.x0 + 1, xb = t.vx1 ? t.x0 + 1 : t.x0;
  g.oaa = ya * W + xa;
  g.oab = ya * W + xb;
  g.oba = yb * W + xa;
  g.obb = yb * W + xb;
  return g;
}

}  // namespace k5
}  // namespace ezb
