// Device-side matching cost of the depth path: computeCost (Derp.cpp:104-226) with computeSSD
// (DerpUtil.cpp:126-162) and cv_util::getPixelBilinear (CvUtil.h:78-120) fused into one function.
//
// Arithmetic contract (bit-exactness with the reference CPU path):
//   * fp64 projection in the reference's operation order, -fmad=false;
//   * fp32 bilinear weights ((1-xw)*(1-yw))*p00 + (xw*(1-yw))*p01 + ((1-xw)*yw)*p10 + (xw*yw)*p11,
//     summed left to right, then truncated to integer (bilerp<ushort> returns ushort);
//   * SSD accumulation order dx outer / dy inner / channel 0..2, cv::Matx::dot order;
//   * robust camera mean in libstdc++'s nth_element order (derp_select.cuh).
//
// HBM layout: every colour image is W*H texels of 4 x u16 (B,G,R,0) = 8 B, so one texel is one
// aligned 64-bit load; warp tables are W*H x float2.
#pragma once

#include <cfloat>
#include <cstdint>

#include "derp_camera.cuh"
#include "derp_select.cuh"

namespace derp {

constexpr int kMaxCams = 32;                         // per-thread SSD arrays are sized by this
constexpr float kMinVarF = 1.0f / 12.0f / 65025.0f;  // DerpUtil.h:32

// One (frame, level, destination) as the cost function sees it.
struct CostView {
  int W, H, S, self;
  const uint2* projColor;   // [S][H][W] texels (self slot = the destination's own colour)
  const uint2* projBias;    // [S][H][W]
  const float2* projWarp;   // [S][H][W]  src px -> dst px at infinity (self slot unused)
  const float* variance;    // destination's own variance [H][W]
  const DevCamera* cams;    // [S] normalised cameras (global memory; staged to smem by kernels)
};

__device__ __forceinline__ int clampIdx(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

// u16 -> float without the conversion pipe: 0x4B000000 | u is the float 2^23 + u exactly.
__device__ __forceinline__ float u16lo(uint32_t w) {
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)) - 8388608.0f;
}
__device__ __forceinline__ float u16hi(uint32_t w) {
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)) - 8388608.0f;
}
// (float)(ushort)v for 0 <= v < 2^23: adding 2^23 with round-toward-zero drops the fraction.
__device__ __forceinline__ float truncU16(float v) { return __fadd_rz(v, 8388608.0f) - 8388608.0f; }

struct Texel {
  float b, g, r;
};
__device__ __forceinline__ Texel unpack(uint2 t) { return Texel{u16lo(t.x), u16hi(t.x), u16lo(t.y)}; }

// bilerp (CvUtil.h:83-86) for one channel, float result
__device__ __forceinline__ float bilerp1(float p00, float p01, float p10, float p11, float w00, float w01,
                                         float w10, float w11) {
  return w00 * p00 + w01 * p01 + w10 * p10 + w11 * p11;  // left-to-right, no FMA (-fmad=false)
}

// getPixelBilinear on a Vec3w image: per-channel truncated result (CvUtil.h:90-120)
__device__ __forceinline__ Texel sampleTexelTrunc(const uint2* __restrict__ img, int W, int H, float x, float y) {
  const float xf = roundf(x), yf = roundf(y);
  const int xi = (int)xf, yi = (int)yf;
  const int x0 = clampIdx(xi - 1, W - 1), x1 = clampIdx(xi, W - 1);
  const int y0 = clampIdx(yi - 1, H - 1), y1 = clampIdx(yi, H - 1);
  const float xw = x - xf + 0.5f, yw = y - yf + 0.5f;
  const float w00 = (1 - xw) * (1 - yw), w01 = xw * (1 - yw), w10 = (1 - xw) * yw, w11 = xw * yw;
  const Texel p00 = unpack(__ldg(img + (size_t)y0 * W + x0));
  const Texel p01 = unpack(__ldg(img + (size_t)y0 * W + x1));
  const Texel p10 = unpack(__ldg(img + (size_t)y1 * W + x0));
  const Texel p11 = unpack(__ldg(img + (size_t)y1 * W + x1));
  Texel o;
  o.b = truncU16(bilerp1(p00.b, p01.b, p10.b, p11.b, w00, w01, w10, w11));
  o.g = truncU16(bilerp1(p00.g, p01.g, p10.g, p11.g, w00, w01, w10, w11));
  o.r = truncU16(bilerp1(p00.r, p01.r, p10.r, p11.r, w00, w01, w10, w11));
  return o;
}

// getPixelBilinear on a Vec2f image (float result, no truncation)
__device__ __forceinline__ float2 sampleWarp(const float2* __restrict__ img, int W, int H, float x, float y) {
  const float xf = roundf(x), yf = roundf(y);
  const int xi = (int)xf, yi = (int)yf;
  const int x0 = clampIdx(xi - 1, W - 1), x1 = clampIdx(xi, W - 1);
  const int y0 = clampIdx(yi - 1, H - 1), y1 = clampIdx(yi, H - 1);
  const float xw = x - xf + 0.5f, yw = y - yf + 0.5f;
  const float w00 = (1 - xw) * (1 - yw), w01 = xw * (1 - yw), w10 = (1 - xw) * yw, w11 = xw * yw;
  const float2 p00 = __ldg(img + (size_t)y0 * W + x0);
  const float2 p01 = __ldg(img + (size_t)y0 * W + x1);
  const float2 p10 = __ldg(img + (size_t)y1 * W + x0);
  const float2 p11 = __ldg(img + (size_t)y1 * W + x1);
  float2 o;
  o.x = bilerp1(p00.x, p01.x, p10.x, p11.x, w00, w01, w10, w11);
  o.y = bilerp1(p00.y, p01.y, p10.y, p11.y, w00, w01, w10, w11);
  return o;
}

// getPixelBilinear on a float image
__device__ __forceinline__ float sampleF32(const float* __restrict__ img, int W, int H, float x, float y) {
  const float xf = roundf(x), yf = roundf(y);
  const int xi = (int)xf, yi = (int)yf;
  const int x0 = clampIdx(xi - 1, W - 1), x1 = clampIdx(xi, W - 1);
  const int y0 = clampIdx(yi - 1, H - 1), y1 = clampIdx(yi, H - 1);
  const float xw = x - xf + 0.5f, yw = y - yf + 0.5f;
  const float w00 = (1 - xw) * (1 - yw), w01 = xw * (1 - yw), w10 = (1 - xw) * yw, w11 = xw * yw;
  return bilerp1(img[(size_t)y0 * W + x0], img[(size_t)y0 * W + x1], img[(size_t)y1 * W + x0],
                 img[(size_t)y1 * W + x1], w00, w01, w10, w11);
}

// Candidate-independent state of one destination pixel.
struct PixelState {
  float cD[27];     // dst colour patch, [dx+1][dy+1][channel] (computeSSD loop order)
  float dBias[3];   // projColorBias(dst,self)(y,x)
  float conf;       // max(variance(y,x), kMinVar)
  double dir[3];    // ray direction of the pixel in rig space
  double org[3];    // destination camera position
};

__device__ __forceinline__ void loadPixelState(const CostView& v, const DevCamera& camDst, int x, int y,
                                               PixelState& ps) {
  const uint2* col = v.projColor + (size_t)v.self * v.W * v.H;
#pragma unroll
  for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const Texel t = unpack(__ldg(col + (size_t)(y + dy) * v.W + (x + dx)));
      const int k = ((dx + 1) * 3 + (dy + 1)) * 3;
      ps.cD[k] = t.b;
      ps.cD[k + 1] = t.g;
      ps.cD[k + 2] = t.r;
    }
  const Texel tb = unpack(__ldg(v.projBias + (size_t)v.self * v.W * v.H + (size_t)y * v.W + x));
  ps.dBias[0] = tb.b;
  ps.dBias[1] = tb.g;
  ps.dBias[2] = tb.r;
  ps.conf = fmaxf(__ldg(v.variance + (size_t)y * v.W + x), kMinVarF);
  // dstToWorldPoint (DerpUtil.cpp:38-52): normalised pixel centre, ray through it
  const double px = (x + 0.5) / v.W, py = (y + 0.5) / v.H;
  pixelRay(camDst, px, py, ps.dir);
  ps.org[0] = camDst.pos[0];
  ps.org[1] = camDst.pos[1];
  ps.org[2] = camDst.pos[2];
}

// computeSSD (DerpUtil.cpp:126-162) against one projected source
__device__ __forceinline__ void ssdAgainst(const uint2* __restrict__ srcColor, int W, int H, const PixelState& ps,
                                           float xDstSrc, float yDstSrc, const Texel& srcBias, float* ssdB,
                                           float* ssdU) {
  const float bias0 = ps.dBias[0] - srcBias.b, bias1 = ps.dBias[1] - srcBias.g, bias2 = ps.dBias[2] - srcBias.r;
  float sB = 0.0f, sU = 0.0f;
#pragma unroll
  for (int dx = -1; dx <= 1; ++dx) {
    const float xs = xDstSrc + (float)dx;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const Texel cs = sampleTexelTrunc(srcColor, W, H, xs, yDstSrc + (float)dy);
      const int k = ((dx + 1) * 3 + (dy + 1)) * 3;
      const float d0 = ps.cD[k] - cs.b, d1 = ps.cD[k + 1] - cs.g, d2 = ps.cD[k + 2] - cs.r;
      const float u0 = d0 - bias0, u1 = d1 - bias1, u2 = d2 - bias2;
      sB += d0 * d0 + d1 * d1 + d2 * d2;
      sU += u0 * u0 + u1 * u1 + u2 * u2;
    }
  }
  const float scaleFactor = 1.0f / (65535.0f * 65535.0f);
  *ssdB = sB * scaleFactor;
  *ssdU = sU * scaleFactor;
}

// computeCost (Derp.cpp:104-226).  `cams` should point to shared memory.  Returns the cost;
// confidence is ps.conf when the return value is not FLT_MAX, 0 otherwise.
__device__ __forceinline__ float evalCost(const CostView& v, const DevCamera* __restrict__ cams,
                                          const PixelState& ps, float disparity, unsigned* hits) {
  const double depth = (double)(1.0f / disparity);
  const double wx = ps.org[0] + ps.dir[0] * depth;
  const double wy = ps.org[1] + ps.dir[1] * depth;
  const double wz = ps.org[2] + ps.dir[2] * depth;
  float ssdB[kMaxCams], ssdU[kMaxCams];
  int n = 0;
  const size_t plane = (size_t)v.W * v.H;
  for (int s = 0; s < v.S; ++s) {
    if (s == v.self) continue;
    double px, py;
    if (!sees(cams[s], wx, wy, wz, &px, &py)) continue;
    px *= v.W;  // worldToSrcPoint: de-normalise (DerpUtil.cpp:67-71)
    py *= v.H;
    const float2 pd = sampleWarp(v.projWarp + s * plane, v.W, v.H, (float)px, (float)py);
    const float xDstSrc = pd.x + 0.5f, yDstSrc = pd.y + 0.5f;
    if (isnan(xDstSrc) || isnan(yDstSrc)) continue;
    const Texel srcBias = sampleTexelTrunc(v.projBias + s * plane, v.W, v.H, xDstSrc, yDstSrc);
    ssdAgainst(v.projColor + s * plane, v.W, v.H, ps, xDstSrc, yDstSrc, srcBias, &ssdB[n], &ssdU[n]);
    ++n;
  }
  *hits += n;
  if (n < 1) return FLT_MAX;  // kMinOverlappingCams - 1
  const int keep = n - 2 > 1 ? n - 2 : 1;
  float cost;
  if (n == 1) {
    cost = 0.0f + ssdU[0];
  } else {
    cost = robustSum(ssdB, ssdU, n, keep);
  }
  cost /= (float)keep;
  const float trustCoef = 1.0f / (float)keep;
  return cost * trustCoef / ps.conf;
}

}  // namespace derp
