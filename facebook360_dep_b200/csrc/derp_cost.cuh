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
// HBM layout: the source colour images are W*H texels of 4 x u16 (B,G,R,0) = 8 B; the per-destination
// pair tables the cost reads (projColor, projBias) hold the same integer values pre-converted to
// float4 (B,G,R,0) = one aligned 128-bit load per texel and no u16->f32 conversion in the inner loop
// (the sweep is instruction-issue bound, not bandwidth bound: profiles/README.md); warp tables are
// W*H x float2.
#pragma once

#include <cfloat>
#include <cstdint>

// Table-driven robust-mean selection (derp_select.cuh::robustSumTable): measured on B200 at 2048^2 x 128 candidates
// 94.0 -> 87.0 ms per sweep launch (25.7 -> 27.8 G triples/s), all parity tests bit-exact; -DDERP_NO_SELECT_TABLE
// builds the general introselect emulation only (the A/B baseline).
#if !defined(DERP_NO_SELECT_TABLE) && !defined(DERP_SELECT_TABLE)
#define DERP_SELECT_TABLE 1
#endif

#include "derp_camera.cuh"
#include "derp_select.cuh"

namespace derp {

constexpr int kMaxCams = 32;                         // per-thread SSD arrays are sized by this
constexpr float kMinVarF = 1.0f / 12.0f / 65025.0f;  // DerpUtil.h:32

// One (frame, level, destination) as the cost function sees it.
struct CostView {
  int W, H, S, self;
  const float4* projColor;  // [S][H][W] texels (self slot = the destination's own colour)
  const float4* projBias;   // [S][H][W]
  // The same two tables as 4 x u16 (B,G,R,R of the texel below) = 8 B per texel, for the compacted fine-level
  // kernels: their gathers are scattered (list entries of one warp span several image rows) and L1-tag bound, and
  // half-size texels halve the cache lines a request touches; the conversion costs 4 instructions per texel.
  const uint2* projColor16;
  const uint2* projBias16;
#ifdef DERP_SELECT_TABLE
  const unsigned* selTab;  // robustSumTable's permutation table (derp_select.cuh)
#endif
  const float2* projWarp;   // [S][H][W]  src px -> dst px at infinity (self slot unused)
  const float* variance;    // destination's own variance [H][W]
  const DevCamera* cams;    // [S] normalised cameras (global memory; staged to smem by kernels)
  // 1.0f and 2^23 passed as kernel parameters: they reach the packed instructions as constant-bank /
  // uniform-register operands, and because ptxas cannot see their values it can neither fold
  // fma(p, one, q) back into an add nor contract it with the multiply that produced p (see f32x2 notes).
  float one, b23;
};

__device__ __forceinline__ int clampIdx(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

// u16 -> float without the conversion pipe: 0x4B000000 | u is the float 2^23 + u exactly.
__device__ __forceinline__ float u16lo(uint32_t w) {
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)) - 8388608.0f;
}
__device__ __forceinline__ float u16hi(uint32_t w) {
  return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)) - 8388608.0f;
}
// (ushort)v for 0 <= v < 2^23, returned as the float 2^23 + trunc(v): adding 2^23 with
// round-toward-zero drops the fraction.  The cost only ever uses DIFFERENCES of such truncated values
// and integer texels, so everything on the dst side is biased by 2^23 as well (exact: all values are
// integers below 2^24) and the second FADD is never needed.
constexpr float kBias23 = 8388608.0f;
__device__ __forceinline__ float truncBiased(float v) { return __fadd_rz(v, kBias23); }

struct Texel {
  float b, g, r;
};
__device__ __forceinline__ Texel unpack(uint2 t) { return Texel{u16lo(t.x), u16hi(t.x), u16lo(t.y)}; }

// bilerp (CvUtil.h:83-86) for one channel, float result
__device__ __forceinline__ float bilerp1(float p00, float p01, float p10, float p11, float w00, float w01,
                                         float w10, float w11) {
  return w00 * p00 + w01 * p01 + w10 * p10 + w11 * p11;  // left-to-right, no FMA (-fmad=false)
}

__device__ __forceinline__ Texel texelOf(float4 t) { return Texel{t.x, t.y, t.z}; }

// One table texel as float4 (B, G, R, R-below), from either table format.  u16 -> f32 is exact.
__device__ __forceinline__ float4 ldTexel(const float4* p) { return __ldg(p); }
__device__ __forceinline__ float4 ldTexel(const uint2* p) {
  const uint2 t = __ldg(p);
#ifdef DERP_U16_PRMT
  // u16 -> f32 without the conversion unit: one byte permute builds the float 2^23 + u (0x4B00'uuuu), one add removes
  // the 2^23 — two full-rate instructions instead of a quarter-rate I2F; exact for every u < 2^16
  const float b = 8388608.0f;
  return make_float4(__uint_as_float(__byte_perm(t.x, 0x4B000000u, 0x7610)) - b,
                     __uint_as_float(__byte_perm(t.x, 0x4B000000u, 0x7632)) - b,
                     __uint_as_float(__byte_perm(t.y, 0x4B000000u, 0x7610)) - b,
                     __uint_as_float(__byte_perm(t.y, 0x4B000000u, 0x7632)) - b);
#else
  return make_float4((float)(t.x & 0xffffu), (float)(t.x >> 16), (float)(t.y & 0xffffu), (float)(t.y >> 16));
#endif
}
template <class TX>
struct TablesOf;
template <>
struct TablesOf<float4> {
  static __device__ __forceinline__ const float4* color(const CostView& v) { return v.projColor; }
  static __device__ __forceinline__ const float4* bias(const CostView& v) { return v.projBias; }
};
template <>
struct TablesOf<uint2> {
  static __device__ __forceinline__ const uint2* color(const CostView& v) { return v.projColor16; }
  static __device__ __forceinline__ const uint2* bias(const CostView& v) { return v.projBias16; }
};

// getPixelBilinear on a Vec3w image: per-channel truncated result (CvUtil.h:90-120), biased by 2^23.
// Generic path: per-tap clamp-to-edge.
template <class TX>
__device__ __forceinline__ Texel sampleTexelTruncBiased(const TX* __restrict__ img, int W, int H, float x, float y) {
  const float xf = roundf(x), yf = roundf(y);
  const int xi = (int)xf, yi = (int)yf;
  const int x0 = clampIdx(xi - 1, W - 1), x1 = clampIdx(xi, W - 1);
  const int y0 = clampIdx(yi - 1, H - 1), y1 = clampIdx(yi, H - 1);
  const float xw = x - xf + 0.5f, yw = y - yf + 0.5f;
  const float w00 = (1 - xw) * (1 - yw), w01 = xw * (1 - yw), w10 = (1 - xw) * yw, w11 = xw * yw;
  const Texel p00 = texelOf(ldTexel(img + (size_t)y0 * W + x0));
  const Texel p01 = texelOf(ldTexel(img + (size_t)y0 * W + x1));
  const Texel p10 = texelOf(ldTexel(img + (size_t)y1 * W + x0));
  const Texel p11 = texelOf(ldTexel(img + (size_t)y1 * W + x1));
  Texel o;
  o.b = truncBiased(bilerp1(p00.b, p01.b, p10.b, p11.b, w00, w01, w10, w11));
  o.g = truncBiased(bilerp1(p00.g, p01.g, p10.g, p11.g, w00, w01, w10, w11));
  o.r = truncBiased(bilerp1(p00.r, p01.r, p10.r, p11.r, w00, w01, w10, w11));
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

// ---- packed fp32x2 arithmetic (sm_100a FADD2 / FMUL2 / FFMA2) ---------------------------------------------
// The sweep is instruction-issue bound and ~40 % of its instructions are the fp32 mul/add of the
// truncated-bilinear SSD.  Blackwell's packed f32x2 instructions do two IEEE fp32 operations per issue slot
// with per-lane rounding, so results stay bit-identical to the scalar reference arithmetic, PROVIDED no
// multiply is contracted into a following add.  ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2
// even with explicit .rn and -fmad=false (verified on SASS), so every add whose operand is a product is
// written as fma(p, one, q) with an opaque one (CostView::one): RN(p*1+q) == RN(p+q) exactly, and an FMA
// cannot absorb a second multiply.  Adds of non-products use add/sub directly.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float lo2(f32x2 v) {
  float a, b;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
  (void)b;
  return a;
}
__device__ __forceinline__ float hi2(f32x2 v) {
  float a, b;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
  (void)a;
  return b;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {  // operands must NOT be products
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {  // operands must NOT be products
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 addrz2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 addp2(f32x2 p, f32x2 q, f32x2 one2) {  // p + q where p and/or q are products
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(p), "l"(one2), "l"(q));
  return r;
}
__device__ __forceinline__ float addp(float p, float q, float one) { return __fmaf_rn(p, one, q); }

// bilerp of CvUtil.h:83-86 on two lanes: ((w00*p00 + w01*p01) + w10*p10) + w11*p11, per-lane RN after every op
__device__ __forceinline__ f32x2 bilerp2(f32x2 p00, f32x2 p01, f32x2 p10, f32x2 p11, f32x2 w00, f32x2 w01, f32x2 w10,
                                         f32x2 w11, f32x2 one2) {
  f32x2 s = addp2(mul2(p01, w01), mul2(p00, w00), one2);
  s = addp2(mul2(p10, w10), s, one2);
  return addp2(mul2(p11, w11), s, one2);
}

// ---- destination patch tile in shared memory ----------------------------------------------------------
// The 3x3 dst colour patches of a CTA's 32x8 pixels overlap: one (32+2)x(8+2) tile of the destination's
// own colour (+2^23, see truncBiased) serves all of them and keeps 27 values per thread out of registers.
// Two float2 planes so that the packed arithmetic gets its operands with 64-bit shared loads:
//   bg[row][col] = (B, G),   rr[row][col] = (R(row), R(row+1))   (vertical pair: see the R channel below)
constexpr int kTileW = 32 + 2;
#ifndef DERP_SWEEP_MAXBY
#define DERP_SWEEP_MAXBY 20  // tallest CTA the sweep is launched with (32 x 20 threads => 96 registers, 20 warps/SM)
#endif
constexpr int kMaxTileH = DERP_SWEEP_MAXBY + 2;
constexpr int kTileFloats = 2 * kMaxTileH * kTileW * 2;

// `addend`: 2^23 for the exact cost (see truncBiased), 0.5 for the lower-bound pass (midpoint of the truncation interval).
__device__ __forceinline__ void loadDstTile(float* tile, const CostView& v, int x0, int y0, float addend = kBias23) {
  const float4* col = v.projColor + (size_t)v.self * v.W * v.H;
  float2* bg = reinterpret_cast<float2*>(tile);
  const int tileH = blockDim.y + 2;
  float2* rr = bg + tileH * kTileW;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nt = blockDim.x * blockDim.y;
  for (int i = tid; i < kTileW * tileH; i += nt) {
    const int ty = i / kTileW, tx = i - ty * kTileW;
    const int gx = clampIdx(x0 + tx - 1, v.W - 1), gy = clampIdx(y0 + ty - 1, v.H - 1);
    const int gy1 = clampIdx(y0 + ty, v.H - 1);
    const float4 t = __ldg(col + (size_t)gy * v.W + gx);
    const float4 t1 = __ldg(col + (size_t)gy1 * v.W + gx);
    bg[i] = make_float2(t.x + addend, t.y + addend);
    rr[i] = make_float2(t.z + addend, t1.z + addend);
  }
  __syncthreads();
}

// Candidate-independent state of one destination pixel.
struct PixelState {
  const float2* bg;    // &bg[threadIdx.y][threadIdx.x]: patch(dx,dy) = bg[(dy+1)*kTileW + dx+1]
  const float2* rr;    // same indexing, (R(dy), R(dy+1))
  float dBias[3];      // projColorBias(dst,self)(y,x) + 2^23
  float conf;          // max(variance(y,x), kMinVar)
  double dir[3];       // ray direction of the pixel in rig space
  float2* sel;         // this thread's (ssdB, ssdU) slots in shared memory: entry i at sel[i * selStride]
  int selStride;       // = threads per CTA
};

// The per-source (biased, unbiased) SSD pairs of one cost evaluation live in shared memory, [slot][thread]: S - 1
// slots per thread, sized at launch (16 cameras, 640-thread sweep CTA: 75 KB).  kSelSlots = how many of them the
// table-driven selection covers.
constexpr int kSelSlots = 8;
struct SmemPairs {
  float2* p;
  int stride;
  __device__ __forceinline__ PairVal get(int i) const {
    const float2 t = p[i * stride];
    return PairVal{t.x, t.y};
  }
  __device__ __forceinline__ void set(int i, PairVal v) const { p[i * stride] = make_float2(v.a, v.b); }
};

__device__ __forceinline__ void loadPixelState(const CostView& v, const DevCamera& camDst, const float* tile, int x,
                                               int y, PixelState& ps, float addend = kBias23) {
  ps.bg = reinterpret_cast<const float2*>(tile) + threadIdx.y * kTileW + threadIdx.x;
  ps.rr = ps.bg + (blockDim.y + 2) * kTileW;
  ps.selStride = blockDim.x * blockDim.y;
  ps.sel = reinterpret_cast<float2*>(const_cast<float*>(tile) + kTileFloats) + threadIdx.y * blockDim.x + threadIdx.x;
  const float4 tb = __ldg(v.projBias + (size_t)v.self * v.W * v.H + (size_t)y * v.W + x);
  ps.dBias[0] = tb.x + addend;
  ps.dBias[1] = tb.y + addend;
  ps.dBias[2] = tb.z + addend;
  ps.conf = fmaxf(__ldg(v.variance + (size_t)y * v.W + x), kMinVarF);
  // dstToWorldPoint (DerpUtil.cpp:38-52): normalised pixel centre, ray through it
  const double px = (x + 0.5) / v.W, py = (y + 0.5) / v.H;
  pixelRay(camDst, px, py, ps.dir);
}

// Compacted kernels (one thread per ACTIVE pixel, pixels of a CTA are not a rectangle): every thread keeps its
// own 3x3 patch in shared memory, laid out [row][col][thread] so that a warp reads consecutive words.
#ifndef DERP_PATCH_THREADS
#define DERP_PATCH_THREADS 256
#endif
#ifndef DERP_PATCH_MINB
// Resident CTAs per SM the compacted kernels are compiled for (register cap).  2 => 128 registers (4 - 56 bytes of
// spills since the table-driven selection), 16 warps/SM: these kernels are L1-bound (scattered 4x4 gathers,
// profiles/README.md), so spill traffic costs more than the lost warps: at 2048^2 proposals 0.98 -> 0.63 ms,
// ping-pong 5.93 -> 5.51 ms against 3 CTAs / 80 registers (round 1); 96-register shapes re-measured in round 2
// (profiles/README.md, "compacted kernels: CTA shape").
#define DERP_PATCH_MINB 2
#endif
constexpr int kPatchThreads = DERP_PATCH_THREADS;
constexpr int kPatchRP = 3 * kPatchThreads, kPatchCP = kPatchThreads;
constexpr int kPatchFloats = 2 * 9 * kPatchThreads * 2;
// pingPongKernel has its own CTA shape: 128 threads x 5 CTAs per SM (96 registers, 20 warps/SM) measured 4.40 ms against
// 4.65 ms for 256 x 2 at 2048^2, while proposalKernel is faster at 256 x 2 (0.52 vs 0.62 ms: it spills more at 96).
#ifndef DERP_PING_THREADS
#define DERP_PING_THREADS 128
#endif
#ifndef DERP_PING_MINB
#define DERP_PING_MINB 5
#endif
constexpr int kPingThreads = DERP_PING_THREADS;

template <int T>
__device__ __forceinline__ void loadPixelStateCompact(const CostView& v, const DevCamera& camDst, float* patches, int x,
                                                      int y, PixelState& ps) {
  const int tid = threadIdx.x;
  float2* bg = reinterpret_cast<float2*>(patches) + tid;
  float2* rr = bg + 9 * T;
  const uint2* col = v.projColor16 + (size_t)v.self * v.W * v.H;  // u16 -> f32 is exact
  float rz[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {  // rows y-1 .. y+1 (interior pixel: always in bounds)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float4 t = ldTexel(col + (size_t)(y - 1 + r) * v.W + (x - 1 + c));
      bg[r * (3 * T) + c * T] = make_float2(t.x + kBias23, t.y + kBias23);
      // rr[r] = (R(row r), R(row r+1)); the fast path reads rr[0] (both lanes) and rr[2].x, the slow path rr[r].x
      if (r >= 1) rr[(r - 1) * (3 * T) + c * T] = make_float2(rz[c] + kBias23, t.z + kBias23);
      if (r == 2) rr[2 * (3 * T) + c * T] = make_float2(t.z + kBias23, 0.f);
      rz[c] = t.z;
    }
  }
  ps.bg = bg;
  ps.rr = rr;
  ps.selStride = T;
  ps.sel = reinterpret_cast<float2*>(patches + 2 * 9 * T * 2) + tid;
  const float4 tb = ldTexel(v.projBias16 + (size_t)v.self * v.W * v.H + (size_t)y * v.W + x);
  ps.dBias[0] = tb.x + kBias23;
  ps.dBias[1] = tb.y + kBias23;
  ps.dBias[2] = tb.z + kBias23;
  ps.conf = fmaxf(__ldg(v.variance + (size_t)y * v.W + x), kMinVarF);
  const double px = (x + 0.5) / v.W, py = (y + 0.5) / v.H;
  pixelRay(camDst, px, py, ps.dir);
}

// The same for the refine pass of the filtered sweep: list entries of the DENSE level, so patches and bias come from the
// float4 tables the sweep already built.
__device__ __forceinline__ void loadPixelStateCompactF32(const CostView& v, const DevCamera& camDst, float* patches, int x,
                                                         int y, PixelState& ps) {
  const int tid = threadIdx.x;
  float2* bg = reinterpret_cast<float2*>(patches) + tid;
  float2* rr = bg + 9 * kPatchThreads;
  const float4* col = v.projColor + (size_t)v.self * v.W * v.H;
  float rz[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float4 t = __ldg(col + (size_t)(y - 1 + r) * v.W + (x - 1 + c));
      bg[r * kPatchRP + c * kPatchCP] = make_float2(t.x + kBias23, t.y + kBias23);
      if (r >= 1) rr[(r - 1) * kPatchRP + c * kPatchCP] = make_float2(rz[c] + kBias23, t.z + kBias23);
      if (r == 2) rr[2 * kPatchRP + c * kPatchCP] = make_float2(t.z + kBias23, 0.f);
      rz[c] = t.z;
    }
  }
  ps.bg = bg;
  ps.rr = rr;
  ps.selStride = kPatchThreads;
  ps.sel = reinterpret_cast<float2*>(patches + kPatchFloats) + tid;
  const float4 tb = __ldg(v.projBias + (size_t)v.self * v.W * v.H + (size_t)y * v.W + x);
  ps.dBias[0] = tb.x + kBias23;
  ps.dBias[1] = tb.y + kBias23;
  ps.dBias[2] = tb.z + kBias23;
  ps.conf = fmaxf(__ldg(v.variance + (size_t)y * v.W + x), kMinVarF);
  const double px = (x + 0.5) / v.W, py = (y + 0.5) / v.H;
  pixelRay(camDst, px, py, ps.dir);
}

// isOutsideFov (Camera.h:154-164) for a world point; true = the cone test passes (camera may see it)
__device__ __forceinline__ bool insideCone(const DevCamera& c, double wx, double wy, double wz) {
  if (c.cosFov == -1) return true;
  const double vx = wx - c.pos[0], vy = wy - c.pos[1], vz = wz - c.pos[2];
  const double camz = c.rot[6] * vx + c.rot[7] * vy + c.rot[8] * vz;
  if (c.cosFov == 0) return !(camz >= 0);  // !isBehind
  const double dot = -camz;
  const double n2 = vx * vx + vy * vy + vz * vz;
  return !(dot * fabs(dot) <= c.cosFov * fabs(c.cosFov) * n2);
}

// The fields of the cone test alone, for kernels that pass them BY VALUE in their parameter block: the candidate loop
// of the dense sweep tests every source for every candidate with a warp-uniform source index, so reading the seven doubles
// from the constant bank (uniform loads / constant operands) instead of shared memory takes that traffic off the
// L1 / shared-memory pipe the sweep is limited by.
struct ConeCam {
  double pos[3];
  double back[3];  // rotation row 2 (= -forward)
  double cosFov;
  double pad;
};
__device__ __forceinline__ bool insideCone(const ConeCam& c, double wx, double wy, double wz) {
  if (c.cosFov == -1) return true;
  const double vx = wx - c.pos[0], vy = wy - c.pos[1], vz = wz - c.pos[2];
  const double camz = c.back[0] * vx + c.back[1] * vy + c.back[2] * vz;
  if (c.cosFov == 0) return !(camz >= 0);  // !isBehind
  const double dot = -camz;
  const double n2 = vx * vx + vy * vy + vz * vz;
  return !(dot * fabs(dot) <= c.cosFov * fabs(c.cosFov) * n2);
}

// Conservative fp32 version of the same test (-DDERP_CONE_F32; built, validated, measured 1.2 % SLOWER than the plain fp64
// test on B200 — the fp64 pipe is not the binding resource of the sweep — and therefore off by default):
// 1 = inside, 0 = outside, -1 = too close to call (the caller then runs insideCone in fp64, so the decision is ALWAYS the
// reference's).
// Error bound (u = 2^-24, |.|_1 the 1-norm, B = |w|_1 + |pos|_1 >= |v|_1 >= |v|_2): every component of
// v = fl(fl(w) - fl(pos)) is off by at most u(|w_i| + |pos_i| + |v_i|) <= 2uB; the forward axis is a unit vector
// rounded to fp32, so dot = f.v (three rounded operations) is off by at most |f|_2 |dv|_2 + 4u|v| < 8uB; n2 = v.v by at most
// 4|v|uB*sqrt(3) + 3u n2 < 10uB^2; hence lhs - rhs = dot|dot| - c2 n2 (|c2| <= 1, c2 rounded: +u) is off by less than
// 2B*8uB + (8uB)^2 + 10uB^2 + 3uB^2 < 30uB^2.  The margin used is 64uB^2 = 2^-18 B^2 (2^-20 B on dot for the hemisphere
// case, bound 8uB = 2^-21 B).  NaN / infinite inputs fail both comparisons and fall through to the fp64 test.
__device__ __forceinline__ int coneClass(const DevCamera& c, float wx, float wy, float wz, float wL1) {
  if (c.coneMode == 1) return 1;
  const float vx = wx - c.conePos[0], vy = wy - c.conePos[1], vz = wz - c.conePos[2];
  const float dot = __fmaf_rn(c.coneFwd[2], vz, __fmaf_rn(c.coneFwd[1], vy, c.coneFwd[0] * vx));
  const float B = wL1 + c.conePosL1;
  if (c.coneMode == 2) {  // inside iff !(camz >= 0) iff forward . v > 0
    const float m = 9.5367431640625e-07f * B;  // 2^-20 B
    return dot > m ? 1 : (dot < -m ? 0 : -1);
  }
  const float n2 = __fmaf_rn(vz, vz, __fmaf_rn(vy, vy, vx * vx));
  const float d = __fmaf_rn(-c.coneC2, n2, dot * fabsf(dot));  // inside iff !(dot|dot| <= c2 n2) iff d > 0
  const float m = 3.814697265625e-06f * (B * B);  // 2^-18 B^2
  return d > m ? 1 : (d < -m ? 0 : -1);
}

// pixel() + isOutsideSensor + de-normalisation + narrowing (Camera.h:121-128,180-190, DerpUtil.cpp:56-73,
// Derp.cpp:175) for a point that already passed the cone test.  Straight-line code (no early exit) so that
// the scheduler can interleave it with the fp32 SSD of the previous source.
struct SrcPoint {
  float x, y;
  bool ok;
};
__device__ __forceinline__ SrcPoint projectToSource(const DevCamera& c, double wx, double wy, double wz, int W, int H) {
  const double vx = wx - c.pos[0], vy = wy - c.pos[1], vz = wz - c.pos[2];
  const double camx = c.rot[0] * vx + c.rot[1] * vy + c.rot[2] * vz;
  const double camy = c.rot[3] * vx + c.rot[4] * vy + c.rot[5] * vz;
  const double camz = c.rot[6] * vx + c.rot[7] * vy + c.rot[8] * vz;
  double sx, sy;
  cameraToSensor(c, camx, camy, camz, &sx, &sy);
  double px = c.focal[0] * sx + c.principal[0];
  double py = c.focal[1] * sy + c.principal[1];
  SrcPoint o;
  o.ok = !(0 > px || px >= c.res[0] || 0 > py || py >= c.res[1]);
  px *= W;  // worldToSrcPoint: de-normalise (cameras are normalised, DerpUtil.cpp:67-71)
  py *= H;
  o.x = (float)px;
  o.y = (float)py;
  return o;
}

#ifdef DERP_EXPERIMENT_PROJ_F32
// TIMING EXPERIMENT ONLY (results of the lower-bound pass are not valid bounds in this build): the projection of the
// lower-bound pass in fp32, to measure what an fp32 projection with a position-error analysis could save at most.
__device__ __forceinline__ SrcPoint projectToSourceF32(const DevCamera& c, float wx, float wy, float wz, int W, int H) {
  const float vx = wx - c.conePos[0], vy = wy - c.conePos[1], vz = wz - c.conePos[2];
  const float r0 = (float)c.rot[0], r1 = (float)c.rot[1], r2 = (float)c.rot[2], r3 = (float)c.rot[3], r4 = (float)c.rot[4],
              r5 = (float)c.rot[5];
  const float camx = __fmaf_rn(r2, vz, __fmaf_rn(r1, vy, r0 * vx));
  const float camy = __fmaf_rn(r5, vz, __fmaf_rn(r4, vy, r3 * vx));
  const float camz = -__fmaf_rn(c.coneFwd[2], vz, __fmaf_rn(c.coneFwd[1], vy, c.coneFwd[0] * vx));
  const float xy2 = __fmaf_rn(camy, camy, camx * camx);
  const float inv = rsqrtf(xy2), xy = xy2 * inv;
  float r = atan2f(xy, -camz);
  const float dm = (float)c.distMax;
  r = r < dm ? r : dm;
  const float q = r * r;
  const float fac = __fmaf_rn(q, __fmaf_rn(q, __fmaf_rn(q, (float)c.dist[2], (float)c.dist[1]), (float)c.dist[0]), 1.0f);
  const float f = fac * r * inv;
  const float px = __fmaf_rn((float)c.focal[0], f * camx, (float)c.principal[0]);
  const float py = __fmaf_rn((float)c.focal[1], f * camy, (float)c.principal[1]);
  SrcPoint o;
  o.ok = !(0 > px || px >= 1.0f || 0 > py || py >= 1.0f);
  o.x = px * W;
  o.y = py * H;
  return o;
}
#endif

// Generic (border / inconsistent-rounding / invalid) path of one source: returns false if the source
// contributes no SSD (warp entry NaN).  Kept out of line: it runs for a few pixels per image.
template <class TX>
__device__ __noinline__ bool ssdSlowPath(const TX* __restrict__ srcColor, const TX* __restrict__ srcBiasImg,
                                         int W, int H, const float2* bg, const float2* rr, int rowPitch, int colPitch,
                                         float dBias0, float dBias1, float dBias2, float xDstSrc, float yDstSrc,
                                         float* ssdB, float* ssdU) {
  if (isnan(xDstSrc) || isnan(yDstSrc)) return false;
  const Texel sbias = sampleTexelTruncBiased(srcBiasImg, W, H, xDstSrc, yDstSrc);
  const float bias0 = dBias0 - sbias.b, bias1 = dBias1 - sbias.g, bias2 = dBias2 - sbias.r;
  float sB = 0.0f, sU = 0.0f;
  for (int dx = -1; dx <= 1; ++dx)
    for (int dy = -1; dy <= 1; ++dy) {
      const Texel cs = sampleTexelTruncBiased(srcColor, W, H, xDstSrc + (float)dx, yDstSrc + (float)dy);
      const float2 pbg = bg[(dy + 1) * rowPitch + (dx + 1) * colPitch];
      const float pr = rr[(dy + 1) * rowPitch + (dx + 1) * colPitch].x;
      const float d0 = pbg.x - cs.b, d1 = pbg.y - cs.g, d2 = pr - cs.r;
      const float u0 = d0 - bias0, u1 = d1 - bias1, u2 = d2 - bias2;
      sB += d0 * d0 + d1 * d1 + d2 * d2;
      sU += u0 * u0 + u1 * u1 + u2 * u2;
    }
  const float scaleFactor = 1.0f / (65535.0f * 65535.0f);
  *ssdB = sB * scaleFactor;
  *ssdU = sU * scaleFactor;
  return true;
}

// roundf(p) for both lanes of p when 0 <= p < 2^22, returned biased by 2^23 (so the integer index is a
// plain integer subtract of the bit patterns): floor(RZ(p + .5)) == floor(p + .5) == roundf(p) for p >= 0.
__device__ __forceinline__ f32x2 roundBiased2(f32x2 p, f32x2 half2, f32x2 b23) { return addrz2(addrz2(p, half2), b23); }


// ---- lower-bound pass of the filtered sweep (derp_refine.cuh) ------------------------------------------------------
// Per sample and channel the reference computes t = (ushort) fl32(bilerp) (CvUtil.h:83-120).  The cheap path computes
// a = fma-based separable bilerp of the SAME four texels with the SAME fp32 weights.  Bounds (texels in [0, 65535],
// weights in [0, 1], u = 2^-24):
//   |fl32(bilerp) - V| <= 0.0273   (V = the real-valued bilerp: three roundings per weight, one per product, three adds)
//   |a - V|            <= 0.0234   (three fused lerps on exact integer differences)
//   t in (fl32(bilerp) - 1, fl32(bilerp)]
// so with the midpoint m = a - 0.5:  |t - m| <= 0.5 + 0.0507, and after the two fp32 subtractions that form the
// difference against the destination texel (+0.5) and the bias: |e_d| <= 0.56 per term, |e_bias| <= 0.56.
// Over the 27 terms of one source: ||d|| >= ||d'|| - ||e_d||, ||e_d|| <= 0.56 sqrt(27) < 2.91 (reverse triangle
// inequality in l2), and for the bias-compensated differences ||e_u|| <= 2 * 2.91.
constexpr float kErrB = 2.91f;
constexpr float kErrU = 5.82f;

__device__ __forceinline__ float sqrtApprox(float x) {  // MUFU.SQRT, relative error < 2^-22: covered by the slack of kErr*
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// a + w (b - a), both lanes
__device__ __forceinline__ f32x2 lerp2(f32x2 a, f32x2 b, f32x2 w) { return fma2(w, sub2(b, a), a); }
__device__ __forceinline__ float lerp1(float a, float b, float w) { return __fmaf_rn(w, b - a, a); }

// Sums of squared differences (biased / bias-compensated) of one source over the 3x3 patch, cheap version: the 4x4
// texel block r0..r3 (4 texels each), the bias block at b1, per-sample weights W0..W2 = (xw, yw) of d = -1, 0, +1.
// The destination patch tile and ps.dBias carry +0.5 (loadDstTile / loadPixelState with addend 0.5).
template <int RP, int CP>
__device__ __forceinline__ void ssdApprox(const float4* r0, const float4* r1, const float4* r2, const float4* r3,
                                          const float4* b1, int W, const PixelState& ps, f32x2 W0, f32x2 W1, f32x2 W2,
                                          float* outB, float* outU) {
  const float xw[3] = {lo2(W0), lo2(W1), lo2(W2)};
  const float yw[3] = {hi2(W0), hi2(W1), hi2(W2)};
  // bias sample (centre sample's footprint): (B, G) packed, R through the (R, R-below) lanes
  f32x2 biasBG;
  float biasR;
  {
    const float4 q00 = __ldg(b1), q01 = __ldg(b1 + 1), q10 = __ldg(b1 + W), q11 = __ldg(b1 + W + 1);
    const f32x2 wx2 = pk(xw[1], xw[1]), wy2 = pk(yw[1], yw[1]);
    const f32x2 top = lerp2(pk(q00.x, q00.y), pk(q01.x, q01.y), wx2);
    const f32x2 bot = lerp2(pk(q10.x, q10.y), pk(q11.x, q11.y), wx2);
    biasBG = sub2(pk(ps.dBias[0], ps.dBias[1]), lerp2(top, bot, wy2));
    biasR = ps.dBias[2] - lerp1(lerp1(q00.z, q01.z, xw[1]), lerp1(q10.z, q11.z, xw[1]), yw[1]);
  }
  const f32x2 biasRR = pk(biasR, biasR);
  f32x2 accB = pk(0.f, 0.f), accU = pk(0.f, 0.f), accBR = pk(0.f, 0.f), accUR = pk(0.f, 0.f);
  float accBs = 0.f, accUs = 0.f;
  float4 A0 = __ldg(r0), A1 = __ldg(r1), A2 = __ldg(r2), A3 = __ldg(r3);
#pragma unroll
  for (int c = 0; c < 3; ++c) {  // dx = c - 1: texel columns c and c + 1
    const float4 B0 = __ldg(r0 + c + 1), B1 = __ldg(r1 + c + 1), B2 = __ldg(r2 + c + 1), B3 = __ldg(r3 + c + 1);
    const f32x2 wx2 = pk(xw[c], xw[c]);
    // x-lerps of the four texel rows: (B, G) lanes; R of rows (0,1) and (2,3) through the (z, w) lanes
    const f32x2 x0 = lerp2(pk(A0.x, A0.y), pk(B0.x, B0.y), wx2), x1 = lerp2(pk(A1.x, A1.y), pk(B1.x, B1.y), wx2);
    const f32x2 x2 = lerp2(pk(A2.x, A2.y), pk(B2.x, B2.y), wx2), x3 = lerp2(pk(A3.x, A3.y), pk(B3.x, B3.y), wx2);
    const f32x2 xr01 = lerp2(pk(A0.z, A0.w), pk(B0.z, B0.w), wx2), xr23 = lerp2(pk(A2.z, A2.w), pk(B2.z, B2.w), wx2);
    const float xr0 = lo2(xr01), xr1 = hi2(xr01), xr2 = lo2(xr23), xr3 = hi2(xr23);
    // y-lerps: sample rows dy = -1, 0, +1
    const f32x2 s0 = lerp2(x0, x1, pk(yw[0], yw[0])), s1 = lerp2(x1, x2, pk(yw[1], yw[1])), s2 = lerp2(x2, x3, pk(yw[2], yw[2]));
    const f32x2 sr01 = lerp2(pk(xr0, xr1), pk(xr1, xr2), pk(yw[0], yw[1]));
    const float sr2 = lerp1(xr2, xr3, yw[2]);
    // differences against the destination patch (+0.5) and the bias
    const f32x2 d0 = sub2(*reinterpret_cast<const f32x2*>(ps.bg + 0 * RP + c * CP), s0);
    const f32x2 d1 = sub2(*reinterpret_cast<const f32x2*>(ps.bg + 1 * RP + c * CP), s1);
    const f32x2 d2 = sub2(*reinterpret_cast<const f32x2*>(ps.bg + 2 * RP + c * CP), s2);
    const f32x2 dr01 = sub2(*reinterpret_cast<const f32x2*>(ps.rr + c * CP), sr01);  // (R(dy=-1), R(dy=0))
    const float dr2 = ps.rr[2 * RP + c * CP].x - sr2;
    const f32x2 u0 = sub2(d0, biasBG), u1 = sub2(d1, biasBG), u2 = sub2(d2, biasBG), ur01 = sub2(dr01, biasRR);
    const float ur2 = dr2 - biasR;
    accB = fma2(d0, d0, accB);
    accB = fma2(d1, d1, accB);
    accB = fma2(d2, d2, accB);
    accBR = fma2(dr01, dr01, accBR);
    accBs = __fmaf_rn(dr2, dr2, accBs);
    accU = fma2(u0, u0, accU);
    accU = fma2(u1, u1, accU);
    accU = fma2(u2, u2, accU);
    accUR = fma2(ur01, ur01, accUR);
    accUs = __fmaf_rn(ur2, ur2, accUs);
    A0 = B0;
    A1 = B1;
    A2 = B2;
    A3 = B3;
  }
  *outB = (lo2(accB) + hi2(accB)) + (lo2(accBR) + hi2(accBR)) + accBs;
  *outU = (lo2(accU) + hi2(accU)) + (lo2(accUR) + hi2(accUR)) + accUs;
}

// Lower bound of computeCost's result from the per-source slots (sqrt of the approximate biased sum, lower bound of the
// unbiased sum), both BEFORE the 1/65535^2 scale.  The reference keeps the `keep` = max(1, n - 2) sources with the
// smallest (biased, unbiased) pairs (Derp.cpp:204-215).  If the biased sums of the dropped and the kept sources are
// separated by more than the error of their difference (2 kErrB on the square roots), the kept set is known and the
// bound is the sum of its lower bounds; otherwise the sum of the `keep` smallest lower bounds, which bounds every
// possible kept set.  The factor 1 - 2^-16 covers the fp32 roundings of both computations (sums, scale, the two
// divisions and the multiplication of Derp.cpp:216-225: < 5e-6 relative together).
template <class V>
__device__ __forceinline__ float lowerBoundOfCost(const V& slots, int n, int keep, float conf) {
  float r1 = -1.f, r2 = -1.f, r3 = -1.f, l1 = 0.f, l2 = 0.f;  // three largest roots; lower bounds of the two largest
  float t1 = 0.f, t2 = 0.f;                                      // two largest lower bounds
  float sum = 0.f;
  for (int i = 0; i < n; ++i) {
    const PairVal p = slots.get(i);
    sum += p.b;
    if (p.a > r1) {
      r3 = r2;
      r2 = r1;
      l2 = l1;
      r1 = p.a;
      l1 = p.b;
    } else if (p.a > r2) {
      r3 = r2;
      r2 = p.a;
      l2 = p.b;
    } else if (p.a > r3) {
      r3 = p.a;
    }
    if (p.b > t1) {
      t2 = t1;
      t1 = p.b;
    } else if (p.b > t2) {
      t2 = p.b;
    }
  }
  float kept;
  if (n == 1) {
    kept = sum;
  } else if (n == 2) {  // keep 1 = the smaller biased sum
    kept = (r1 - r2 > 2.f * kErrB) ? sum - l1 : sum - t1;
  } else {              // drop the two largest biased sums
    kept = (r2 - r3 > 2.f * kErrB) ? sum - l1 - l2 : sum - t1 - t2;
  }
  kept = fmaxf(kept, 0.f);
  const float scaleFactor = 1.0f / (65535.0f * 65535.0f);
  const float k = (float)keep;
  return ((kept * scaleFactor) / k) * (1.0f / k) / conf * 0.9999847412109375f;  // 1 - 2^-16
}

// computeCost (Derp.cpp:104-226).  `cams` points to shared memory.  Returns the cost; confidence is
// ps.conf when the return value is not FLT_MAX, 0 otherwise.
//
// Structure (instruction-issue bound kernel, see profiles/README.md):
//   phase A  cone test of every source -> bitmask (short independent fp64 chains, unrolled);
//   phase B  for each set bit: the fp64 projection of the NEXT source is issued in the same basic block as
//            the fp32 SSD of the CURRENT one, so the long dependent fp64 chain (sqrt, atan2, divide) and the
//            two dependent gathers (warp entry -> 4x4 block) hide behind each other.
// SSD fast path: the 36 taps of the nine 2x2 footprints are one interior 4x4 block read column by column
// (= the reference's accumulation order, dx outer / dy inner).  Channels B,G ride the two lanes of the
// packed instructions; channel R rides them as (sample dy=-1, sample dy=0) thanks to the table's w lane
// holding R of the texel below; the dy=+1 R sample is scalar.  Weights are formed per sample exactly like
// the reference, so the result is bit-identical to the generic per-tap path.
// RP / CP: row and column pitch (in float2) of the thread's 3x3 destination patch in shared memory —
// (kTileW, 1) for the dense CTA tile, (3*256, 256) for the per-thread patches of the compacted kernels.
// TX: table texel type (float4 for the dense sweep, uint2 = 4 x u16 for the compacted kernels).
// LOWER = true: the LOWER-BOUND pass of the filtered sweep (derp_refine.cuh).  Visibility, projection, warp fetch and
// sample positions are the exact path's; only the SSD arithmetic is replaced by a cheap approximation with a proven
// error bound, and the return value is a number that is <= the exact cost (0 = "unknown", FLT_MAX = no source).
template <int RP, int CP, class TX = float4, bool LOWER = false>
__device__ __forceinline__ float evalCost(const CostView& v, const DevCamera* __restrict__ cams,
                                          const PixelState& ps, float disparity, unsigned* hits,
                                          const ConeCam* __restrict__ cone = nullptr) {
  const double depth = (double)(1.0f / disparity);
  const DevCamera& cd = cams[v.self];
  const double wx = cd.pos[0] + ps.dir[0] * depth;
  const double wy = cd.pos[1] + ps.dir[1] * depth;
  const double wz = cd.pos[2] + ps.dir[2] * depth;
  const int W = v.W, H = v.H;
  const size_t plane = (size_t)W * H;
  const float one = v.one, b23 = v.b23;
  const f32x2 one2 = pk(one, one), b232 = pk(b23, b23), half2 = pk(0.5f, 0.5f);

  unsigned mask = 0;
  {
    const float fx = (float)wx, fy = (float)wy, fz = (float)wz;
    const float wL1 = fabsf(fx) + fabsf(fy) + fabsf(fz);
#pragma unroll 4
    for (int s = 0; s < v.S; ++s) {
#if defined(DERP_EXPERIMENT_PROJ_F32)
      int cls = -1;
      if constexpr (LOWER) {
        cls = coneClass(cams[s], fx, fy, fz, wL1);
        cls = cls < 0 ? 1 : cls;
      }
#elif defined(DERP_CONE_F32)  // measured on B200 (2048^2 x 128): 89.1 ms per launch with the pre-test, 88.0 without: not kept
      const int cls = coneClass(cams[s], fx, fy, fz, wL1);
#else
      const int cls = -1;
#endif
      const bool in = cls < 0 ? (cone ? insideCone(cone[s], wx, wy, wz) : insideCone(cams[s], wx, wy, wz)) : (cls != 0);
      if (in) mask |= 1u << s;
    }
  }
  mask &= ~(1u << v.self);

  // (biased, unbiased) SSD of every contributing source: this thread's column of the [slot][thread] array in shared
  // memory, S - 1 slots (sized at launch), so no evaluation ever touches local memory.
  int n = 0;
  bool unknown = false;  // LOWER only: a source took the generic (border) path, no bound is formed
  auto pushPair = [&](float b, float u) {
    ps.sel[n * ps.selStride] = make_float2(b, u);
    ++n;
  };
  // The four warp-table taps of a projected point (getPixelBilinear on the Vec2f table, CvUtil.h:107-120) and its
  // bilinear weights.  A point inside the sensor has coordinates >= 0, so the RZ rounding applies; for a point
  // outside (!ok) the clamped, harmless fetch result is discarded by the caller.
  struct WarpTaps {
    float2 p00, p01, p10, p11;
    float xw, yw;
  };
  auto fetchWarp = [&](int s, const SrcPoint& sp) {
    const f32x2 P = pk(sp.x, sp.y);
    const f32x2 T = roundBiased2(P, half2, b232);
    const f32x2 Wt = add2(sub2(P, sub2(T, b232)), half2);  // (xw, yw) = p - round(p) + 0.5
    const int wxi = __float_as_int(lo2(T)) - 0x4B000000, wyi = __float_as_int(hi2(T)) - 0x4B000000;
    const float2* wt = v.projWarp + s * plane;
    const int x0 = clampIdx(wxi - 1, W - 1), x1 = clampIdx(wxi, W - 1);
    const int y0 = clampIdx(wyi - 1, H - 1), y1 = clampIdx(wyi, H - 1);
    WarpTaps t;
    t.p00 = __ldg(wt + (size_t)y0 * W + x0);
    t.p01 = __ldg(wt + (size_t)y0 * W + x1);
    t.p10 = __ldg(wt + (size_t)y1 * W + x0);
    t.p11 = __ldg(wt + (size_t)y1 * W + x1);
    t.xw = lo2(Wt);
    t.yw = hi2(Wt);
    return t;
  };
#ifdef DERP_EXPERIMENT_PROJ_F32
#define DERP_PROJECT(cam) (LOWER ? projectToSourceF32(cam, (float)wx, (float)wy, (float)wz, W, H) : projectToSource(cam, wx, wy, wz, W, H))
#else
#define DERP_PROJECT(cam) projectToSource(cam, wx, wy, wz, W, H)
#endif
  if (mask) {
    int s = __ffs(mask) - 1;
    mask &= mask - 1;
    SrcPoint cur = DERP_PROJECT(cams[s]);
    while (true) {
      const int sNext = mask ? __ffs(mask) - 1 : s;  // tail: harmless re-projection of the same source
      const bool more = mask != 0;
      mask &= mask - 1;
      // ---- current source: warp entry ---------------------------------------------------------------------
      float2 pd;
      {
        const WarpTaps tp = fetchWarp(s, cur);
        const float2 p00 = tp.p00, p01 = tp.p01, p10 = tp.p10, p11 = tp.p11;
        const float xw = tp.xw, yw = tp.yw;
        const float xm = 1 - xw, ym = 1 - yw;
        const f32x2 r = bilerp2(pk(p00.x, p00.y), pk(p01.x, p01.y), pk(p10.x, p10.y), pk(p11.x, p11.y),
                                pk(xm * ym, xm * ym), pk(xw * ym, xw * ym), pk(xm * yw, xm * yw), pk(xw * yw, xw * yw), one2);
        pd.x = lo2(r);
        pd.y = hi2(r);
      }
      const float xDstSrc = pd.x + 0.5f, yDstSrc = pd.y + 0.5f;
      // ---- footprint of the nine samples: (x,y) pairs for d = -1, 0, +1 ----------------------------------
      const f32x2 C = pk(xDstSrc, yDstSrc);
      const f32x2 P0 = add2(C, pk(-1.0f, -1.0f)), P2 = add2(C, pk(1.0f, 1.0f));  // xDstSrc + dx as float adds
      const f32x2 T0 = roundBiased2(P0, half2, b232), T1 = roundBiased2(C, half2, b232), T2 = roundBiased2(P2, half2, b232);
      const f32x2 W0 = add2(sub2(P0, sub2(T0, b232)), half2), W1 = add2(sub2(C, sub2(T1, b232)), half2),
                  W2 = add2(sub2(P2, sub2(T2, b232)), half2);
      const int xi0 = __float_as_int(lo2(T0)) - 0x4B000000, yi0 = __float_as_int(hi2(T0)) - 0x4B000000;
      const int xi1 = __float_as_int(lo2(T1)) - 0x4B000000, yi1 = __float_as_int(hi2(T1)) - 0x4B000000;
      const int xi2 = __float_as_int(lo2(T2)) - 0x4B000000, yi2 = __float_as_int(hi2(T2)) - 0x4B000000;
      const int X0 = xi0 - 1, Y0 = yi0 - 1;
      // the RZ rounding needs 0 <= p < 2^22; NaN fails every comparison -> slow path, which rejects it
      const bool fast = (xDstSrc >= 1.5f) & (yDstSrc >= 1.5f) & (xDstSrc < 4.0e6f) & (yDstSrc < 4.0e6f) &
          (xi1 == xi0 + 1) & (xi2 == xi0 + 2) & (yi1 == yi0 + 1) & (yi2 == yi0 + 2) & (X0 + 3 <= W - 1) & (Y0 + 3 <= H - 1);
      const TX* srcColor = TablesOf<TX>::color(v) + s * plane;
      const TX* srcBiasImg = TablesOf<TX>::bias(v) + s * plane;
      SrcPoint nxt;
      if (cur.ok) {
        if (fast) {
          // ---- main block: 20 gathers + fp32 SSD of source s, fp64 projection of source sNext --------------
          const size_t off = (size_t)Y0 * W + X0;
          const TX* r0 = srcColor + off;
          const TX* r1 = r0 + W;
          const TX* r2 = r1 + W;
          const TX* r3 = r2 + W;
          const TX* b1 = srcBiasImg + off + W + 1;  // bias sample = centre sample's 2x2 footprint
          if constexpr (LOWER) {
            nxt = DERP_PROJECT(cams[sNext]);
            float sB, sU;
            ssdApprox<RP, CP>(r0, r1, r2, r3, b1, W, ps, W0, W1, W2, &sB, &sU);
            // slot = (sqrt of the biased sum, lower bound of the unbiased sum), see lowerBoundOfCost
            const float rB = sqrtApprox(sB), rU = sqrtApprox(sU);
            const float ul = fmaxf(rU - kErrU, 0.0f);
            pushPair(rB, ul * ul);
          } else {
#ifdef DERP_EXPERIMENT_NOBIAS  // measurement-only variant (breaks parity): upper bound of not reading the bias table
          const float4 q00 = make_float4(1.f, 2.f, 3.f, 0.f), q01 = q00, q10 = q00, q11 = q00;
          (void)b1;
#else
          const float4 q00 = ldTexel(b1), q01 = ldTexel(b1 + 1), q10 = ldTexel(b1 + W), q11 = ldTexel(b1 + W + 1);
#endif
          float4 colA[4], colB[4];
          colA[0] = ldTexel(r0);
          colA[1] = ldTexel(r1);
          colA[2] = ldTexel(r2);
          colA[3] = ldTexel(r3);
          nxt = projectToSource(cams[sNext], wx, wy, wz, W, H);
          // y weights: per sample row r (dy = r-1): yw, 1-yw; rows 0,1 also as a packed pair for channel R
          const float yw0 = hi2(W0), yw1 = hi2(W1), yw2 = hi2(W2);
          const float ym0 = 1 - yw0, ym1 = 1 - yw1, ym2 = 1 - yw2;
          const f32x2 ywP = pk(yw0, yw1), ymP = pk(ym0, ym1);
          const float xwv[3] = {lo2(W0), lo2(W1), lo2(W2)};
          // bias = float(dstBias) - float(srcBias): both carry +2^23, the difference is exact
          f32x2 biasBG;
          float biasR;
          {
            const float xwc = xwv[1], xwm = 1 - xwc;
            const float w00 = xwm * ym1, w01 = xwc * ym1, w10 = xwm * yw1, w11 = xwc * yw1;
            const f32x2 sbg = addrz2(bilerp2(pk(q00.x, q00.y), pk(q01.x, q01.y), pk(q10.x, q10.y), pk(q11.x, q11.y),
                                            pk(w00, w00), pk(w01, w01), pk(w10, w10), pk(w11, w11), one2), b232);
            biasBG = sub2(pk(ps.dBias[0], ps.dBias[1]), sbg);
            biasR = ps.dBias[2] - truncBiased(bilerp1(q00.z, q01.z, q10.z, q11.z, w00, w01, w10, w11));
          }
          const f32x2 biasRR = pk(biasR, biasR);
          float sB = 0.0f, sU = 0.0f;
#pragma unroll
          for (int c = 0; c < 3; ++c) {  // dx = c-1: texel columns c and c+1
            colB[0] = ldTexel(r0 + c + 1);
            colB[1] = ldTexel(r1 + c + 1);
            colB[2] = ldTexel(r2 + c + 1);
            colB[3] = ldTexel(r3 + c + 1);
            const float xwc = xwv[c], xwm = 1 - xwc;
            const f32x2 xwc2 = pk(xwc, xwc), xwm2 = pk(xwm, xwm);
            // weights of samples dy=-1,0 as pairs (lane = sample), dy=+1 scalar: w00=(1-xw)(1-yw) ...
            const f32x2 w00P = mul2(ymP, xwm2), w01P = mul2(ymP, xwc2), w10P = mul2(ywP, xwm2), w11P = mul2(ywP, xwc2);
            const float w00s = xwm * ym2, w01s = xwc * ym2, w10s = xwm * yw2, w11s = xwc * yw2;
            // channel R of samples dy=-1 and dy=0 in one go: the texels' (z,w) lanes are (R(row), R(row+1))
            const f32x2 sR01 = addrz2(bilerp2(pk(colA[0].z, colA[0].w), pk(colB[0].z, colB[0].w), pk(colA[1].z, colA[1].w),
                                              pk(colB[1].z, colB[1].w), w00P, w01P, w10P, w11P, one2), b232);
            const float sR2 = truncBiased(bilerp1(colA[2].z, colB[2].z, colA[3].z, colB[3].z, w00s, w01s, w10s, w11s));
            const f32x2 pR01 = *reinterpret_cast<const f32x2*>(ps.rr + c * CP);
            const float pR2 = ps.rr[2 * RP + c * CP].x;
            const f32x2 dR01 = sub2(pR01, sR01);
            const f32x2 uR01 = sub2(dR01, biasRR);
            const f32x2 ddR01 = mul2(dR01, dR01), uuR01 = mul2(uR01, uR01);
            const float dR2 = pR2 - sR2, uR2 = dR2 - biasR;
            const float ddR[3] = {lo2(ddR01), hi2(ddR01), dR2 * dR2};
            const float uuR[3] = {lo2(uuR01), hi2(uuR01), uR2 * uR2};
#pragma unroll
            for (int r = 0; r < 3; ++r) {  // dy = r-1: texel rows r and r+1; channels (B,G) on the two lanes
              const float w00 = r == 0 ? lo2(w00P) : (r == 1 ? hi2(w00P) : w00s);
              const float w01 = r == 0 ? lo2(w01P) : (r == 1 ? hi2(w01P) : w01s);
              const float w10 = r == 0 ? lo2(w10P) : (r == 1 ? hi2(w10P) : w10s);
              const float w11 = r == 0 ? lo2(w11P) : (r == 1 ? hi2(w11P) : w11s);
              const f32x2 sBG = addrz2(bilerp2(pk(colA[r].x, colA[r].y), pk(colB[r].x, colB[r].y), pk(colA[r + 1].x, colA[r + 1].y),
                                               pk(colB[r + 1].x, colB[r + 1].y), pk(w00, w00), pk(w01, w01), pk(w10, w10),
                                               pk(w11, w11), one2), b232);
              const f32x2 pBG = *reinterpret_cast<const f32x2*>(ps.bg + r * RP + c * CP);
              const f32x2 dBG = sub2(pBG, sBG);
              const f32x2 uBG = sub2(dBG, biasBG);
              const f32x2 dd = mul2(dBG, dBG), uu = mul2(uBG, uBG);
              // cv::Matx::dot order (d0*d0 + d1*d1) + d2*d2, then ssd += s (DerpUtil.cpp:150-151)
              const float s1 = addp(ddR[r], addp(hi2(dd), lo2(dd), one), one);
              const float s2 = addp(uuR[r], addp(hi2(uu), lo2(uu), one), one);
              sB += s1;
              sU += s2;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) colA[j] = colB[j];
          }
          const float scaleFactor = 1.0f / (65535.0f * 65535.0f);
          pushPair(sB * scaleFactor, sU * scaleFactor);
          }
        } else {
          nxt = projectToSource(cams[sNext], wx, wy, wz, W, H);
          if constexpr (LOWER) {
            if (!(isnan(xDstSrc) || isnan(yDstSrc))) {  // the source contributes, through the generic path
              unknown = true;
              ++n;
            }
          } else {
            float slowB, slowU;
            if (ssdSlowPath(srcColor, srcBiasImg, W, H, ps.bg, ps.rr, RP, CP, ps.dBias[0], ps.dBias[1], ps.dBias[2], xDstSrc,
                            yDstSrc, &slowB, &slowU))
              pushPair(slowB, slowU);
          }
        }
      } else {
        nxt = projectToSource(cams[sNext], wx, wy, wz, W, H);
      }
      if (!more) break;
      cur = nxt;
      s = sNext;
    }
  }
  *hits += n;
  if (n < 1) return FLT_MAX;  // kMinOverlappingCams - 1
  const int keep = n - 2 > 1 ? n - 2 : 1;
  if constexpr (LOWER) {
    if (unknown) return 0.0f;
    return lowerBoundOfCost(SmemPairs{ps.sel, ps.selStride}, n, keep, ps.conf);
  }
  float cost;
  if (n <= 3) {
    // keep == 1: nth_element(v, v+1, v+n) on <= 3 elements is an insertion sort, v[0] = the smallest pair
    float2 m = ps.sel[0];
    for (int i = 1; i < n; ++i) {
      const float2 t = ps.sel[i * ps.selStride];
      if (pairLess(t.x, t.y, m.x, m.y)) m = t;
    }
    cost = 0.0f + m.y;
  } else if (n <= kSelSlots) {
#ifdef DERP_SELECT_TABLE  // host-validated on every permutation (tests/test_host_units.py), GPU-validated by the parity suite
    static_assert(kSelSlots <= kSelTabMaxN, "the table path covers every evaluation that fits the shared-memory slots");
    // the 15-compare instance when no lane of the warp that got here holds more than 6 pairs
    const bool small = __reduce_max_sync(__activemask(), (unsigned)n) <= 6u;
    const bool done = small ? robustSumTable<6>(SmemPairs{ps.sel, ps.selStride}, n, keep, v.selTab, &cost)
                            : robustSumTable<8>(SmemPairs{ps.sel, ps.selStride}, n, keep, v.selTab, &cost);
    if (!done)
#endif
      cost = robustSum(SmemPairs{ps.sel, ps.selStride}, n, keep);
  } else {  // more than kSelTabMaxN sources: the general algorithm on the shared-memory slots
    cost = robustSum(SmemPairs{ps.sel, ps.selStride}, n, keep);
  }
  cost /= (float)keep;
  const float trustCoef = 1.0f / (float)keep;
  return cost * trustCoef / ps.conf;
}

}  // namespace derp
