// ORACLE — TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library, and only as the checker / the timed
// CPU baseline — never as the product path.
//
// CPU restatement of facebook360_dep's depth-estimation hot path (source/depth_estimation +
// the source/util functions it calls) behind the same C ABI as the CUDA library
// (include/derp_b200.h), so that a parity test is "same calls, two libraries, compare buffers".
// Arithmetic follows the reference exactly: fp64 camera, fp32 non-FMA cost in the reference's
// operation order (build: g++ -O3 -funroll-loops -ffp-contract=off, no -march, mirroring
// CMakeLists.txt:39-40), integer-truncating bilinear sampler, libstdc++ std::nth_element /
// std::minstd_rand0 / std::uniform_real_distribution called literally.
//
// PARITY PIN STATUS:
//   camera model            pinned — reference gtest known answers + reference numpy port vectors
//   OpenCV primitives       pinned to cv2 4.13.0 outputs (the reference pins no OpenCV version)
//   computeCost / brute force / random proposals / ping-pong / mismatch / bilateral / median /
//   temporal / upsample     PARITY UNPINNED by the reference: it ships no golden vectors for these
//                           (its only test is MSSIM 90%±5% on an S3 dataset) and cannot be compiled
//                           here (OpenCV C++, Eigen, Boost, gflags, glog, folly all absent).
//                           This file is the pin; every function cites the lines it restates.
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <random>
#include <string>
#include <thread>
#include <tuple>
#include <utility>
#include <vector>

#include "../include/derp_b200.h"
#include "camera.h"
#include "cvprims.h"

namespace oracle {

static thread_local std::string g_err;
static int g_threads = -1;

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

static int threadCount() {  // ThreadPool.h:19-32
  if (g_threads < 0) return std::max<int>(1, (int)std::thread::hardware_concurrency());
  return g_threads;
}

// Row-range parallel loop.  The reference spawns one std::thread per row / slice and joins in
// batches (ThreadPool.h:36-52); results are identical because tasks write disjoint rows.
static void parallelFor(int begin, int end, const std::function<void(int, int)>& fn) {
  const int n = end - begin;
  const int t = std::min(threadCount(), n);
  if (t <= 1) {
    if (n > 0) fn(begin, end);
    return;
  }
  std::atomic<int> next(begin);
  const int chunk = std::max(1, n / (t * 8));
  std::vector<std::thread> ths;
  for (int i = 0; i < t; ++i)
    ths.emplace_back([&] {
      for (;;) {
        const int b = next.fetch_add(chunk);
        if (b >= end) break;
        fn(b, std::min(end, b + chunk));
      }
    });
  for (auto& th : ths) th.join();
}

static inline int clampi(int x, int a, int b) { return x < a ? a : x > b ? b : x; }  // MathUtil.h:35-38

struct Vec2f {
  float x, y;
};

// ---- cv_util::getPixelBilinear (CvUtil.h:78-120) ---------------------------------------------
template <typename T>
static inline T bilerpScalar(T p00, T p01, T p10, T p11, float xw, float yw) {
  // the return statement converts float -> T: truncation for T = uint16_t (CvUtil.h:83-86)
  return (T)((1 - xw) * (1 - yw) * p00 + xw * (1 - yw) * p01 + (1 - xw) * yw * p10 + xw * yw * p11);
}

static inline Vec2f bilinearVec2f(const Vec2f* img, int w, int h, float x, float y) {
  const float xf = roundf(x), yf = roundf(y);
  const int xi = (int)xf, yi = (int)yf;
  const Vec2f& p00 = img[(size_t)clampi(yi - 1, 0, h - 1) * w + clampi(xi - 1, 0, w - 1)];
  const Vec2f& p01 = img[(size_t)clampi(yi - 1, 0, h - 1) * w + clampi(xi, 0, w - 1)];
  const Vec2f& p10 = img[(size_t)clampi(yi, 0, h - 1) * w + clampi(xi - 1, 0, w - 1)];
  const Vec2f& p11 = img[(size_t)clampi(yi, 0, h - 1) * w + clampi(xi, 0, w - 1)];
  const float xw = x - xf + 0.5f, yw = y - yf + 0.5f;
  Vec2f r;
  r.x = bilerpScalar<float>(p00.x, p01.x, p10.x, p11.x, xw, yw);
  r.y = bilerpScalar<float>(p00.y, p01.y, p10.y, p11.y, xw, yw);
  return r;
}

static inline float bilinearF32(const float* img, int w, int h, float x, float y) {
  const float xf = roundf(x), yf = roundf(y);
  const int xi = (int)xf, yi = (int)yf;
  const float p00 = img[(size_t)clampi(yi - 1, 0, h - 1) * w + clampi(xi - 1, 0, w - 1)];
  const float p01 = img[(size_t)clampi(yi - 1, 0, h - 1) * w + clampi(xi, 0, w - 1)];
  const float p10 = img[(size_t)clampi(yi, 0, h - 1) * w + clampi(xi - 1, 0, w - 1)];
  const float p11 = img[(size_t)clampi(yi, 0, h - 1) * w + clampi(xi, 0, w - 1)];
  return bilerpScalar<float>(p00, p01, p10, p11, x - xf + 0.5f, y - yf + 0.5f);
}

// Vec3w: per channel bilerp<ushort> -> truncated, returned as float (then cast back to Vec3w)
static inline void bilinearU16C3(const uint16_t* img, int w, int h, float x, float y, float* out) {
  const float xf = roundf(x), yf = roundf(y);
  const int xi = (int)xf, yi = (int)yf;
  const uint16_t* p00 = img + ((size_t)clampi(yi - 1, 0, h - 1) * w + clampi(xi - 1, 0, w - 1)) * 3;
  const uint16_t* p01 = img + ((size_t)clampi(yi - 1, 0, h - 1) * w + clampi(xi, 0, w - 1)) * 3;
  const uint16_t* p10 = img + ((size_t)clampi(yi, 0, h - 1) * w + clampi(xi - 1, 0, w - 1)) * 3;
  const uint16_t* p11 = img + ((size_t)clampi(yi, 0, h - 1) * w + clampi(xi, 0, w - 1)) * 3;
  const float xw = x - xf + 0.5f, yw = y - yf + 0.5f;
  for (int c = 0; c < 3; ++c)
    out[c] = (float)bilerpScalar<uint16_t>(p00[c], p01[c], p10[c], p11[c], xw, yw);
}

// ---- state -----------------------------------------------------------------------------------
struct Ctx {
  std::vector<Camera> camsFull;  // as loaded
  std::vector<Camera> cams;      // normalised (Camera::normalizeRig)
  std::vector<int> dst2src;
  int S = 0, Sd = 0;
  // level
  bool levelOpen = false;
  DerpLevelParams lp{};
  int W = 0, H = 0;
  float varNoiseFloor = 0;
  std::vector<std::vector<uint16_t>> color;    // [S] HxWx3
  std::vector<std::vector<float>> variance;    // [S]
  std::vector<std::vector<uint8_t>> fgMask;    // [S] (all-pass unless use_foreground_masks)
  std::vector<std::vector<float>> bgDisp;      // [Sd] (empty unless use_foreground_masks)
  std::vector<std::vector<uint8_t>> fovMask;   // [Sd]
  std::vector<std::vector<float>> disp, cost, conf;
  std::vector<std::vector<float>> kept;  // derp_level_keep
  int keptW = 0, keptH = 0;  // [Sd]
  std::vector<std::vector<float>> gathered;          // [S] all-camera disparities of a sharded mismatch stage
  bool haveGathered = false;
  std::vector<std::vector<uint8_t>> mismatch;  // [Sd]
  bool haveColors = false;
  // tables of the current destination
  int projDst = -1;
  std::vector<std::vector<Vec2f>> projWarp;       // [S]
  std::vector<std::vector<uint16_t>> projColor;   // [S]
  std::vector<std::vector<uint16_t>> projBias;    // [S]
  // counters
  std::atomic<uint64_t> costEvals{0}, srcHits{0};
};

// ---- DerpUtil.cpp:38-73 ------------------------------------------------------------------------
static inline void dstToWorldPoint(const Camera& camDst, int x, int y, float disparity, int dstW,
                                   int dstH, double* out) {
  double p[2] = {(x + 0.5) / dstW, (y + 0.5) / dstH};
  if (!camDst.isNormalized()) {
    p[0] *= camDst.resolution[0];
    p[1] *= camDst.resolution[1];
  }
  camDst.rig(p, (double)(1.0f / disparity), out);
}

static inline bool worldToSrcPoint(double* pSrc, const double* pWorld, const Camera& camSrc, int srcW,
                                   int srcH) {
  if (!camSrc.sees(pWorld, pSrc)) return false;
  if (camSrc.isNormalized()) {
    pSrc[0] *= srcW;
    pSrc[1] *= srcH;
  }
  return true;
}

// ---- computeSSD (DerpUtil.cpp:126-162) ---------------------------------------------------------
static inline std::pair<float, float> computeSSD(const uint16_t* dstColor, int W, int H, int x, int y,
                                                 const uint16_t* dstBias, const uint16_t* dstSrcColor,
                                                 float xDstSrc, float yDstSrc,
                                                 const float* dstSrcBias, int radius) {
  float bias[3];
  for (int c = 0; c < 3; ++c) bias[c] = (float)dstBias[c] - dstSrcBias[c];
  std::pair<float, float> ssd = {0.0f, 0.0f};
  for (int dx = -radius; dx <= radius; ++dx) {
    for (int dy = -radius; dy <= radius; ++dy) {
      const uint16_t* cD = dstColor + ((size_t)(y + dy) * W + (x + dx)) * 3;
      float cSrc[3];
      bilinearU16C3(dstSrcColor, W, H, xDstSrc + dx, yDstSrc + dy, cSrc);
      float diffBias[3], diffNoBias[3];
      for (int c = 0; c < 3; ++c) {
        diffBias[c] = (float)cD[c] - cSrc[c];
        diffNoBias[c] = diffBias[c] - bias[c];
      }
      float s1 = 0, s2 = 0;  // cv::Matx::dot: s = 0; s += a[i]*b[i]
      for (int c = 0; c < 3; ++c) s1 += diffBias[c] * diffBias[c];
      for (int c = 0; c < 3; ++c) s2 += diffNoBias[c] * diffNoBias[c];
      ssd.first += s1;
      ssd.second += s2;
    }
  }
  const float maxDepth = 65535.0f;
  const float scaleFactor = 1.0f / (maxDepth * maxDepth);
  ssd.first *= scaleFactor;
  ssd.second *= scaleFactor;
  return ssd;
}

static const float kMinVar = 1.0f / 12.0f / 65025.0f;  // DerpUtil.h:32
static const int kSearchWindowRadius = 1;              // Derp.h:26
static const int kMinOverlappingCams = 2;              // Derp.h:28

// ---- computeCost (Derp.cpp:104-226) ------------------------------------------------------------
static inline void computeCost(Ctx& c, int dstIdx, float disparity, int x, int y, float* outCost,
                               float* outConf, uint64_t* hits) {
  const int W = c.W, H = c.H;
  const int self = c.dst2src[dstIdx];
  const uint16_t* dstColor = c.projColor[self].data();
  const Camera& camDst = c.cams[self];
  double pWorld[3];
  dstToWorldPoint(camDst, x, y, disparity, W, H, pWorld);

  std::pair<float, float> SSDs[64];
  int ssdCount = 0;
  const uint16_t* dstColorBias = c.projBias[self].data();
  for (int srcIdx = 0; srcIdx < c.S; ++srcIdx) {
    if (srcIdx == self) continue;
    const Camera& camSrc = c.cams[srcIdx];
    double pSrc[2];
    if (!worldToSrcPoint(pSrc, pWorld, camSrc, W, H)) continue;
    const Vec2f pDstSrc = bilinearVec2f(c.projWarp[srcIdx].data(), W, H, (float)pSrc[0], (float)pSrc[1]);
    const float xDstSrc = pDstSrc.x + 0.5;  // float + double -> rounded once to float
    const float yDstSrc = pDstSrc.y + 0.5;
    if (std::isnan(xDstSrc) || std::isnan(yDstSrc)) continue;
    float dstSrcBias[3];
    bilinearU16C3(c.projBias[srcIdx].data(), W, H, xDstSrc, yDstSrc, dstSrcBias);
    const uint16_t* dstBias = dstColorBias + ((size_t)y * W + x) * 3;
    SSDs[ssdCount] = computeSSD(dstColor, W, H, x, y, dstBias, c.projColor[srcIdx].data(), xDstSrc,
                                yDstSrc, dstSrcBias, kSearchWindowRadius);
    ++ssdCount;
  }
  *hits += ssdCount;
  int keep = kMinOverlappingCams - 1;
  if (ssdCount < keep) {
    *outCost = FLT_MAX;
    *outConf = 0.0f;
    return;
  }
  keep = std::max<int>(keep, ssdCount - 2);
  std::nth_element(SSDs, SSDs + keep, SSDs + ssdCount);
  float cost = 0;
  for (int i = 0; i < keep; ++i) cost += SSDs[i].second;
  cost /= keep;
  const float trustCoef = 1.0f / keep;
  const float dstVariance = c.variance[self][(size_t)y * W + x];
  const float confidence = std::max(dstVariance, kMinVar);
  *outCost = cost * trustCoef / confidence;
  *outConf = confidence;
}

// ---- generateFovMasks (DerpUtil.cpp:239-276) ---------------------------------------------------
static void generateFovMask(const Camera& cam, int W, int H, uint8_t* mask) {
  parallelFor(0, H, [&](int y0, int y1) {
    for (int y = y0; y < y1; ++y)
      for (int x = 0; x < W; ++x) {
        double p[2] = {x + 0.5, y + 0.5};
        if (cam.isNormalized()) {
          p[0] = p[0] / W;
          p[1] = p[1] / H;
        }
        mask[(size_t)y * W + x] = !cam.isOutsideImageCircle(p);
      }
  });
}

// ---- computeWarpDstToSrc (ImageUtil.cpp:142-167) -----------------------------------------------
// `dst` and `src` are cameras rescaled to pixel resolution; map has dst's size.
static void computeWarpDstToSrc(const Camera& dst, const Camera& src, bool sameId, Vec2f* warp) {
  const int W = (int)dst.resolution[0], H = (int)dst.resolution[1];
  const float nan = std::numeric_limits<float>::quiet_NaN();
  parallelFor(0, H, [&](int y0, int y1) {
    for (int y = y0; y < y1; ++y)
      for (int x = 0; x < W; ++x) {
        Vec2f& o = warp[(size_t)y * W + x];
        o.x = nan;
        o.y = nan;
        if (sameId) continue;
        const double dstPixel[2] = {x + 0.5, y + 0.5};
        if (dst.isOutsideImageCircle(dstPixel)) continue;
        double rig[3];
        dst.rigNearInfinity(dstPixel, rig);
        double srcPixel[2];
        if (!src.sees(rig, srcPixel)) continue;
        o.x = (float)(srcPixel[0] - 0.5f);
        o.y = (float)(srcPixel[1] - 0.5f);
      }
  });
}

}  // namespace oracle

using namespace oracle;

struct DerpCtx {
  Ctx c;
};

extern "C" {

const char* derp_backend(void) { return "oracle-cpu"; }
const char* derp_last_error(void) { return g_err.c_str(); }
int derp_set_threads(int threads) {
  g_threads = threads;
  return DERP_OK;
}

int derp_create(const DerpCameraDesc* cams, int num_cams, const int32_t* dst_to_src, int num_dsts,
                int device, DerpCtx** out) {
  (void)device;
  if (!cams || !dst_to_src || !out || num_cams <= 0 || num_dsts <= 0 || num_cams > 64)
    return fail(DERP_EINVAL, "derp_create: bad arguments");
  std::unique_ptr<DerpCtx> h(new DerpCtx);
  Ctx& c = h->c;
  c.S = num_cams;
  c.Sd = num_dsts;
  c.camsFull.resize(num_cams);
  for (int i = 0; i < num_cams; ++i)
    if (!c.camsFull[i].init(cams[i])) return fail(DERP_EINVAL, "derp_create: invalid camera " + std::to_string(i));
  // PyramidLevel::checkParams (PyramidLevel.h:169-184): all cameras share one resolution
  for (int i = 1; i < num_cams; ++i)
    if (c.camsFull[i].resolution[0] != c.camsFull[0].resolution[0] ||
        c.camsFull[i].resolution[1] != c.camsFull[0].resolution[1])
      return fail(DERP_EINVAL, "derp_create: cameras must share one resolution");
  c.cams = c.camsFull;
  for (auto& cam : c.cams)
    if (!cam.isNormalized()) cam.normalize();
  c.dst2src.assign(dst_to_src, dst_to_src + num_dsts);
  for (int d : c.dst2src)
    if (d < 0 || d >= num_cams) return fail(DERP_EINVAL, "derp_create: dst_to_src out of range");
  *out = h.release();
  return DERP_OK;
}

void derp_destroy(DerpCtx* ctx) { delete ctx; }
int derp_set_stream(DerpCtx*, void*) { return DERP_OK; }
int derp_sync(DerpCtx*) { return DERP_OK; }
int derp_profile(DerpCtx*, int) { return DERP_OK; }
int derp_set_sweep_mode(DerpCtx*, int) { return DERP_OK; }
int derp_get_profile_ping_pong(DerpCtx*, double* ms, uint64_t* n, uint64_t* e, uint64_t* h) {
  if (ms) *ms = 0;
  if (n) *n = 0;
  if (e) *e = 0;
  if (h) *h = 0;
  return DERP_OK;
}
int derp_get_sweep_stats(DerpCtx*, uint64_t* a, uint64_t* b) {
  if (a) *a = 0;
  if (b) *b = 0;
  return DERP_OK;
}
int derp_get_profile(DerpCtx*, double* ms, uint64_t* n) {
  if (ms) *ms = 0;
  if (n) *n = 0;
  return DERP_OK;
}
int derp_get_launch_count(DerpCtx*, uint64_t* out) {
  if (out) *out = 0;  // the oracle launches no GPU kernels
  return DERP_OK;
}

int derp_level_begin(DerpCtx* ctx, const DerpLevelParams* p) {
  if (!ctx || !p || p->width < 3 || p->height < 3 || p->num_levels <= 0 || p->full_height <= 0)
    return fail(DERP_EINVAL, "derp_level_begin: bad arguments");
  Ctx& c = ctx->c;
  c.lp = *p;
  c.W = p->width;
  c.H = p->height;
  const size_t n = (size_t)c.W * c.H;
  // PyramidLevel::computeVariances (PyramidLevel.h:232-236) — width / heightFullSize, as written
  const float scale = float(c.W) / p->full_height;
  const float scaleVar = scale * scale;
  c.varNoiseFloor = std::max(p->var_noise_floor * scaleVar, kMinVar);
  c.color.assign(c.S, {});
  c.variance.assign(c.S, {});
  c.fgMask.assign(c.S, std::vector<uint8_t>(n, 1));
  c.bgDisp.assign(c.Sd, {});
  c.fovMask.assign(c.Sd, std::vector<uint8_t>(n));
  c.disp.assign(c.Sd, std::vector<float>(n, 0.f));
  c.cost.assign(c.Sd, std::vector<float>(n, 0.f));
  c.conf.assign(c.Sd, std::vector<float>(n, 0.f));
  c.mismatch.assign(c.Sd, std::vector<uint8_t>(n, 0));
  for (int d = 0; d < c.Sd; ++d) generateFovMask(c.cams[c.dst2src[d]], c.W, c.H, c.fovMask[d].data());
  c.projDst = -1;
  c.projWarp.assign(c.S, {});
  c.projColor.assign(c.S, {});
  c.projBias.assign(c.S, {});
  c.haveColors = false;
  c.levelOpen = true;
  return DERP_OK;
}

int derp_set_colors(DerpCtx* ctx, const uint16_t* const* colors) {
  if (!ctx || !colors) return fail(DERP_EINVAL, "derp_set_colors: bad arguments");
  Ctx& c = ctx->c;
  if (!c.levelOpen) return fail(DERP_ESTATE, "derp_set_colors: no level");
  const size_t n = (size_t)c.W * c.H;
  for (int s = 0; s < c.S; ++s) {
    if (!colors[s]) return fail(DERP_EINVAL, "derp_set_colors: null image");
    c.color[s].assign(colors[s], colors[s] + n * 3);
    c.variance[s].resize(n);
    const uint16_t* src = c.color[s].data();
    float* var = c.variance[s].data();
    parallelFor(0, c.H, [&](int y0, int y1) { imageVarianceU16C3(src, c.W, c.H, var, y0, y1); });
  }
  c.haveColors = true;
  c.projDst = -1;
  return DERP_OK;
}

int derp_set_foreground_masks(DerpCtx* ctx, const uint8_t* const* masks) {
  if (!ctx || !masks) return fail(DERP_EINVAL, "derp_set_foreground_masks: bad arguments");
  Ctx& c = ctx->c;
  if (!c.levelOpen) return fail(DERP_ESTATE, "no level");
  const size_t n = (size_t)c.W * c.H;
  for (int s = 0; s < c.S; ++s) {
    c.fgMask[s].resize(n);
    for (size_t i = 0; i < n; ++i) c.fgMask[s][i] = masks[s][i] ? 1 : 0;
  }
  return DERP_OK;
}

int derp_set_background_disparity(DerpCtx* ctx, const float* const* background) {
  if (!ctx || !background) return fail(DERP_EINVAL, "derp_set_background_disparity: bad arguments");
  Ctx& c = ctx->c;
  if (!c.levelOpen) return fail(DERP_ESTATE, "no level");
  const size_t n = (size_t)c.W * c.H;
  for (int d = 0; d < c.Sd; ++d) c.bgDisp[d].assign(background[d], background[d] + n);
  return DERP_OK;
}

static int checkDst(Ctx& c, int dst, const char* who, bool needProj) {
  if (!c.levelOpen) return fail(DERP_ESTATE, std::string(who) + ": no level");
  if (dst < 0 || dst >= c.Sd) return fail(DERP_EINVAL, std::string(who) + ": dst out of range");
  if (needProj && c.projDst != dst)
    return fail(DERP_ESTATE, std::string(who) + ": derp_reproject(dst) must precede this stage");
  return DERP_OK;
}

// precomputeProjections + reprojectColors for one dst (Derp.cpp:955-1003)
int derp_reproject(DerpCtx* ctx, int dst) {
  if (!ctx) return fail(DERP_EINVAL, "null ctx");
  Ctx& c = ctx->c;
  int rc = checkDst(c, dst, "derp_reproject", false);
  if (rc) return rc;
  if (!c.haveColors) return fail(DERP_ESTATE, "derp_reproject: colours not set");
  const int W = c.W, H = c.H;
  const size_t n = (size_t)W * H;
  const int self = c.dst2src[dst];
  const Camera camDst = c.cams[self].rescale(W, H);
  std::vector<Vec2f> warpInv(n);
  for (int s = 0; s < c.S; ++s) {
    const Camera camSrc = c.cams[s].rescale(W, H);
    c.projWarp[s].resize(n);
    // dstProjWarp = computeWarpDstToSrc(camSrc, camDst): src-sized, src px -> dst px (Derp.cpp:970)
    computeWarpDstToSrc(camSrc, camDst, s == self, c.projWarp[s].data());
    if (s == self) {
      c.projColor[s] = c.color[s];  // Derp.cpp:989-991
    } else {
      // dstProjWarpInv = computeWarpDstToSrc(camDst, camSrc): dst px -> src px (Derp.cpp:971)
      computeWarpDstToSrc(camDst, camSrc, false, warpInv.data());
      c.projColor[s].resize(n * 3);
      const uint16_t* src = c.color[s].data();
      uint16_t* out = c.projColor[s].data();
      const float* map = (const float*)warpInv.data();
      parallelFor(0, H, [&](int y0, int y1) { remapBicubicU16C3(src, W, H, map, W, H, out, y0, y1); });
    }
    c.projBias[s].resize(n * 3);
    const uint16_t* pc = c.projColor[s].data();
    uint16_t* pb = c.projBias[s].data();
    parallelFor(0, H, [&](int y0, int y1) { blur3x3U16C3(pc, W, H, pb, y0, y1); });
  }
  c.projDst = dst;
  return DERP_OK;
}

int derp_eval_cost(DerpCtx* ctx, int dst, const float* disparity, float* out_cost, float* out_conf) {
  if (!ctx || !disparity) return fail(DERP_EINVAL, "derp_eval_cost: bad arguments");
  Ctx& c = ctx->c;
  int rc = checkDst(c, dst, "derp_eval_cost", true);
  if (rc) return rc;
  const int W = c.W, H = c.H;
  const float nan = std::numeric_limits<float>::quiet_NaN();
  c.costEvals = 0;
  c.srcHits = 0;
  parallelFor(0, H, [&](int y0, int y1) {
    uint64_t hits = 0, evals = 0;
    for (int y = y0; y < y1; ++y)
      for (int x = 0; x < W; ++x) {
        float co = nan, cf = nan;
        if (x >= 1 && x < W - 1 && y >= 1 && y < H - 1) {
          computeCost(c, dst, disparity[(size_t)y * W + x], x, y, &co, &cf, &hits);
          ++evals;
        }
        if (out_cost) out_cost[(size_t)y * W + x] = co;
        if (out_conf) out_conf[(size_t)y * W + x] = cf;
      }
    c.costEvals += evals;
    c.srcHits += hits;
  });
  return DERP_OK;
}

// computeBruteForceDisparity (Derp.cpp:264-382), pixel-major: identical results to the reference's
// slice-major volume + strict-< scan because the per-pixel candidate order is the same.
int derp_brute_force(DerpCtx* ctx, int dst, int num_depths, float min_depth_m, float max_depth_m,
                     int partial_coverage, int32_t* best_index) {
  if (!ctx || num_depths < 2) return fail(DERP_EINVAL, "derp_brute_force: bad arguments");
  Ctx& c = ctx->c;
  int rc = checkDst(c, dst, "derp_brute_force", true);
  if (rc) return rc;
  const int W = c.W, H = c.H;
  const int self = c.dst2src[dst];
  const bool useFg = c.lp.use_foreground_masks != 0;
  if (useFg && c.bgDisp[dst].empty()) return fail(DERP_ESTATE, "derp_brute_force: background disparity not set");
  std::vector<float> disparities(num_depths);
  const float minDisparity = 1.0f / max_depth_m;
  const float maxDisparity = 1.0f / min_depth_m;
  for (int i = 0; i < num_depths; ++i) {
    // probeDisparity (ImageUtil.cpp:100-107), double math, stored as float
    const double fraction = double(i) / double(num_depths - 1);
    disparities[i] = (float)(fraction * (double)minDisparity + (1 - fraction) * (double)maxDisparity);
  }
  float* dstDisparity = c.disp[dst].data();
  float* dstCosts = c.cost[dst].data();
  float* dstConf = c.conf[dst].data();
  const uint8_t* fov = c.fovMask[dst].data();
  const uint8_t* fg = c.fgMask[self].data();
  const float* bg = useFg ? c.bgDisp[dst].data() : nullptr;
  std::vector<int32_t> idxLocal;
  int32_t* idx = best_index;
  if (!idx) {
    idxLocal.resize((size_t)W * H);
    idx = idxLocal.data();
  }
  const int margin = kSearchWindowRadius;
  std::atomic<int> uncovered(0);
  c.costEvals = 0;
  c.srcHits = 0;
  parallelFor(margin, H - margin, [&](int y0, int y1) {
    uint64_t hits = 0, evals = 0;
    for (int y = y0; y < y1; ++y)
      for (int x = margin; x < W - margin; ++x) {
        const size_t p = (size_t)y * W + x;
        if (!fov[p]) {
          dstDisparity[p] = std::numeric_limits<float>::quiet_NaN();
          idx[p] = -2;
          continue;
        }
        if (!fg[p]) {
          dstDisparity[p] = bg[p];
          idx[p] = -3;
          continue;
        }
        float minCost = FLT_MAX, minCostConfidence = 0;
        int best = -1;
        for (int i = 0; i < num_depths; ++i) {
          // computeBruteForceCosts (Derp.cpp:240-258): closerMask = bg < disparity
          if (useFg && !(bg[p] < disparities[i])) continue;  // cost NaN, never wins
          float cost, conf;
          computeCost(c, dst, disparities[i], x, y, &cost, &conf, &hits);
          ++evals;
          if (cost < minCost) {
            minCost = cost;
            minCostConfidence = conf;
            best = i;
          }
        }
        if (best == -1) {
          if (!(partial_coverage || useFg)) uncovered++;
          dstDisparity[p] = minDisparity;
        } else {
          dstDisparity[p] = disparities[best];
        }
        idx[p] = best;
        dstCosts[p] = minCost;
        dstConf[p] = minCostConfidence;
      }
    c.costEvals += evals;
    c.srcHits += hits;
  });
  if (uncovered.load() > 0)  // Derp.cpp:339 CHECK(partialCoverage || useForegroundMasks)
    return fail(DERP_ECOVERAGE, "Insufficient coverage at " + std::to_string(uncovered.load()) + " pixels");
  // Extend disparities to margin (Derp.cpp:359-381)
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x)
      if (x < margin || x >= W - margin || y < margin || y >= H - margin) {
        const size_t p = (size_t)y * W + x;
        if (!fg[p]) {
          dstDisparity[p] = bg[p];
          idx[p] = -3;
          continue;
        }
        const size_t q = (size_t)clampi(y, margin, H - margin - 1) * W + clampi(x, margin, W - margin - 1);
        dstDisparity[p] = dstDisparity[q];
        dstCosts[p] = dstCosts[q];
        dstConf[p] = dstConf[q];
        idx[p] = idx[q];
      }
  return DERP_OK;
}

// randomProposal / randomProposals (Derp.cpp:750-873)
int derp_random_proposals(DerpCtx* ctx, int dst, int num_proposals, float min_depth_m, float max_depth_m) {
  if (!ctx) return fail(DERP_EINVAL, "null ctx");
  Ctx& c = ctx->c;
  int rc = checkDst(c, dst, "derp_random_proposals", true);
  if (rc) return rc;
  const int W = c.W, H = c.H;
  const int self = c.dst2src[dst];
  const bool useFg = c.lp.use_foreground_masks != 0;
  float* dstDisparity = c.disp[dst].data();
  float* dstCosts = c.cost[dst].data();
  float* dstConfidence = c.conf[dst].data();
  const float* variance = c.variance[self].data();
  const uint8_t* fov = c.fovMask[dst].data();
  const uint8_t* fg = c.fgMask[self].data();
  const float* bg = useFg ? c.bgDisp[dst].data() : nullptr;
  const float kRandomPropMaxCost = 5.0f, kRandomPropHighVarDeviation = 0.1f;  // Derp.h:36-37
  const int level = c.lp.level;
  c.costEvals = 0;
  c.srcHits = 0;
  parallelFor(kSearchWindowRadius, H - kSearchWindowRadius, [&](int y0, int y1) {
    uint64_t hits = 0, evals = 0;
    for (int y = y0; y < y1; ++y) {
      std::default_random_engine engine;
      engine.seed(y * level);
      for (int x = kSearchWindowRadius; x < W - kSearchWindowRadius; ++x) {
        const size_t p = (size_t)y * W + x;
        if (!fov[p]) continue;
        float currDisp = dstDisparity[p];
        if (!fg[p]) {
          dstDisparity[p] = bg[p];
          continue;
        }
        const float varHighDev = kRandomPropHighVarDeviation * c.lp.var_high_thresh;
        const float varHighThresh = std::max(varHighDev, c.varNoiseFloor);
        if (variance[p] < varHighThresh) continue;
        float currCost, currConfidence;
        computeCost(c, dst, currDisp, x, y, &currCost, &currConfidence, &hits);
        ++evals;
        const float costThresh = std::fmin(0.5f * currCost, kRandomPropMaxCost);
        const float minDisp = useFg ? bg[p] : (1.0f / max_depth_m);
        const float maxDisp = 1.0f / min_depth_m;
        float amplitude = (maxDisp - minDisp) / 2.0f;
        for (int i = 0; i < num_proposals; ++i) {
          float propDisp = std::uniform_real_distribution<float>(
              std::max(float(minDisp), currDisp - amplitude),
              std::min(float(maxDisp), currDisp + amplitude))(engine);
          float propCost, propConfidence;
          computeCost(c, dst, propDisp, x, y, &propCost, &propConfidence, &hits);
          ++evals;
          if (propCost < currCost && propCost < costThresh) {
            currCost = propCost;
            currDisp = propDisp;
            currConfidence = propConfidence;
            amplitude /= 2.0f;
          }
        }
        dstDisparity[p] = currDisp;
        dstCosts[p] = currCost;
        dstConfidence[p] = currConfidence;
      }
    }
    c.costEvals += evals;
    c.srcHits += hits;
  });
  return DERP_OK;
}

// pingPongRectangle / pingPong (Derp.cpp:403-538)
int derp_ping_pong(DerpCtx* ctx, int dst, int iterations) {
  if (!ctx) return fail(DERP_EINVAL, "null ctx");
  Ctx& c = ctx->c;
  int rc = checkDst(c, dst, "derp_ping_pong", true);
  if (rc) return rc;
  const int W = c.W, H = c.H;
  const size_t n = (size_t)W * H;
  const int self = c.dst2src[dst];
  const bool useFg = c.lp.use_foreground_masks != 0;
  static const int kOffsets[9][2] = {{0, 0}, {-1, 0}, {1, 0}, {0, -1}, {0, 1},
                                     {-2, -2}, {2, -2}, {-2, 2}, {2, 2}};  // DerpUtil.h:34-43
  std::vector<float>& disp = c.disp[dst];
  std::vector<float>& costs = c.cost[dst];
  const std::vector<float>& confidences = c.conf[dst];
  const uint8_t* maskFov = c.fovMask[dst].data();
  const uint8_t* fg = c.fgMask[self].data();
  const float* dispBackground = useFg ? c.bgDisp[dst].data() : nullptr;
  const float* variance = c.variance[self].data();
  std::vector<float> dispRes(disp);
  std::vector<float> costsRes(n, INFINITY);
  std::vector<float> confidencesRes(n, 0.f);
  std::vector<uint8_t> changed(n, 1);
  c.costEvals = 0;
  c.srcHits = 0;
  for (int it = 1; it <= iterations; ++it) {
    const int radius = kSearchWindowRadius;
    parallelFor(radius, H - radius, [&](int y0, int y1) {
      uint64_t hits = 0, evals = 0;
      for (int y = y0; y < y1; ++y)
        for (int x = radius; x < W - radius; ++x) {
          const size_t p = (size_t)y * W + x;
          if (!maskFov[p]) continue;
          if (!fg[p]) {
            dispRes[p] = dispBackground[p];
            continue;
          }
          if (variance[p] < c.varNoiseFloor) continue;
          float bestCost = INFINITY;
          float bestDisparity = disp[p];
          float bestConfidence = confidences[p];
          const float backgroundDisparity = useFg ? dispBackground[p] : 0;
          for (int k = 0; k < 9; ++k) {
            const int xx = clampi(x + kOffsets[k][0], 0, W - 1);
            const int yy = clampi(y + kOffsets[k][1], 0, H - 1);
            const size_t q = (size_t)yy * W + xx;
            if (maskFov[q]) {
              const float d = disp[q];
              if (d >= backgroundDisparity && changed[q]) {
                float cost, conf;
                computeCost(c, dst, d, x, y, &cost, &conf, &hits);
                ++evals;
                if (cost < bestCost) {
                  bestCost = cost;
                  bestDisparity = d;
                  bestConfidence = conf;
                }
              }
            }
          }
          dispRes[p] = bestDisparity;
          costsRes[p] = bestCost;
          confidencesRes[p] = bestConfidence;
        }
      c.costEvals += evals;
      c.srcHits += hits;
    });
    // changed = disp != dispRes (cv compare: NaN != NaN is true)
    for (size_t i = 0; i < n; ++i) changed[i] = (disp[i] != dispRes[i]) ? 1 : 0;
    disp = dispRes;
    costs = costsRes;  // INF where skipped; confidence is not written back (Derp.cpp:527-529)
  }
  return DERP_OK;
}

// getSrcMismatches / updateDstDisparityAndMismatchMask / handleDisparityMismatch(es) (Derp.cpp:553-748)
// allDisp[s] = pre-update disparity of rig camera s (every camera); updates this context's destinations.
static int mismatchesImpl(Ctx& c, const std::vector<const float*>& allDisp) {
  const int W = c.W, H = c.H;
  const size_t n = (size_t)W * H;
  std::vector<std::vector<float>> newDisp(c.Sd);
  for (int dstIdx = 0; dstIdx < c.Sd; ++dstIdx) {
    const float* dstDisp = allDisp[c.dst2src[dstIdx]];
    std::vector<uint8_t>& dstMask = c.mismatch[dstIdx];
    std::vector<float>& dstDispNew = newDisp[dstIdx];
    dstDispNew.assign(n, std::numeric_limits<float>::quiet_NaN());
    const std::vector<float>& dstVar = c.variance[c.dst2src[dstIdx]];
    const Camera& camDst = c.cams[c.dst2src[dstIdx]];
    const uint8_t* fov = c.fovMask[dstIdx].data();
    const uint8_t* fg = c.fgMask[c.dst2src[dstIdx]].data();
    parallelFor(0, H, [&](int y0, int y1) {
      std::vector<float> dispMatches, dispMismatches;
      for (int y = y0; y < y1; ++y)
        for (int x = 0; x < W; ++x) {
          const size_t p = (size_t)y * W + x;
          if (!fov[p]) continue;
          dispMatches.clear();
          dispMismatches.clear();
          if (fg[p]) {
            double ptWorld[3];
            dstToWorldPoint(camDst, x, y, dstDisp[p], W, H, ptWorld);
            for (int srcIdx = 0; srcIdx < c.S; ++srcIdx) {
              if (srcIdx == c.dst2src[dstIdx]) continue;
              double ptSrc[2];
              if (!worldToSrcPoint(ptSrc, ptWorld, c.cams[srcIdx], W, H)) continue;
              const float dSrc = bilinearF32(allDisp[srcIdx], W, H, (float)ptSrc[0], (float)ptSrc[1]);
              static const float kFractionChange = 0.1f;
              const float dDstMin = (1.0f - kFractionChange) * dstDisp[p];
              const float dDstMax = (1.0f + kFractionChange) * dstDisp[p];
              if (dDstMin <= dSrc && dSrc <= dDstMax) dispMatches.push_back(dSrc);
              else dispMismatches.push_back(dSrc);
            }
          }
          const float dispCurr = dstDisp[p];
          if (dispMatches.size() + dispMismatches.size() == 0) {
            dstMask[p] = 0;
            dstDispNew[p] = dispCurr;
            continue;
          }
          static const int kNumMinSrcCams = kMinOverlappingCams - 1;
          if (int(dispMatches.size()) >= kNumMinSrcCams || c.lp.var_high_thresh < dstVar[p] ||
              dstVar[p] < c.varNoiseFloor) {
            dstMask[p] = 0;
            dstDispNew[p] = dispCurr;
          } else {
            dstMask[p] = 1;
            std::sort(dispMismatches.begin(), dispMismatches.end());
            int closer;
            for (closer = 0; closer < int(dispMismatches.size()); ++closer)
              if (dispMismatches[closer] >= dispCurr) break;
            const int median = closer / 2;
            dstDispNew[p] = std::min(dispCurr, dispMismatches[median]);
          }
        }
    });
  }
  for (int d = 0; d < c.Sd; ++d) c.disp[d] = newDisp[d];
  return DERP_OK;
}

int derp_mismatches(DerpCtx* ctx) {
  if (!ctx) return fail(DERP_EINVAL, "null ctx");
  Ctx& c = ctx->c;
  if (!c.levelOpen) return fail(DERP_ESTATE, "derp_mismatches: no level");
  if (c.Sd != c.S) return fail(DERP_EINVAL, "Mismatches only valid when considering all cameras");
  for (int d = 0; d < c.Sd; ++d)
    if (c.dst2src[d] != d) return fail(DERP_EINVAL, "derp_mismatches: dst list must equal camera list");
  std::vector<const float*> all(c.S);
  for (int s = 0; s < c.S; ++s) all[s] = c.disp[s].data();
  return mismatchesImpl(c, all);
}

// Destination cameras dealt to several contexts (include/derp_b200.h): same Jacobi update, the other
// cameras' pre-update planes come from the caller's all-gather.
const float* derp_disparity_device_ptr(DerpCtx* ctx, int dst) {
  if (!ctx) return nullptr;
  Ctx& c = ctx->c;
  if (checkDst(c, dst, "derp_disparity_device_ptr", false)) return nullptr;
  return c.disp[dst].data();
}

int derp_gather_disparities(DerpCtx* ctx, const float* const* planes) {
  if (!ctx || !planes) return fail(DERP_EINVAL, "derp_gather_disparities: bad arguments");
  Ctx& c = ctx->c;
  if (!c.levelOpen) return fail(DERP_ESTATE, "derp_gather_disparities: no level");
  const size_t n = (size_t)c.W * c.H;
  std::vector<int> ownDst(c.S, -1);
  for (int d = 0; d < c.Sd; ++d) ownDst[c.dst2src[d]] = d;
  c.gathered.assign(c.S, std::vector<float>());
  for (int s = 0; s < c.S; ++s) {
    if (planes[s]) c.gathered[s].assign(planes[s], planes[s] + n);
    else if (ownDst[s] >= 0) c.gathered[s] = c.disp[ownDst[s]];
    else return fail(DERP_EINVAL, "derp_gather_disparities: no plane for a camera this context does not own");
  }
  c.haveGathered = true;
  return DERP_OK;
}

int derp_mismatches_gathered(DerpCtx* ctx) {
  if (!ctx) return fail(DERP_EINVAL, "null ctx");
  Ctx& c = ctx->c;
  if (!c.levelOpen) return fail(DERP_ESTATE, "derp_mismatches_gathered: no level");
  if (!c.haveGathered) return fail(DERP_ESTATE, "derp_mismatches_gathered: derp_gather_disparities not called for this level");
  c.haveGathered = false;
  std::vector<const float*> all(c.S);
  for (int s = 0; s < c.S; ++s) all[s] = c.gathered[s].data();
  return mismatchesImpl(c, all);
}

// generalizedJointBilateralFilter (TemporalBilateralFilter.h:39-124).  guideScale = 1/maxPixelValue.
extern "C++" {
template <typename TG>
static void jointBilateral(const float* image, const TG* guide, float guideFactor, const uint8_t* mask,
                           int W, int H, int radius, float sigma, float weight0, float weight1,
                           float weight2, float* dest) {
  parallelFor(0, H, [&](int y0, int y1) {
    for (int y = y0; y < y1; ++y)
      for (int x = 0; x < W; ++x) {
        const size_t p = (size_t)y * W + x;
        if (!mask[p]) {
          dest[p] = image[p];
          continue;
        }
        const TG* guideColor = guide + p * 3;
        float sumWeight = 0.0f;
        float weightedAvg = 0.0f;
        for (int v = -radius; v <= radius; ++v)
          for (int u = -radius; u <= radius; ++u) {
            const int sampleX = clampi(x + u, 0, W - 1);
            const int sampleY = clampi(y + v, 0, H - 1);
            const size_t q = (size_t)sampleY * W + sampleX;
            if (!mask[q]) continue;
            const TG* nb = guide + q * 3;
            const float d0 = (guideColor[0] * guideFactor) - (nb[0] * guideFactor);
            const float d1 = (guideColor[1] * guideFactor) - (nb[1] * guideFactor);
            const float d2 = (guideColor[2] * guideFactor) - (nb[2] * guideFactor);
            const float colorDiffSq = weight0 * (d0 * d0) + weight1 * (d1 * d1) + weight2 * (d2 * d2);
            const float weight = expf((-colorDiffSq / 3.0f) / (2.0f * (sigma * sigma)));
            sumWeight += weight;
            weightedAvg += weight * image[q];
          }
        if (sumWeight != 0.0f) {
          weightedAvg /= sumWeight;
          dest[p] = weightedAvg;
        } else {
          dest[p] = image[p];
        }
      }
  });
}
}  // extern "C++"

// bilateralFilter (Derp.cpp:875-902)
int derp_bilateral(DerpCtx* ctx, int dst) {
  if (!ctx) return fail(DERP_EINVAL, "null ctx");
  Ctx& c = ctx->c;
  int rc = checkDst(c, dst, "derp_bilateral", false);
  if (rc) return rc;
  if (!c.haveColors) return fail(DERP_ESTATE, "derp_bilateral: colours not set");
  const int W = c.W, H = c.H;
  const size_t n = (size_t)W * H;
  const int self = c.dst2src[dst];
  const float kLevelScale = 0.9f;
  const float scale = std::pow(kLevelScale, c.lp.level);
  const int spaceRadius = std::max(std::ceil(5 * scale), float(1));
  std::vector<uint8_t> mask(n);
  for (size_t i = 0; i < n; ++i) mask[i] = c.fovMask[dst][i] & c.fgMask[self][i];
  std::vector<float> filtered(n);
  jointBilateral<uint16_t>(c.disp[dst].data(), c.color[self].data(), 1 / 65535.0f, mask.data(), W, H,
                           spaceRadius, 0.005f, 0.5f, 1.0f, 1.0f, filtered.data());
  for (size_t i = 0; i < n; ++i)
    if (c.fgMask[self][i]) c.disp[dst][i] = filtered[i];
  return DERP_OK;
}

// maskedMedianBlur (CvUtil.h:336-385) / medianFilter (Derp.cpp:904-920)
int derp_median(DerpCtx* ctx, int dst) {
  if (!ctx) return fail(DERP_EINVAL, "null ctx");
  Ctx& c = ctx->c;
  int rc = checkDst(c, dst, "derp_median", false);
  if (rc) return rc;
  const int W = c.W, H = c.H;
  const size_t n = (size_t)W * H;
  const int self = c.dst2src[dst];
  const int radius = 1;
  const std::vector<float>& mat = c.disp[dst];
  const std::vector<float>& background = c.bgDisp[dst];
  std::vector<uint8_t> mask(n);
  for (size_t i = 0; i < n; ++i) mask[i] = c.fovMask[dst][i] & c.fgMask[self][i];
  std::vector<float> blurred(n, 0.0f);
  parallelFor(0, H, [&](int y0, int y1) {
    std::vector<float> values;
    for (int y = y0; y < y1; ++y)
      for (int x = 0; x < W; ++x) {
        const size_t p = (size_t)y * W + x;
        values.clear();
        if (!mask[p]) {
          if (!background.empty()) blurred[p] = background[p];
          continue;
        }
        for (int yy = y - radius; yy <= y + radius; ++yy)
          for (int xx = x - radius; xx <= x + radius; ++xx) {
            if (0 > yy || yy >= H || 0 > xx || xx >= W) continue;
            const size_t q = (size_t)yy * W + xx;
            if (!mask[q]) continue;
            if (std::isnan(mat[q]) || mat[q] == 0) continue;
            values.push_back(mat[q]);
          }
        if (!values.empty()) {
          const size_t m = values.size() / 2;
          std::partial_sort(values.begin(), values.begin() + m + 1, values.end());
          if (values.size() % 2 == 1) blurred[p] = values[m];
          else blurred[p] = (values[m - 1] + values[m]) / 2.0;
        }
      }
  });
  c.disp[dst] = blurred;
  return DERP_OK;
}

// maskFov (Derp.cpp:940-951)
int derp_mask_fov(DerpCtx* ctx, int dst) {
  if (!ctx) return fail(DERP_EINVAL, "null ctx");
  Ctx& c = ctx->c;
  int rc = checkDst(c, dst, "derp_mask_fov", false);
  if (rc) return rc;
  const size_t n = (size_t)c.W * c.H;
  for (size_t i = 0; i < n; ++i)
    if (!c.fovMask[dst][i]) c.disp[dst][i] = std::numeric_limits<float>::quiet_NaN();
  return DERP_OK;
}

// ---- upsampling (UpsampleDisparityLib.cpp:27-182) ----------------------------------------------
static std::vector<std::pair<int, int>> spiral(const int w) {
  int x = 0, y = 0, dx = 0, dy = -1, t = w;
  const int samples = t * t;
  std::vector<std::pair<int, int>> locs;
  for (int i = 0; i < samples; ++i) {
    const bool isValidX = (-w / 2 <= x) && (x <= w / 2);
    const bool isValidY = (-w / 2 <= y) && (y <= w / 2);
    if (isValidX && isValidY) locs.emplace_back(x, y);
    const bool isCorner = x == y;
    const bool isEdgeLeftX = (x < 0) && (x == -y);
    const bool isEdgeRightX = (x > 0) && (x == 1 - y);
    if (isCorner || isEdgeLeftX || isEdgeRightX) {
      t = dx;
      dx = -dy;
      dy = t;
    }
    x += dx;
    y += dy;
  }
  return locs;
}

static void replaceNans(const std::vector<float>& dispUp, const float* bgDispUp, const std::vector<uint8_t>& maskUp,
                        int W, int H, int radius, std::vector<float>& dispOut) {
  dispOut = dispUp;
  const std::vector<std::pair<int, int>> spiralLocs = spiral(radius * 2 + 1);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const size_t p = (size_t)y * W + x;
      // maskNan = maskUp with (dispUp > 0) cleared
      if (!maskUp[p] || dispUp[p] > 0) continue;
      for (const auto& loc : spiralLocs) {
        const int xx = clampi(x + loc.first, 0, W - 1);
        const int yy = clampi(y + loc.second, 0, H - 1);
        const float d = dispUp[(size_t)yy * W + xx];
        if (d > 0) {
          dispOut[p] = d;
          break;
        }
      }
    }
  for (size_t i = 0; i < dispOut.size(); ++i)
    if (std::isnan(dispOut[i]) || dispOut[i] == 0) dispOut[i] = bgDispUp ? bgDispUp[i] : dispOut[i];
}

// upsampleDisparityInPlace (UpsampleDisparityLib.cpp:98-147); mask/maskUp already AND-ed with fov
static int upsampleOne(const float* disp, int cw, int ch, const float* bgDispUp, const uint8_t* mask,
                       const uint8_t* maskUp, int W, int H, bool useForegroundMasks, float* dispUp) {
  if (useForegroundMasks) {
    if (!bgDispUp) return fail(DERP_EINVAL, "upsample: background disparity required with foreground masks");
    // getRadius (UpsampleDisparityLib.cpp:93-96): float(sizeUp.width)/float(mask.width)
    const float scale = float(W) / float(cw);
    const int radius = scale * scale + 1;
    std::vector<float> dispSmallMasked(disp, disp + (size_t)cw * ch);
    for (size_t i = 0; i < dispSmallMasked.size(); ++i)
      if (!mask[i]) dispSmallMasked[i] = NAN;
    std::vector<float> dispUpMasked((size_t)W * H);
    resizeNearest<float>(dispSmallMasked.data(), cw, ch, dispUpMasked.data(), W, H);
    std::vector<uint8_t> mu(maskUp, maskUp + (size_t)W * H);
    for (size_t i = 0; i < dispUpMasked.size(); ++i)
      if (!mu[i]) dispUpMasked[i] = NAN;
    std::vector<float> out;
    replaceNans(dispUpMasked, bgDispUp, mu, W, H, radius, out);
    std::memcpy(dispUp, out.data(), out.size() * sizeof(float));
  } else {
    const float minDisp = 1e-4;
    std::vector<float> dispSmallMasked(disp, disp + (size_t)cw * ch);
    for (float& v : dispSmallMasked)
      if (v != v) v = minDisp;
    resizeLanczos4F32(dispSmallMasked.data(), cw, ch, dispUp, W, H);
  }
  return DERP_OK;
}

int derp_upsample_from(DerpCtx* ctx, int dst, const float* coarse, int coarse_w, int coarse_h,
                       const uint8_t* coarse_mask, const uint8_t* fine_mask) {
  if (!ctx || !coarse || coarse_w < 1 || coarse_h < 1) return fail(DERP_EINVAL, "derp_upsample_from: bad arguments");
  Ctx& c = ctx->c;
  int rc = checkDst(c, dst, "derp_upsample_from", false);
  if (rc) return rc;
  const bool useFg = c.lp.use_foreground_masks != 0;
  const int W = c.W, H = c.H;
  std::vector<uint8_t> m, mu;
  if (useFg) {
    if (!coarse_mask || !fine_mask) return fail(DERP_EINVAL, "derp_upsample_from: masks required");
    // fovMasks at both sizes (UpsampleDisparityLib.cpp:163-176)
    const Camera& cam = c.cams[c.dst2src[dst]];
    std::vector<uint8_t> fovC((size_t)coarse_w * coarse_h);
    generateFovMask(cam, coarse_w, coarse_h, fovC.data());
    m.resize(fovC.size());
    for (size_t i = 0; i < m.size(); ++i) m[i] = fovC[i] && coarse_mask[i];
    mu.resize((size_t)W * H);
    for (size_t i = 0; i < mu.size(); ++i) mu[i] = c.fovMask[dst][i] && fine_mask[i];
  }
  return upsampleOne(coarse, coarse_w, coarse_h, useFg ? c.bgDisp[dst].data() : nullptr,
                     useFg ? m.data() : nullptr, useFg ? mu.data() : nullptr, W, H, useFg,
                     c.disp[dst].data());
}

int derp_level_keep(DerpCtx* ctx) {
  if (!ctx) return fail(DERP_EINVAL, "derp_level_keep: null context");
  Ctx& c = ctx->c;
  if (c.W < 1) return fail(DERP_ESTATE, "derp_level_keep: no level is open");
  c.kept = c.disp;
  c.keptW = c.W;
  c.keptH = c.H;
  return DERP_OK;
}
int derp_upsample_from_kept(DerpCtx* ctx, int dst, const uint8_t* coarse_mask, const uint8_t* fine_mask) {
  if (!ctx) return fail(DERP_EINVAL, "derp_upsample_from_kept: null context");
  Ctx& c = ctx->c;
  if (c.keptW < 1 || dst < 0 || dst >= (int)c.kept.size()) return fail(DERP_ESTATE, "derp_upsample_from_kept: nothing kept");
  return derp_upsample_from(ctx, dst, c.kept[dst].data(), c.keptW, c.keptH, coarse_mask, fine_mask);
}

int derp_upsample_disparity(int device, const DerpCameraDesc* cam, const float* coarse, int coarse_w,
                            int coarse_h, const float* background_up, const uint8_t* coarse_mask,
                            const uint8_t* fine_mask, int out_w, int out_h, int use_foreground_masks,
                            float* out) {
  (void)device;
  if (!cam || !coarse || !out) return fail(DERP_EINVAL, "derp_upsample_disparity: bad arguments");
  std::vector<uint8_t> m, mu;
  if (use_foreground_masks) {
    if (!coarse_mask || !fine_mask) return fail(DERP_EINVAL, "derp_upsample_disparity: masks required");
    Camera c;
    if (!c.init(*cam)) return fail(DERP_EINVAL, "derp_upsample_disparity: invalid camera");
    c.normalize();
    std::vector<uint8_t> fovC((size_t)coarse_w * coarse_h), fovU((size_t)out_w * out_h);
    generateFovMask(c, coarse_w, coarse_h, fovC.data());
    generateFovMask(c, out_w, out_h, fovU.data());
    m.resize(fovC.size());
    mu.resize(fovU.size());
    for (size_t i = 0; i < m.size(); ++i) m[i] = fovC[i] && coarse_mask[i];
    for (size_t i = 0; i < mu.size(); ++i) mu[i] = fovU[i] && fine_mask[i];
  }
  return upsampleOne(coarse, coarse_w, coarse_h, background_up, use_foreground_masks ? m.data() : nullptr,
                     use_foreground_masks ? mu.data() : nullptr, out_w, out_h, use_foreground_masks != 0, out);
}

// processLevel (Derp.cpp:1004-1035) minus saveResults, in the two halves include/derp_b200.h declares
int derp_level_estimate(DerpCtx* ctx, const DerpProcessOpts* o) {
  if (!ctx || !o) return fail(DERP_EINVAL, "derp_level_estimate: bad arguments");
  Ctx& c = ctx->c;
  if (!c.levelOpen || !c.haveColors) return fail(DERP_ESTATE, "derp_level_estimate: level/colours not set");
  const bool coarsest = c.lp.level == c.lp.num_levels - 1;
  uint64_t evals = 0, hits = 0;
  for (int d = 0; d < c.Sd; ++d) {
    int rc = derp_reproject(ctx, d);
    if (rc) return rc;
    if (coarsest) {  // preprocessLevel (Derp.cpp:826-842)
      rc = derp_brute_force(ctx, d, o->num_depths, o->min_depth_m, o->max_depth_m, o->partial_coverage, nullptr);
      if (rc) return rc;
      evals += c.costEvals;
      hits += c.srcHits;
    }
    if (o->random_proposals > 0 && !coarsest) {  // Derp.cpp:851-853
      rc = derp_random_proposals(ctx, d, o->random_proposals, o->min_depth_m, o->max_depth_m);
      if (rc) return rc;
      evals += c.costEvals;
      hits += c.srcHits;
    }
    if (!coarsest) {  // Derp.cpp:545-547
      rc = derp_ping_pong(ctx, d, o->ping_pong_iterations);
      if (rc) return rc;
      evals += c.costEvals;
      hits += c.srcHits;
    }
  }
  c.costEvals = evals;
  c.srcHits = hits;
  return DERP_OK;
}

int derp_level_filter(DerpCtx* ctx, const DerpProcessOpts* o) {
  if (!ctx || !o) return fail(DERP_EINVAL, "derp_level_filter: bad arguments");
  Ctx& c = ctx->c;
  if (!c.levelOpen || !c.haveColors) return fail(DERP_ESTATE, "derp_level_filter: level/colours not set");
  for (int d = 0; d < c.Sd; ++d) {
    int rc;
    if (o->do_bilateral_filter && (rc = derp_bilateral(ctx, d))) return rc;
    if (o->do_median_filter && (rc = derp_median(ctx, d))) return rc;
    if ((rc = derp_mask_fov(ctx, d))) return rc;
  }
  return DERP_OK;
}

int derp_process_level(DerpCtx* ctx, const DerpProcessOpts* o) {
  int rc = derp_level_estimate(ctx, o);
  if (rc) return rc;
  Ctx& c = ctx->c;
  const bool coarsest = c.lp.level == c.lp.num_levels - 1;
  if (!(c.lp.level > o->mismatches_start_level || coarsest)) {  // Derp.cpp:726-728
    if ((rc = derp_mismatches(ctx))) return rc;
  }
  return derp_level_filter(ctx, o);
}

// ---- state access ------------------------------------------------------------------------------
int derp_set_disparity(DerpCtx* ctx, int dst, const float* disparity, const float* cost, const float* confidence) {
  if (!ctx) return fail(DERP_EINVAL, "null ctx");
  Ctx& c = ctx->c;
  int rc = checkDst(c, dst, "derp_set_disparity", false);
  if (rc) return rc;
  const size_t n = (size_t)c.W * c.H;
  if (disparity) c.disp[dst].assign(disparity, disparity + n);
  if (cost) c.cost[dst].assign(cost, cost + n);
  if (confidence) c.conf[dst].assign(confidence, confidence + n);
  return DERP_OK;
}
int derp_get_disparity(DerpCtx* ctx, int dst, float* disparity, float* cost, float* confidence) {
  if (!ctx) return fail(DERP_EINVAL, "null ctx");
  Ctx& c = ctx->c;
  int rc = checkDst(c, dst, "derp_get_disparity", false);
  if (rc) return rc;
  const size_t n = (size_t)c.W * c.H * sizeof(float);
  if (disparity) std::memcpy(disparity, c.disp[dst].data(), n);
  if (cost) std::memcpy(cost, c.cost[dst].data(), n);
  if (confidence) std::memcpy(confidence, c.conf[dst].data(), n);
  return DERP_OK;
}
int derp_get_fov_mask(DerpCtx* ctx, int dst, uint8_t* mask) {
  if (!ctx || !mask) return fail(DERP_EINVAL, "bad arguments");
  Ctx& c = ctx->c;
  int rc = checkDst(c, dst, "derp_get_fov_mask", false);
  if (rc) return rc;
  std::memcpy(mask, c.fovMask[dst].data(), (size_t)c.W * c.H);
  return DERP_OK;
}
int derp_get_mismatch_mask(DerpCtx* ctx, int dst, uint8_t* mask) {
  if (!ctx || !mask) return fail(DERP_EINVAL, "bad arguments");
  Ctx& c = ctx->c;
  int rc = checkDst(c, dst, "derp_get_mismatch_mask", false);
  if (rc) return rc;
  std::memcpy(mask, c.mismatch[dst].data(), (size_t)c.W * c.H);
  return DERP_OK;
}
int derp_get_variance(DerpCtx* ctx, int src, float* variance) {
  if (!ctx || !variance) return fail(DERP_EINVAL, "bad arguments");
  Ctx& c = ctx->c;
  if (!c.levelOpen || !c.haveColors) return fail(DERP_ESTATE, "derp_get_variance: colours not set");
  if (src < 0 || src >= c.S) return fail(DERP_EINVAL, "src out of range");
  std::memcpy(variance, c.variance[src].data(), (size_t)c.W * c.H * sizeof(float));
  return DERP_OK;
}
int derp_get_var_noise_floor(DerpCtx* ctx, float* out) {
  if (!ctx || !out || !ctx->c.levelOpen) return fail(DERP_EINVAL, "bad arguments");
  *out = ctx->c.varNoiseFloor;
  return DERP_OK;
}
static int checkProj(Ctx& c, int src, const char* who) {
  if (!c.levelOpen || c.projDst < 0) return fail(DERP_ESTATE, std::string(who) + ": no projection tables");
  if (src < 0 || src >= c.S) return fail(DERP_EINVAL, std::string(who) + ": src out of range");
  return DERP_OK;
}
int derp_get_proj_warp(DerpCtx* ctx, int src, float* warp_xy) {
  if (!ctx || !warp_xy) return fail(DERP_EINVAL, "bad arguments");
  Ctx& c = ctx->c;
  int rc = checkProj(c, src, "derp_get_proj_warp");
  if (rc) return rc;
  std::memcpy(warp_xy, c.projWarp[src].data(), (size_t)c.W * c.H * sizeof(Vec2f));
  return DERP_OK;
}
int derp_get_proj_color(DerpCtx* ctx, int src, uint16_t* bgr) {
  if (!ctx || !bgr) return fail(DERP_EINVAL, "bad arguments");
  Ctx& c = ctx->c;
  int rc = checkProj(c, src, "derp_get_proj_color");
  if (rc) return rc;
  std::memcpy(bgr, c.projColor[src].data(), (size_t)c.W * c.H * 6);
  return DERP_OK;
}
int derp_get_proj_bias(DerpCtx* ctx, int src, uint16_t* bgr) {
  if (!ctx || !bgr) return fail(DERP_EINVAL, "bad arguments");
  Ctx& c = ctx->c;
  int rc = checkProj(c, src, "derp_get_proj_bias");
  if (rc) return rc;
  std::memcpy(bgr, c.projBias[src].data(), (size_t)c.W * c.H * 6);
  return DERP_OK;
}
int derp_get_counters(DerpCtx* ctx, uint64_t* cost_evals, uint64_t* src_hits) {
  if (!ctx) return fail(DERP_EINVAL, "null ctx");
  if (cost_evals) *cost_evals = ctx->c.costEvals.load();
  if (src_hits) *src_hits = ctx->c.srcHits.load();
  return DERP_OK;
}

// temporalJointBilateralFilterCol / temporalJointBilateralFilter (TemporalBilateralFilter.h:126-215)
int derp_temporal_filter(int device, int width, int height, int num_frames, const uint16_t* const* guides,
                         const float* const* disps, const uint8_t* const* masks, int frame_offset,
                         float sigma, int spatial_radius, float weight0, float weight1, float weight2,
                         float* out) {
  (void)device;
  if (!guides || !disps || !masks || !out || num_frames < 1 || frame_offset < 0 || frame_offset >= num_frames)
    return fail(DERP_EINVAL, "derp_temporal_filter: bad arguments");
  const int W = width, H = height;
  const float maxImageValue = 65535.0f;
  parallelFor(0, H, [&](int y0, int y1) {
    for (int y = y0; y < y1; ++y)
      for (int x = 0; x < W; ++x) {
        const size_t p = (size_t)y * W + x;
        if (!masks[frame_offset][p]) {
          out[p] = disps[frame_offset][p];
          continue;
        }
        float weightedSumPix = 0.0f, sumWeight = 0.0f;
        const uint16_t* referenceColor = guides[frame_offset] + p * 3;
        for (int t = 0; t < num_frames; ++t)
          for (int u = -spatial_radius; u <= spatial_radius; ++u)
            for (int v = -spatial_radius; v <= spatial_radius; ++v) {
              const int sampleX = clampi(x + u, 0, W - 1);
              const int sampleY = clampi(y + v, 0, H - 1);
              const size_t q = (size_t)sampleY * W + sampleX;
              if (!masks[t][q]) continue;
              const uint16_t* sampleColor = guides[t] + q * 3;
              // ushort - ushort -> int, then / float
              const float e0 = (referenceColor[0] - sampleColor[0]) / maxImageValue;
              const float e1 = (referenceColor[1] - sampleColor[1]) / maxImageValue;
              const float e2 = (referenceColor[2] - sampleColor[2]) / maxImageValue;
              const float weightedDiff = weight0 * (e0 * e0) + weight1 * (e1 * e1) + weight2 * (e2 * e2);
              const float weight = expf(-weightedDiff / (sigma * sigma));
              weightedSumPix += disps[t][p] * weight;
              sumWeight += weight;
            }
        out[p] = (weightedSumPix / sumWeight);
      }
  });
  return DERP_OK;
}

int derp_joint_bilateral_f32(int device, int width, int height, const float* image, const float* guide_bgr,
                             const uint8_t* mask, int radius, float sigma, float weight0, float weight1,
                             float weight2, float* out) {
  (void)device;
  if (!image || !guide_bgr || !mask || !out || radius < 0) return fail(DERP_EINVAL, "derp_joint_bilateral_f32: bad arguments");
  jointBilateral<float>(image, guide_bgr, 1.0f / 1.0f, mask, width, height, radius, sigma, weight0, weight1,
                        weight2, out);
  return DERP_OK;
}

// ---- oracle-only test hooks (not part of derp_b200.h) ----------------------------------------
int oracle_camera_project(const DerpCameraDesc* d, int normalized, const double* rig_pts, int n, double* pix,
                          uint8_t* sees) {
  Camera c;
  if (!c.init(*d)) return fail(DERP_EINVAL, "invalid camera");
  if (normalized) c.normalize();
  for (int i = 0; i < n; ++i) {
    double p[2] = {NAN, NAN};
    c.pixel(rig_pts + 3 * i, p);
    double q[2];
    sees[i] = c.sees(rig_pts + 3 * i, q) ? 1 : 0;
    pix[2 * i] = p[0];
    pix[2 * i + 1] = p[1];
  }
  return DERP_OK;
}
int oracle_camera_unproject(const DerpCameraDesc* d, const double* pix, int n, double depth, double* rig_pts,
                            uint8_t* outside_circle) {
  Camera c;
  if (!c.init(*d)) return fail(DERP_EINVAL, "invalid camera");
  for (int i = 0; i < n; ++i) {
    c.rig(pix + 2 * i, depth, rig_pts + 3 * i);
    outside_circle[i] = c.isOutsideImageCircle(pix + 2 * i) ? 1 : 0;
  }
  return DERP_OK;
}
int oracle_camera_info(const DerpCameraDesc* d, double* rotation9, double* distortion_max, double* cos_fov) {
  Camera c;
  if (!c.init(*d)) return fail(DERP_EINVAL, "invalid camera");
  for (int i = 0; i < 9; ++i) rotation9[i] = c.rotation[i / 3][i % 3];
  *distortion_max = c.distortionMax;
  *cos_fov = c.cosFov;
  return DERP_OK;
}
int oracle_camera_distort(const DerpCameraDesc* d, const double* r, int n, double* distorted, double* undistorted) {
  Camera c;
  if (!c.init(*d)) return fail(DERP_EINVAL, "invalid camera");
  for (int i = 0; i < n; ++i) {
    distorted[i] = c.distort(r[i]);
    undistorted[i] = c.undistort(r[i]);
  }
  return DERP_OK;
}
void oracle_remap_bicubic(const uint16_t* src, int sw, int sh, const float* map, int dw, int dh, uint16_t* dst) {
  remapBicubicU16C3(src, sw, sh, map, dw, dh, dst);
}
void oracle_blur3(const uint16_t* src, int w, int h, uint16_t* dst) { blur3x3U16C3(src, w, h, dst); }
int derp_downscale_area(int /*device*/, const uint16_t* src, int src_w, int src_h, uint16_t* dst, int dst_w, int dst_h) {
  if (!src || !dst || !resizeAreaU16C3(src, src_w, src_h, dst, dst_w, dst_h)) return DERP_EINVAL;
  return DERP_OK;
}
int derp_device_alloc(int /*device*/, size_t bytes, void** out) {
  if (!out) return DERP_EINVAL;
  *out = std::malloc(bytes ? bytes : 1);
  return *out ? DERP_OK : DERP_ENOMEM;
}
int derp_device_free(int /*device*/, void* p) {
  std::free(p);
  return DERP_OK;
}
int derp_device_copy(int /*device*/, void* dst, const void* src, size_t bytes) {
  if (!dst || !src) return DERP_EINVAL;
  std::memmove(dst, src, bytes);
  return DERP_OK;
}
int derp_foreground_mask(int /*device*/, const uint16_t* templ, const uint16_t* frame, int w, int h, int blur_radius,
                         float threshold, int morph_closing_size, uint8_t* mask) {
  if (!templ || !frame || !mask || w < 1 || h < 1 ||
      !foregroundMaskU16C3(templ, frame, w, h, blur_radius, threshold, morph_closing_size, mask))
    return DERP_EINVAL;
  return DERP_OK;
}
// ---- camera mesh (first slice): ConvertToBinary.cpp:150-183 before simplification -----------
namespace mesh {
// getTriangleMask, source/render/MeshUtil.h:162-221 (isRigCoordinates = false: z is the distance measure)
static unsigned triangleMask(const std::vector<double>& z, int base, int width, float tearRatio) {
  const double tl = z[base], tr = z[base + 1], bl = z[base + width], br = z[base + width + 1];
  std::vector<std::tuple<double, int>> v = {std::make_tuple(tl, 0), std::make_tuple(tr, 1), std::make_tuple(bl, 2),
                                            std::make_tuple(br, 3)};
  std::sort(v.begin(), v.end());  // literal: with NaN depths the order is whatever libstdc++'s insertion sort leaves
  if (std::get<0>(v.front()) / std::get<0>(v.back()) > tearRatio) {
    if (std::abs(tl - br) < std::abs(tr - bl)) return 1 << 1 | 1 << 2;
    return 1 << 0 | 1 << 3;
  }
  const double lo = std::get<0>(v.front()) / std::get<0>(v[2]);
  const double hi = std::get<0>(v[1]) / std::get<0>(v.back());
  if (lo >= tearRatio && lo > hi) return 1 << (std::get<1>(v.back()) ^ 0x3);
  if (hi >= tearRatio) return 1 << (std::get<1>(v.front()) ^ 0x3);
  return 0;
}
// addTriangle, MeshUtil.h:224-251
static void triangle(int which, int base, int width, int* f) {
  switch (which) {
    case 0: f[0] = base + width, f[1] = base + 1, f[2] = base; break;
    case 1: f[0] = base, f[1] = base + width + 1, f[2] = base + 1; break;
    case 2: f[0] = base + width + 1, f[1] = base, f[2] = base + width; break;
    default: f[0] = base + 1, f[1] = base + width, f[2] = base + width + 1; break;
  }
}
}  // namespace mesh

int derp_camera_mesh_size(int width, int height, double depth_scale, int* mesh_width, int* mesh_height) {
  if (width < 1 || height < 1 || !(depth_scale > 0) || depth_scale > 1 || !mesh_width || !mesh_height) return DERP_EINVAL;
  *mesh_width = depth_scale < 1 ? cvRoundD(width * depth_scale) : width;
  *mesh_height = depth_scale < 1 ? cvRoundD(height * depth_scale) : height;
  return DERP_OK;
}

int derp_camera_mesh(int /*device*/, const float* disparity, int width, int height, double depth_scale, double resolution_x,
                     double resolution_y, double scalar_focal, float tear_ratio, const uint8_t* foreground_mask,
                     int mask_width, int mask_height, float* vertexes, uint32_t* faces, uint64_t* num_vertexes,
                     uint64_t* num_faces) {
  int W = 0, H = 0;
  if (derp_camera_mesh_size(width, height, depth_scale, &W, &H) || !disparity || !vertexes || !faces || !num_vertexes ||
      !num_faces || W < 1 || H < 1 || (foreground_mask && (mask_width < 1 || mask_height < 1)))
    return DERP_EINVAL;
  // depth = 1.0f / disparity (cv::divide on floats: IEEE since OpenCV 4), ConvertToBinary.cpp:151-152
  std::vector<float> depth((size_t)width * height);
  for (size_t i = 0; i < depth.size(); ++i) depth[i] = 1.0f / disparity[i];
  if (depth_scale < 1) {  // :153-156
    std::vector<float> small;
    int w2, h2;
    resizeNearestScaled<float>(depth.data(), width, height, depth_scale, depth_scale, small, &w2, &h2);
    depth.swap(small);
  }
  const size_t n = (size_t)W * H;
  // getVertexesEquiError, MeshUtil.h:313-338
  const double scale = scalar_focal * 1.0;
  std::vector<double> vx(n), vy(n), vz(n);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const size_t i = (size_t)y * W + x;
      vx[i] = resolution_x / W * (x + 0.5);
      vy[i] = resolution_y / H * (y + 0.5);
      vz[i] = scale / depth[i];
    }
  // getFaces(wrapHorizontally = false, isRigCoordinates = false), MeshUtil.h:267-298
  std::vector<int> fc;
  fc.reserve(n * 6);
  for (int y = 0; y < H - 1; ++y)
    for (int x = 0; x < W - 1; ++x) {
      const int base = y * W + x;
      const unsigned m = mesh::triangleMask(vz, base, W, tear_ratio);
      for (int t = 0; t < 4; ++t)
        if ((m >> t) & 1) {
          int f[3];
          mesh::triangle(t, base, W, f);
          fc.insert(fc.end(), f, f + 3);
        }
    }
  // vertex mask, ConvertToBinary.cpp:163-176
  std::vector<uint8_t> vmask(n);
  for (size_t i = 0; i < n; ++i) vmask[i] = !std::isnan(depth[i]);
  if (foreground_mask) {
    std::vector<uint8_t> fm(n);
    resizeNearest<uint8_t>(foreground_mask, mask_width, mask_height, fm.data(), W, H);
    for (size_t i = 0; i < n; ++i) vmask[i] = vmask[i] & fm[i];  // Mat_<bool> & Mat_<bool>: bitwise on the bytes
  }
  // applyMaskToVertexesAndFaces, MeshUtil.h:342-403
  std::vector<size_t> keep;
  for (size_t i = 0; i < fc.size() / 3; ++i)
    if (vmask[fc[3 * i]] && vmask[fc[3 * i + 1]] && vmask[fc[3 * i + 2]]) keep.push_back(i);
  std::vector<uint8_t> usedv(n, 0);
  for (size_t i : keep)
    for (int j = 0; j < 3; ++j) usedv[fc[3 * i + j]] = 1;
  std::vector<int> newIndex(n, -1);
  uint64_t nv = 0;
  for (size_t i = 0; i < n; ++i)
    if (usedv[i]) {
      newIndex[i] = (int)nv;
      // writeDepth's vertexes.cast<float>(), MeshUtil.h:80-83
      vertexes[nv * 3 + 0] = (float)vx[i];
      vertexes[nv * 3 + 1] = (float)vy[i];
      vertexes[nv * 3 + 2] = (float)vz[i];
      ++nv;
    }
  for (size_t k = 0; k < keep.size(); ++k)
    for (int j = 0; j < 3; ++j) faces[k * 3 + j] = (uint32_t)newIndex[fc[3 * keep[k] + j]];
  *num_vertexes = nv;
  *num_faces = keep.size();
  return DERP_OK;
}

// The simplifier is checked against the reference's own MeshSimplifier.cpp (oracle/_ref) only: no second restatement here.
int derp_camera_mesh_simplified(int, const float*, int, int, double, double, double, double, float, const uint8_t*, int, int,
                                int, float*, uint32_t*, uint64_t*, uint64_t*) {
  g_err = "derp_camera_mesh_simplified: the oracle restatement stops before simplification; use oracle/_ref";
  return DERP_EINVAL;
}

// BC7 colour is checked against the reference's own encoder (oracle/_ref: kernel.ispc compiled by the ispc binary the
// reference vendors) and, for arithmetic identity, oracle/libbc7_x86.so: no second restatement here.
int derp_bc7_compress(int, const uint8_t*, int, int, uint8_t*) {
  g_err = "derp_bc7_compress: no oracle restatement; the checker is oracle/_ref (the reference's own encoder)";
  return DERP_EINVAL;
}
int derp_bc7_compress_image(int, const void*, int, int, int, int, float, uint8_t*) {
  g_err = "derp_bc7_compress_image: no oracle restatement; the checker is oracle/_ref (the reference's own encoder)";
  return DERP_EINVAL;
}

int oracle_resize_area(const uint16_t* src, int sw, int sh, uint16_t* dst, int dw, int dh) {
  return resizeAreaU16C3(src, sw, sh, dst, dw, dh) ? 0 : -1;
}
void oracle_variance(const uint16_t* src, int w, int h, float* var) { imageVarianceU16C3(src, w, h, var); }
void oracle_lanczos4(const float* src, int sw, int sh, float* dst, int dw, int dh) {
  resizeLanczos4F32(src, sw, sh, dst, dw, dh);
}
void oracle_nearest_f32(const float* src, int sw, int sh, float* dst, int dw, int dh) {
  resizeNearest<float>(src, sw, sh, dst, dw, dh);
}
// libstdc++ reference streams for the GPU's skip-ahead RNG and introselect emulation tests
void oracle_minstd_uniform(uint32_t seed, int n, float a, float b, float* out) {
  std::default_random_engine engine;
  engine.seed(seed);
  for (int i = 0; i < n; ++i) out[i] = std::uniform_real_distribution<float>(a, b)(engine);
}
float oracle_nth_element_sum(const float* first, const float* second, int count, int keep) {
  std::pair<float, float> v[64];
  for (int i = 0; i < count; ++i) v[i] = {first[i], second[i]};
  std::nth_element(v, v + keep, v + count);
  float cost = 0;
  for (int i = 0; i < keep; ++i) cost += v[i].second;
  return cost;
}

}  // extern "C"
