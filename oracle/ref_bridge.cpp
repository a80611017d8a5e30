// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// oracle/_ref/libderp_ref.so = the REFERENCE'S OWN depth-path sources, compiled where they lie under
// /root/reference (never copied into this repository), behind the same C ABI as the product and the oracle
// (include/derp_b200.h), so that a parity test is "same calls, three libraries, compare buffers":
//
//   /root/reference/source/depth_estimation/Derp.cpp               computeCost, brute force, proposals, ping-pong,
//                                                                  mismatches, bilateral / median / maskFov, processLevel
//   /root/reference/source/depth_estimation/DerpUtil.cpp           computeSSD, dstToWorldPoint, worldToSrcPoint, project,
//                                                                  colorBias, computeImageVariance, generateFovMasks
//   /root/reference/source/depth_estimation/UpsampleDisparityLib.cpp   upsampleDisparities
//   /root/reference/source/depth_estimation/{PyramidLevel,TemporalBilateralFilter}.h
//   /root/reference/source/util/Camera.cpp + Camera.h              the camera model and rig JSON
//   /root/reference/source/util/CvUtil.cpp + CvUtil.h              getPixelBilinear, maskedMedianBlur, PFM writer
//   /root/reference/source/util/ImageUtil.cpp                      probeDisparity, computeWarpDstToSrc
//
// The third-party libraries those sources include (OpenCV, Eigen, glog, gflags, fmt, folly, Boost) are not in
// this image; oracle/refshim/ holds minimal stand-ins with the same interfaces (each header says what it
// restates and what pins it).  This file is only glue: it builds a PyramidLevel<cv::Vec3w> from the ABI's
// buffers the way DerpCLI.cpp:229-271 does and calls the reference's functions.  Per-destination entry points
// run the reference function on a one-destination VIEW of the level (a copy of the PyramidLevel struct whose
// cv::Mat members share their pixels with the full level).
//
// kNumDepths is a compile-time constant of the reference (Derp.h:33): brute force accepts num_depths == 150 only.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <unistd.h>

#include "source/depth_estimation/Derp.h"
#include "source/depth_estimation/TemporalBilateralFilter.h"
#include "source/depth_estimation/UpsampleDisparityLib.h"
#include "source/render/MeshSimplifier.h"
#include "source/render/MeshUtil.h"
#include "source/util/Camera.h"

#include "../include/derp_b200.h"

using namespace fb360_dep;
using namespace fb360_dep::depth_estimation;
using Level = PyramidLevel<PixelType>;

// functions of Derp.cpp with external linkage that Derp.h does not declare
namespace fb360_dep::depth_estimation {
void randomProposal(Level& pyramidLevel, const int dstIdx, const int y, const int numProposals,
                    const float minDepthMeters, const float maxDepthMeters);
void pingPong(Level& pyramidLevel, const int iterations, const int numThreads);
}  // namespace fb360_dep::depth_estimation

namespace {

thread_local std::string g_err;
int g_threads = -1;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

std::string num17(double v) {
  char b[64];
  snprintf(b, sizeof b, "%.17g", v);
  return b;
}
std::string vecJson(const double* v, int n) {
  std::string s = "[";
  for (int i = 0; i < n; ++i) s += (i ? "," : "") + num17(v[i]);
  return s + "]";
}

}  // namespace

struct DerpCtx {
  Camera::Rig rigSrc, rigDst;  // normalised (DerpCLI.cpp:216-218)
  std::vector<int> dst2src;
  int fullW = 0, fullH = 0;
  DerpLevelParams lp{};
  bool haveLevel = false;
  std::vector<cv::Mat_<bool>> fovMasks;
  std::vector<cv::Mat_<bool>> fgMasks;
  std::vector<cv::Mat_<float>> bgDisps;
  std::unique_ptr<Level> level;
  bool projected = false;
  int currentDst = -1;
  std::string tmpDir;
  std::vector<cv::Mat_<float>> kept;  // derp_level_keep
  // values the caller set before the level object existed
  std::vector<cv::Mat_<float>> pendDisp, pendCost, pendConf;
};

namespace {

Level viewOf(const Level& full, int dst) {
  Level v = full;  // cv::Mat members share their pixels
  const int S = (int)full.rigSrc.size();
  v.rigDst = {full.rigDst[dst]};
  v.dsts = {full.dsts[dst]};
  v.dst2srcIdxs = {full.dst2srcIdxs[dst]};
  v.projs.assign(full.projs.begin() + (size_t)dst * S, full.projs.begin() + (size_t)(dst + 1) * S);
  return v;
}

template <class F>
int guarded(F&& f) {
  try {
    return f();
  } catch (const refshim::Fatal& e) {
    const std::string m = e.what();
    return fail(m.find("Insufficient coverage") != std::string::npos ? DERP_ECOVERAGE : DERP_EINVAL, m);
  } catch (const std::exception& e) {
    return fail(DERP_EINVAL, e.what());
  }
}

int needLevel(DerpCtx* c) {
  if (!c || !c->level) return fail(DERP_ESTATE, "derp_set_colors has not been called for this level");
  return DERP_OK;
}
int needProjected(DerpCtx* c, int dst) {
  if (int rc = needLevel(c)) return rc;
  if (dst < 0 || dst >= (int)c->rigDst.size()) return fail(DERP_EINVAL, "bad destination index");
  if (!c->projected) return fail(DERP_ESTATE, "derp_reproject has not been called for this level");
  return DERP_OK;
}

template <class T>
cv::Mat_<T> wrapCopy(const T* p, int w, int h) {
  cv::Mat_<T> m(h, w);
  std::memcpy(m.data, p, (size_t)w * h * sizeof(T));
  return m;
}

}  // namespace

extern "C" {

const char* derp_backend(void) { return "reference-cpu"; }
const char* derp_last_error(void) { return g_err.c_str(); }
int derp_set_threads(int threads) {
  g_threads = threads;
  return DERP_OK;
}

int derp_create(const DerpCameraDesc* cams, int num_cams, const int32_t* dst_to_src, int num_dsts, int /*device*/,
                DerpCtx** out) {
  return guarded([&] {
    if (!cams || num_cams < 1 || !dst_to_src || num_dsts < 1 || !out) return fail(DERP_EINVAL, "bad arguments");
    static const char* kTypes[] = {"FTHETA", "RECTILINEAR", "EQUISOLID", "ORTHOGRAPHIC"};
    std::string json = "{\"cameras\":[";
    for (int i = 0; i < num_cams; ++i) {
      const DerpCameraDesc& d = cams[i];
      if (d.type < 0 || d.type > 3) return fail(DERP_EINVAL, "bad camera type");
      json += std::string(i ? "," : "") + "{\"version\":1,\"type\":\"" + kTypes[d.type] + "\",\"id\":\"cam" + std::to_string(i) +
          "\",\"origin\":" + vecJson(d.origin, 3) + ",\"forward\":" + vecJson(d.forward, 3) + ",\"up\":" + vecJson(d.up, 3) +
          ",\"right\":" + vecJson(d.right, 3) + ",\"resolution\":" + vecJson(d.resolution, 2) + ",\"focal\":" + vecJson(d.focal, 2);
      if (d.has_principal) json += ",\"principal\":" + vecJson(d.principal, 2);
      json += ",\"distortion\":" + vecJson(d.distortion, 3);
      if (d.has_fov) json += ",\"fov\":" + num17(d.fov);
      json += "}";
    }
    json += "]}";
    auto c = std::make_unique<DerpCtx>();
    c->rigSrc = Camera::loadRigFromJsonString(json);
    for (int i = 0; i < num_dsts; ++i) {
      if (dst_to_src[i] < 0 || dst_to_src[i] >= num_cams) return fail(DERP_EINVAL, "bad dst_to_src entry");
      c->rigDst.push_back(c->rigSrc[dst_to_src[i]]);
    }
    c->dst2src = mapSrcToDstIndexes(c->rigSrc, c->rigDst);
    c->fullW = (int)c->rigDst[0].resolution.x();  // DerpCLI.cpp:212-214
    c->fullH = (int)c->rigDst[0].resolution.y();
    Camera::normalizeRig(c->rigSrc);
    Camera::normalizeRig(c->rigDst);
    *out = c.release();
    return DERP_OK;
  });
}

void derp_destroy(DerpCtx* c) {
  if (!c) return;
  if (!c->tmpDir.empty()) {
    std::error_code ec;
    std::filesystem::remove_all(c->tmpDir, ec);
  }
  delete c;
}

int derp_set_stream(DerpCtx*, void*) { return DERP_OK; }
int derp_sync(DerpCtx*) { return DERP_OK; }
int derp_get_launch_count(DerpCtx*, uint64_t* out) {
  if (out) *out = 0;
  return DERP_OK;
}
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

int derp_level_begin(DerpCtx* c, const DerpLevelParams* p) {
  return guarded([&] {
    if (!c || !p || p->width < 3 || p->height < 3) return fail(DERP_EINVAL, "bad level parameters");
    c->lp = *p;
    c->haveLevel = true;
    c->level.reset();
    c->projected = false;
    c->currentDst = -1;
    const cv::Size size(p->width, p->height);
    c->fovMasks = generateFovMasks(c->rigDst, size, g_threads);  // DerpCLI.cpp:226-227
    c->fgMasks = cv_util::generateAllPassMasks(size, (int)c->rigSrc.size());
    c->bgDisps.assign(c->rigDst.size(), cv::Mat_<float>());
    c->pendDisp.assign(c->rigDst.size(), cv::Mat_<float>());
    c->pendCost = c->pendDisp;
    c->pendConf = c->pendDisp;
    return DERP_OK;
  });
}

int derp_set_foreground_masks(DerpCtx* c, const uint8_t* const* masks) {
  return guarded([&] {
    if (!c || !c->haveLevel || !masks) return fail(DERP_ESTATE, "derp_level_begin first");
    if (c->level) return fail(DERP_ESTATE, "masks must be set before derp_set_colors");
    for (size_t s = 0; s < c->rigSrc.size(); ++s) {
      cv::Mat_<bool> m(c->lp.height, c->lp.width);
      for (size_t i = 0, n = m.total(); i < n; ++i) m.data[i] = masks[s][i] ? 1 : 0;
      c->fgMasks[s] = m;
    }
    return DERP_OK;
  });
}

int derp_set_background_disparity(DerpCtx* c, const float* const* background) {
  return guarded([&] {
    if (!c || !c->haveLevel || !background) return fail(DERP_ESTATE, "derp_level_begin first");
    if (c->level) return fail(DERP_ESTATE, "background must be set before derp_set_colors");
    for (size_t d = 0; d < c->rigDst.size(); ++d) c->bgDisps[d] = wrapCopy(background[d], c->lp.width, c->lp.height);
    return DERP_OK;
  });
}

int derp_set_colors(DerpCtx* c, const uint16_t* const* colors) {
  return guarded([&] {
    if (!c || !c->haveLevel || !colors) return fail(DERP_ESTATE, "derp_level_begin first");
    const int W = c->lp.width, H = c->lp.height;
    std::vector<cv::Mat_<PixelType>> imgs;
    for (size_t s = 0; s < c->rigSrc.size(); ++s) {
      cv::Mat_<PixelType> m(H, W);
      std::memcpy(m.data, colors[s], (size_t)W * H * 6);
      imgs.push_back(m);
    }
    std::map<int, cv::Size> sizes;
    sizes[c->lp.level] = cv::Size(W, H);
    if (c->tmpDir.empty()) {
      char tmpl[] = "/tmp/derp_ref_XXXXXX";
      if (!mkdtemp(tmpl)) return fail(DERP_EINVAL, "mkdtemp failed");
      c->tmpDir = tmpl;
    }
    c->level = std::make_unique<Level>(  // DerpCLI.cpp:250-271
        0, "000000", 1, c->lp.level, c->lp.num_levels, sizes, c->rigSrc, c->rigDst, c->dst2src, imgs, c->fgMasks,
        c->fovMasks, c->bgDisps, c->lp.full_width, c->lp.full_height, std::string(), c->lp.var_noise_floor,
        c->lp.var_high_thresh, c->lp.use_foreground_masks != 0, c->tmpDir, g_threads);
    for (size_t d = 0; d < c->rigDst.size(); ++d) {
      if (!c->pendDisp[d].empty()) c->pendDisp[d].copyTo(c->level->dstDisparity((int)d));
      if (!c->pendCost[d].empty()) c->pendCost[d].copyTo(c->level->dstCost((int)d));
      if (!c->pendConf[d].empty()) c->pendConf[d].copyTo(c->level->dstConfidence((int)d));
    }
    c->projected = false;
    return DERP_OK;
  });
}

int derp_reproject(DerpCtx* c, int dst) {
  return guarded([&] {
    if (int rc = needLevel(c)) return rc;
    if (dst < 0 || dst >= (int)c->rigDst.size()) return fail(DERP_EINVAL, "bad destination index");
    if (!c->projected) {  // the reference keeps every (dst, src) pair resident: built once per level
      precomputeProjections(*c->level, g_threads);
      reprojectColors(*c->level, g_threads);
      c->projected = true;
    }
    c->currentDst = dst;
    return DERP_OK;
  });
}

int derp_brute_force(DerpCtx* c, int dst, int num_depths, float min_depth_m, float max_depth_m, int partial_coverage,
                     int32_t* best_index) {
  return guarded([&] {
    if (int rc = needProjected(c, dst)) return rc;
    if (num_depths != kNumDepths) return fail(DERP_EINVAL, "the reference's candidate count is the compile-time constant 150 (Derp.h:33)");
    Level v = viewOf(*c->level, dst);
    computeBruteForceDisparity(v, 0, min_depth_m, max_depth_m, partial_coverage != 0, c->lp.use_foreground_masks != 0, g_threads);
    if (best_index) {
      // the reference does not keep the index; candidates are distinct, so it is recovered from the disparity
      std::vector<float> table(kNumDepths);
      const float minD = 1.0f / max_depth_m, maxD = 1.0f / min_depth_m;
      for (int i = 0; i < kNumDepths; ++i) table[i] = image_util::probeDisparity(i, kNumDepths, minD, maxD);
      const int W = c->lp.width, H = c->lp.height;
      const cv::Mat_<float>& disp = c->level->dstDisparity(dst);
      const cv::Mat_<float>& cost = c->level->dstCost(dst);
      const cv::Mat_<bool>& fov = c->level->dstFovMask(dst);
      const cv::Mat_<bool>& fg = c->level->dstForegroundMask(dst);
      auto interior = [&](int x, int y) -> int32_t {
        if (!fov(y, x)) return -2;
        if (!fg(y, x)) return -3;
        if (cost(y, x) == FLT_MAX) return -1;
        for (int i = 0; i < kNumDepths; ++i)
          if (table[i] == disp(y, x)) return i;
        return -4;
      };
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
          const bool border = x < 1 || x >= W - 1 || y < 1 || y >= H - 1;
          if (border && !fg(y, x)) {
            best_index[(size_t)y * W + x] = -3;
            continue;
          }
          best_index[(size_t)y * W + x] = interior(std::min(std::max(x, 1), W - 2), std::min(std::max(y, 1), H - 2));
        }
    }
    return DERP_OK;
  });
}

int derp_random_proposals(DerpCtx* c, int dst, int num_proposals, float min_depth_m, float max_depth_m) {
  return guarded([&] {
    if (int rc = needProjected(c, dst)) return rc;
    if (num_proposals <= 0) return DERP_OK;
    Level v = viewOf(*c->level, dst);
    // randomProposals (Derp.cpp:844-873) minus its level test: one task per row
    ThreadPool threadPool(g_threads);
    const cv::Size size = v.dstDisparity(0).size();
    for (int y = kSearchWindowRadius; y < size.height - kSearchWindowRadius; ++y)
      threadPool.spawn(&randomProposal, std::ref(v), 0, y, num_proposals, min_depth_m, max_depth_m);
    threadPool.join();
    return DERP_OK;
  });
}

int derp_ping_pong(DerpCtx* c, int dst, int iterations) {
  return guarded([&] {
    if (int rc = needProjected(c, dst)) return rc;
    Level v = viewOf(*c->level, dst);
    pingPong(v, iterations, g_threads);
    return DERP_OK;
  });
}

int derp_mismatches(DerpCtx* c) {
  return guarded([&] {
    if (int rc = needLevel(c)) return rc;
    Level v = *c->level;
    v.numLevels = v.level + 2;  // the ABI leaves the level test to the caller (processLevel applies it)
    handleDisparityMismatches(v, v.level, g_threads);
    return DERP_OK;
  });
}
const float* derp_disparity_device_ptr(DerpCtx* c, int dst) {
  if (!c || !c->level || dst < 0 || dst >= (int)c->rigDst.size()) return nullptr;
  return c->level->dstDisparity(dst).ptr<float>();
}
int derp_gather_disparities(DerpCtx*, const float* const*) { return fail(DERP_ESTATE, "camera sharding is not a reference feature"); }
int derp_mismatches_gathered(DerpCtx*) { return fail(DERP_ESTATE, "camera sharding is not a reference feature"); }

int derp_bilateral(DerpCtx* c, int dst) {
  return guarded([&] {
    if (int rc = needLevel(c)) return rc;
    Level v = viewOf(*c->level, dst);
    bilateralFilter(v, g_threads);
    return DERP_OK;
  });
}
int derp_median(DerpCtx* c, int dst) {
  return guarded([&] {
    if (int rc = needLevel(c)) return rc;
    Level v = viewOf(*c->level, dst);
    medianFilter(v, g_threads);
    return DERP_OK;
  });
}
int derp_mask_fov(DerpCtx* c, int dst) {
  return guarded([&] {
    if (int rc = needLevel(c)) return rc;
    Level v = viewOf(*c->level, dst);
    maskFov(v, g_threads);
    return DERP_OK;
  });
}

int derp_upsample_from(DerpCtx* c, int dst, const float* coarse, int coarse_w, int coarse_h, const uint8_t* coarse_mask,
                       const uint8_t* fine_mask) {
  return guarded([&] {
    if (int rc = needLevel(c)) return rc;
    if (dst < 0 || dst >= (int)c->rigDst.size() || !coarse) return fail(DERP_EINVAL, "bad arguments");
    const bool useFg = c->lp.use_foreground_masks != 0;
    std::vector<cv::Mat_<bool>> cm(1), fm(1);  // DerpCLI.cpp:277-285: empty unless foreground masks are used
    if (useFg) {
      if (!coarse_mask || !fine_mask) return fail(DERP_EINVAL, "foreground masks required");
      cm[0] = cv::Mat_<bool>(coarse_h, coarse_w);
      for (size_t i = 0, n = cm[0].total(); i < n; ++i) cm[0].data[i] = coarse_mask[i] ? 1 : 0;
      fm[0] = cv::Mat_<bool>(c->lp.height, c->lp.width);
      for (size_t i = 0, n = fm[0].total(); i < n; ++i) fm[0].data[i] = fine_mask[i] ? 1 : 0;
    }
    const std::vector<cv::Mat_<float>> up = upsampleDisparities(  // DerpCLI.cpp:290-298
        {c->rigDst[dst]}, {wrapCopy(coarse, coarse_w, coarse_h)}, {c->level->dstBackgroundDisparity(dst)}, cm, fm,
        cv::Size(c->lp.width, c->lp.height), useFg, g_threads);
    c->level->dsts[dst].disparity = up[0];  // DerpCLI.cpp:300-302
    return DERP_OK;
  });
}

int derp_level_keep(DerpCtx* c) {
  if (int rc = needLevel(c)) return rc;
  c->kept.clear();
  for (size_t d = 0; d < c->rigDst.size(); ++d) c->kept.push_back(c->level->dstDisparity((int)d).clone());
  return DERP_OK;
}
int derp_upsample_from_kept(DerpCtx* c, int dst, const uint8_t* coarse_mask, const uint8_t* fine_mask) {
  if (!c || dst < 0 || dst >= (int)c->kept.size()) return fail(DERP_ESTATE, "nothing kept");
  const cv::Mat_<float> k = c->kept[dst];
  return derp_upsample_from(c, dst, k.ptr<float>(), k.cols, k.rows, coarse_mask, fine_mask);
}

static int runHalves(DerpCtx* c, const DerpProcessOpts* o, bool estimate, bool mismatch, bool filter, bool save) {
  return guarded([&] {
    if (int rc = needLevel(c)) return rc;
    if (!o) return fail(DERP_EINVAL, "null options");
    Level& L = *c->level;
    const bool useFg = c->lp.use_foreground_masks != 0;
    if (estimate) {
      if (L.level == L.numLevels - 1 && o->num_depths != kNumDepths)
        return fail(DERP_EINVAL, "the reference's candidate count is the compile-time constant 150 (Derp.h:33)");
      if (!c->projected) precomputeProjections(L, g_threads);  // DerpCLI.cpp:274
      c->projected = true;
      reprojectColors(L, g_threads);  // processLevel, Derp.cpp:1020-1034
      preprocessLevel(L, o->min_depth_m, o->max_depth_m, o->partial_coverage != 0, useFg, g_threads);
      randomProposals(L, o->random_proposals, o->min_depth_m, o->max_depth_m, g_threads, "");
      pingPongPropagation(L, o->ping_pong_iterations, g_threads, "");
    }
    if (mismatch) handleDisparityMismatches(L, o->mismatches_start_level, g_threads);
    if (filter) {
      if (o->do_bilateral_filter) bilateralFilter(L, g_threads);
      if (o->do_median_filter) medianFilter(L, g_threads);
      maskFov(L, g_threads);
    }
    if (save) saveResults(L, false, "pfm");  // the reference's own PFM writer into the context's scratch directory
    return DERP_OK;
  });
}
int derp_process_level(DerpCtx* c, const DerpProcessOpts* o) { return runHalves(c, o, true, true, true, true); }
int derp_level_estimate(DerpCtx* c, const DerpProcessOpts* o) { return runHalves(c, o, true, false, false, false); }
int derp_level_filter(DerpCtx* c, const DerpProcessOpts* o) { return runHalves(c, o, false, false, true, false); }

int derp_eval_cost(DerpCtx* c, int dst, const float* disparity, float* out_cost, float* out_conf) {
  return guarded([&] {
    if (int rc = needProjected(c, dst)) return rc;
    const int W = c->lp.width, H = c->lp.height;
    const float nan = std::numeric_limits<float>::quiet_NaN();
    for (size_t i = 0; i < (size_t)W * H; ++i) out_cost[i] = out_conf[i] = nan;
    const Level& L = *c->level;
    int nt = g_threads < 0 ? (int)std::max(1u, std::thread::hardware_concurrency()) : std::max(1, g_threads);
    std::atomic<int> next(1);
    std::vector<std::thread> pool;
    for (int t = 0; t < nt; ++t)
      pool.emplace_back([&] {
        for (int y = next++; y < H - 1; y = next++)
          for (int x = 1; x < W - 1; ++x) {
            const auto r = computeCost(L, dst, disparity[(size_t)y * W + x], x, y);
            out_cost[(size_t)y * W + x] = std::get<0>(r);
            out_conf[(size_t)y * W + x] = std::get<1>(r);
          }
      });
    for (auto& t : pool) t.join();
    return DERP_OK;
  });
}

int derp_set_disparity(DerpCtx* c, int dst, const float* disparity, const float* cost, const float* confidence) {
  return guarded([&] {
    if (!c || !c->haveLevel || dst < 0 || dst >= (int)c->rigDst.size()) return fail(DERP_EINVAL, "bad arguments");
    const int W = c->lp.width, H = c->lp.height;
    auto put = [&](const float* p, cv::Mat_<float>& pend, cv::Mat_<float>* live) {
      if (!p) return;
      if (live) std::memcpy(live->data, p, (size_t)W * H * 4);
      else pend = wrapCopy(p, W, H);
    };
    put(disparity, c->pendDisp[dst], c->level ? &c->level->dstDisparity(dst) : nullptr);
    put(cost, c->pendCost[dst], c->level ? &c->level->dstCost(dst) : nullptr);
    put(confidence, c->pendConf[dst], c->level ? &c->level->dstConfidence(dst) : nullptr);
    return DERP_OK;
  });
}
int derp_get_disparity(DerpCtx* c, int dst, float* disparity, float* cost, float* confidence) {
  if (int rc = needLevel(c)) return rc;
  if (dst < 0 || dst >= (int)c->rigDst.size()) return fail(DERP_EINVAL, "bad destination index");
  const size_t bytes = (size_t)c->lp.width * c->lp.height * 4;
  if (disparity) std::memcpy(disparity, c->level->dstDisparity(dst).data, bytes);
  if (cost) std::memcpy(cost, c->level->dstCost(dst).data, bytes);
  if (confidence) std::memcpy(confidence, c->level->dstConfidence(dst).data, bytes);
  return DERP_OK;
}
int derp_get_fov_mask(DerpCtx* c, int dst, uint8_t* mask) {
  if (!c || !c->haveLevel || dst < 0 || dst >= (int)c->rigDst.size()) return fail(DERP_EINVAL, "bad arguments");
  const cv::Mat_<bool>& m = c->fovMasks[dst];
  for (size_t i = 0, n = m.total(); i < n; ++i) mask[i] = m.data[i] ? 1 : 0;
  return DERP_OK;
}
int derp_get_mismatch_mask(DerpCtx* c, int dst, uint8_t* mask) {
  if (int rc = needLevel(c)) return rc;
  const cv::Mat_<bool>& m = c->level->dstMismatchedDisparityMask(dst);
  for (size_t i = 0, n = m.total(); i < n; ++i) mask[i] = m.data[i] ? 1 : 0;
  return DERP_OK;
}
int derp_get_variance(DerpCtx* c, int src, float* variance) {
  if (int rc = needLevel(c)) return rc;
  std::memcpy(variance, c->level->srcVariance(src).data, (size_t)c->lp.width * c->lp.height * 4);
  return DERP_OK;
}
int derp_get_var_noise_floor(DerpCtx* c, float* out) {
  if (int rc = needLevel(c)) return rc;
  *out = c->level->varNoiseFloor;
  return DERP_OK;
}
int derp_get_proj_warp(DerpCtx* c, int src, float* warp_xy) {
  if (int rc = needProjected(c, c ? c->currentDst : -1)) return rc;
  std::memcpy(warp_xy, c->level->dstProjWarp(c->currentDst, src).data, (size_t)c->lp.width * c->lp.height * 8);
  return DERP_OK;
}
int derp_get_proj_color(DerpCtx* c, int src, uint16_t* bgr) {
  if (int rc = needProjected(c, c ? c->currentDst : -1)) return rc;
  std::memcpy(bgr, c->level->dstProjColor(c->currentDst, src).data, (size_t)c->lp.width * c->lp.height * 6);
  return DERP_OK;
}
int derp_get_proj_bias(DerpCtx* c, int src, uint16_t* bgr) {
  if (int rc = needProjected(c, c ? c->currentDst : -1)) return rc;
  std::memcpy(bgr, c->level->dstProjColorBias(c->currentDst, src).data, (size_t)c->lp.width * c->lp.height * 6);
  return DERP_OK;
}
int derp_get_counters(DerpCtx*, uint64_t* a, uint64_t* b) {
  if (a) *a = 0;
  if (b) *b = 0;
  return DERP_OK;
}

int derp_temporal_filter(int /*device*/, int width, int height, int num_frames, const uint16_t* const* guides,
                         const float* const* disps, const uint8_t* const* masks, int frame_offset, float sigma,
                         int spatial_radius, float weight0, float weight1, float weight2, float* out) {
  return guarded([&] {
    std::vector<cv::Mat_<PixelType>> g;
    std::vector<cv::Mat_<float>> d;
    std::vector<cv::Mat_<bool>> m;
    for (int t = 0; t < num_frames; ++t) {
      cv::Mat_<PixelType> gi(height, width);
      std::memcpy(gi.data, guides[t], (size_t)width * height * 6);
      g.push_back(gi);
      d.push_back(wrapCopy(disps[t], width, height));
      cv::Mat_<bool> mi(height, width);
      for (size_t i = 0, n = mi.total(); i < n; ++i) mi.data[i] = masks[t][i] ? 1 : 0;
      m.push_back(mi);
    }
    cv::Mat_<float> result;
    temporalJointBilateralFilter<PixelType>(g, d, m, frame_offset, sigma, spatial_radius, weight0, weight1, weight2, result, g_threads);
    std::memcpy(out, result.data, (size_t)width * height * 4);
    return DERP_OK;
  });
}

int derp_joint_bilateral_f32(int /*device*/, int width, int height, const float* image, const float* guide_bgr,
                             const uint8_t* mask, int radius, float sigma, float weight0, float weight1, float weight2,
                             float* out) {
  return guarded([&] {
    cv::Mat_<cv::Vec3f> g(height, width);
    std::memcpy(g.data, guide_bgr, (size_t)width * height * 12);
    cv::Mat_<bool> m(height, width);
    for (size_t i = 0, n = m.total(); i < n; ++i) m.data[i] = mask[i] ? 1 : 0;
    const cv::Mat_<float> r = generalizedJointBilateralFilter<float, cv::Vec3f>(
        wrapCopy(image, width, height), g, g, m, radius, sigma, weight0, weight1, weight2, g_threads);
    std::memcpy(out, r.data, (size_t)width * height * 4);
    return DERP_OK;
  });
}

int derp_upsample_disparity(int /*device*/, const DerpCameraDesc* cam, const float* coarse, int coarse_w, int coarse_h,
                            const float* background_up, const uint8_t* coarse_mask, const uint8_t* fine_mask, int out_w,
                            int out_h, int use_foreground_masks, float* out) {
  DerpCtx* c = nullptr;
  const int32_t d2s = 0;
  if (int rc = derp_create(cam, 1, &d2s, 1, 0, &c)) return rc;
  const int rc = guarded([&] {
    std::vector<cv::Mat_<bool>> cm(1), fm(1);
    std::vector<cv::Mat_<float>> bg(1);
    if (use_foreground_masks) {
      cm[0] = cv::Mat_<bool>(coarse_h, coarse_w);
      for (size_t i = 0, n = cm[0].total(); i < n; ++i) cm[0].data[i] = coarse_mask[i] ? 1 : 0;
      fm[0] = cv::Mat_<bool>(out_h, out_w);
      for (size_t i = 0, n = fm[0].total(); i < n; ++i) fm[0].data[i] = fine_mask[i] ? 1 : 0;
      if (background_up) bg[0] = wrapCopy(background_up, out_w, out_h);
    }
    const std::vector<cv::Mat_<float>> up = upsampleDisparities(c->rigDst, {wrapCopy(coarse, coarse_w, coarse_h)}, bg, cm, fm,
                                                                cv::Size(out_w, out_h), use_foreground_masks != 0, g_threads);
    std::memcpy(out, up[0].data, (size_t)out_w * out_h * 4);
    return DERP_OK;
  });
  derp_destroy(c);
  return rc;
}

/* scripts/render/resize.py is Python + cv2, not part of the compiled reference: the stand-in's INTER_AREA
 * (oracle/cvprims.h, pinned to cv2 4.13) answers for it so that the library exports the whole ABI */
int derp_downscale_area(int /*device*/, const uint16_t* src, int src_w, int src_h, uint16_t* dst, int dst_w, int dst_h) {
  if (!src || !dst || !oracle::resizeAreaU16C3(src, src_w, src_h, dst, dst_w, dst_h)) return fail(DERP_EINVAL, "bad arguments");
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
/* source/render/BackgroundSubtractionUtil.h is outside the sources compiled here (it needs cv::GaussianBlur /
 * morphologyEx): the stand-in's restatement (oracle/cvprims.h, pinned to cv2 4.13) answers for it */
int derp_foreground_mask(int /*device*/, const uint16_t* templ, const uint16_t* frame, int w, int h, int blur_radius,
                         float threshold, int morph_closing_size, uint8_t* mask) {
  if (!templ || !frame || !mask || !oracle::foregroundMaskU16C3(templ, frame, w, h, blur_radius, threshold, morph_closing_size, mask))
    return fail(DERP_EINVAL, "bad arguments");
  return DERP_OK;
}

/* Camera mesh: the reference's own mesh_util functions (source/render/MeshUtil.h, compiled unmodified) called in the order
 * convertDepth calls them (source/mesh_stream/ConvertToBinary.cpp:150-183; that file itself is an executable with BC7 and
 * MeshSimplifier dependencies, so its dozen lines of glue are restated here), up to the casts of mesh_util::writeDepth. */
int derp_camera_mesh_size(int width, int height, double depth_scale, int* mesh_width, int* mesh_height) {
  if (width < 1 || height < 1 || !(depth_scale > 0) || depth_scale > 1 || !mesh_width || !mesh_height)
    return fail(DERP_EINVAL, "bad arguments");
  *mesh_width = depth_scale < 1 ? oracle::cvRoundD(width * depth_scale) : width;
  *mesh_height = depth_scale < 1 ? oracle::cvRoundD(height * depth_scale) : height;
  return DERP_OK;
}

static int cameraMeshRef(const float* disparity, int width, int height, double depth_scale, double resolution_x,
                         double resolution_y, double scalar_focal, float tear_ratio, const uint8_t* foreground_mask,
                         int mask_width, int mask_height, int triangles, float* vertexesOut, uint32_t* facesOut,
                         uint64_t* num_vertexes, uint64_t* num_faces) {
  int W = 0, H = 0;
  if (derp_camera_mesh_size(width, height, depth_scale, &W, &H) || !disparity || !vertexesOut || !facesOut ||
      !num_vertexes || !num_faces || W < 1 || H < 1 || (foreground_mask && (mask_width < 1 || mask_height < 1)))
    return fail(DERP_EINVAL, "bad arguments");
  return guarded([&] {
    cv::Mat_<float> depth(height, width);
    for (int y = 0; y < height; ++y)
      for (int x = 0; x < width; ++x) depth(y, x) = 1.0f / disparity[(size_t)y * width + x];  // cv::Mat 1.0f / disparity
    if (depth_scale < 1) cv::resize(depth, depth, cv::Size(), depth_scale, depth_scale, cv::INTER_NEAREST);
    const Camera cam(Camera::Type::FTHETA, Camera::Vector2(resolution_x, resolution_y),
                     Camera::Vector2(scalar_focal, -scalar_focal));
    Eigen::MatrixXd vertexes = mesh_util::getVertexesEquiError(depth, cam);
    Eigen::MatrixXi faces = mesh_util::getFaces(vertexes, depth.cols, depth.rows, false, false, tear_ratio);
    cv::Mat_<bool> vertexMask(depth.size());
    for (int i = 0; i < depth.rows; ++i)
      for (int j = 0; j < depth.cols; ++j) vertexMask(i, j) = !std::isnan(depth(i, j));
    if (foreground_mask) {
      cv::Mat_<bool> foregroundMask(mask_height, mask_width);
      std::memcpy(foregroundMask.data, foreground_mask, (size_t)mask_width * mask_height);
      cv::resize(foregroundMask, foregroundMask, depth.size(), 0, 0, cv::INTER_NEAREST);
      vertexMask = vertexMask & foregroundMask;
    }
    mesh_util::applyMaskToVertexesAndFaces(vertexes, faces, vertexMask);
    if (triangles > 0) {  // ConvertToBinary.cpp:186-203
      static const bool kIsEquierror = true;
      static const int kThreads = 1;
      render::MeshSimplifier ms(vertexes, faces, kIsEquierror, kThreads);
      static const float kStrictness = 0.2;
      static const bool kRemoveBoundaryEdges = false;
      ms.simplify(triangles, kStrictness, kRemoveBoundaryEdges);
      vertexes = ms.getVertexes();
      faces = ms.getFaces();
      for (int i = 0; i < vertexes.rows(); ++i) {
        if (vertexes.row(i).z() < 0) {
          vertexes.row(i).z() = FLT_MIN;
        }
      }
    }
    Eigen::Matrix<float, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> v = vertexes.cast<float>();  // writeDepth
    Eigen::Matrix<uint32_t, Eigen::Dynamic, Eigen::Dynamic, Eigen::RowMajor> f = faces.cast<uint32_t>();
    std::memcpy(vertexesOut, v.data(), v.size() * sizeof(float));
    std::memcpy(facesOut, f.data(), f.size() * sizeof(uint32_t));
    *num_vertexes = (uint64_t)vertexes.rows();
    *num_faces = (uint64_t)faces.rows();
    return (int)DERP_OK;
  });
}

int derp_camera_mesh(int /*device*/, const float* disparity, int width, int height, double depth_scale, double resolution_x,
                     double resolution_y, double scalar_focal, float tear_ratio, const uint8_t* foreground_mask,
                     int mask_width, int mask_height, float* vertexes, uint32_t* faces, uint64_t* num_vertexes,
                     uint64_t* num_faces) {
  return cameraMeshRef(disparity, width, height, depth_scale, resolution_x, resolution_y, scalar_focal, tear_ratio,
                       foreground_mask, mask_width, mask_height, 0, vertexes, faces, num_vertexes, num_faces);
}

/* + the reference's own MeshSimplifier (source/render/MeshSimplifier.cpp, compiled unmodified) with convertDepth's
 * constants: equi-error costs, one thread, strictness 0.2, boundary edges kept, negative z -> FLT_MIN. */
int derp_camera_mesh_simplified(int /*device*/, const float* disparity, int width, int height, double depth_scale,
                                double resolution_x, double resolution_y, double scalar_focal, float tear_ratio,
                                const uint8_t* foreground_mask, int mask_width, int mask_height, int triangles,
                                float* vertexes, uint32_t* faces, uint64_t* num_vertexes, uint64_t* num_faces) {
  return cameraMeshRef(disparity, width, height, depth_scale, resolution_x, resolution_y, scalar_focal, tear_ratio,
                       foreground_mask, mask_width, mask_height, triangles, vertexes, faces, num_vertexes, num_faces);
}

/* test hook (not part of derp_b200.h): the reference's MeshSimplifier on an arbitrary mesh */
int derp_ref_simplify(const double* xyz, uint64_t nv, const uint32_t* idx, uint64_t nf, int triangles, float strictness,
                      int remove_boundary_edges, double* out_xyz, uint32_t* out_idx, uint64_t* out_nv, uint64_t* out_nf) {
  return guarded([&] {
    Eigen::MatrixXd vertexes((Eigen::Index)nv, 3);
    Eigen::MatrixXi faces((Eigen::Index)nf, 3);
    for (uint64_t i = 0; i < nv; ++i)
      for (int j = 0; j < 3; ++j) vertexes((Eigen::Index)i, j) = xyz[3 * i + j];
    for (uint64_t i = 0; i < nf; ++i)
      for (int j = 0; j < 3; ++j) faces((Eigen::Index)i, j) = (int)idx[3 * i + j];
    render::MeshSimplifier ms(vertexes, faces, true, 1);
    ms.simplify(triangles, strictness, remove_boundary_edges != 0);
    vertexes = ms.getVertexes();
    faces = ms.getFaces();
    for (Eigen::Index i = 0; i < vertexes.rows(); ++i)
      for (int j = 0; j < 3; ++j) out_xyz[3 * i + j] = vertexes(i, j);
    for (Eigen::Index i = 0; i < faces.rows(); ++i)
      for (int j = 0; j < 3; ++j) out_idx[3 * i + j] = (uint32_t)faces(i, j);
    *out_nv = (uint64_t)vertexes.rows();
    *out_nf = (uint64_t)faces.rows();
    return (int)DERP_OK;
  });
}

/* test hook: mesh_util::writePfm (MeshUtil.h:35-70), the rasterised "pfm" format of ConvertToBinary, on a mesh given as
 * arrays.  calcBarycentrics solves its 2 x 2 system with Eigen's colPivHouseholderQr; the stand-in header solves it by
 * elimination with partial pivoting, so coverage exactly ON triangle edges can differ from a build with Eigen itself. */
int derp_ref_write_raster_pfm(const double* xyz, uint64_t nv, const uint32_t* idx, uint64_t nf, int width, int height,
                              double resolution_x, double resolution_y, const char* path) {
  return guarded([&] {
    Eigen::MatrixXd vertexes((Eigen::Index)nv, 3);
    Eigen::MatrixXi faces((Eigen::Index)nf, 3);
    for (uint64_t i = 0; i < nv; ++i)
      for (int j = 0; j < 3; ++j) vertexes((Eigen::Index)i, j) = xyz[3 * i + j];
    for (uint64_t i = 0; i < nf; ++i)
      for (int j = 0; j < 3; ++j) faces((Eigen::Index)i, j) = (int)idx[3 * i + j];
    const cv::Mat_<float> original(height, width, 0.0f);
    mesh_util::writePfm(original, Camera::Vector2(resolution_x, resolution_y), vertexes, faces, filesystem::path(path));
    return (int)DERP_OK;
  });
}

/* bench / test hook (not part of derp_b200.h): candidate slices of the brute-force cost volume the way the reference
 * builds them (Derp.cpp:288-304): ONE ThreadPool task per candidate, each writing a full-size cost and confidence map
 * (NaN where ignored).  Rows [y0, y1) only, so that a bench step can be a bounded band of the frame: with the full
 * interior (y0 = 1, y1 = H - 1) the reference's own computeBruteForceCosts runs; for a band its loop
 * (Derp.cpp:246-259) is restated around the reference's computeCost.  costs / confs: optional outputs, rows y0..y1-1
 * of every slice: [n][y1 - y0][W] (y0, y1 after clamping to the interior). */
int derp_ref_cost_slices(DerpCtx* c, int dst, const float* disparities, int n, int y0, int y1, float* costs, float* confs) {
  return guarded([&] {
    if (int rc = needProjected(c, dst)) return rc;
    Level& L = *c->level;
    const int W = c->lp.width, H = c->lp.height;
    y0 = std::max(y0, kSearchWindowRadius);
    y1 = std::min(y1, H - kSearchWindowRadius);
    const bool full = y0 == kSearchWindowRadius && y1 == H - kSearchWindowRadius;
    std::vector<cv::Mat_<float>> cm(n), fm(n);
    ThreadPool threadPool(g_threads);
    for (int i = 0; i < n; ++i) {
      cm[i].create(L.sizeLevel);
      cm[i].setTo(NAN);
      fm[i].create(L.sizeLevel);
      fm[i].setTo(NAN);
      if (full) {
        threadPool.spawn(&computeBruteForceCosts, std::ref(L), dst, disparities[i], std::ref(cm[i]), std::ref(fm[i]));
      } else {
        threadPool.spawn([&, i] {
          const float disparity = disparities[i];
          const bool useBg = L.hasForegroundMasks;
          for (int y = y0; y < y1; ++y)
            for (int x = kSearchWindowRadius; x < W - kSearchWindowRadius; ++x) {
              const bool ignore = !L.dstFovMask(dst)(y, x) || !L.dstForegroundMask(dst)(y, x) ||
                  (useBg && !(L.dstBackgroundDisparity(dst)(y, x) < disparity));
              if (ignore) continue;  // maps are pre-filled with NaN
              std::tie(cm[i](y, x), fm[i](y, x)) = computeCost(L, dst, disparity, x, y);
            }
        });
      }
    }
    threadPool.join();
    const size_t band = (size_t)(y1 - y0) * W;
    for (int i = 0; i < n; ++i) {
      if (costs) std::memcpy(costs + i * band, cm[i].ptr<float>(y0), band * 4);
      if (confs) std::memcpy(confs + i * band, fm[i].ptr<float>(y0), band * 4);
    }
    return DERP_OK;
  });
}

/* test hook (not part of derp_b200.h): the scratch directory processLevel's saveResults wrote its PFMs into */
const char* derp_ref_output_dir(DerpCtx* c) { return c ? c->tmpDir.c_str() : ""; }

}  // extern "C"
