// TemporalBilateralFilter — drop-in for source/depth_estimation/TemporalBilateralFilter.cpp on B200.
// Same flags, same directory contract (reads <disparity>/level_L/<cam>/<frame>, writes
// <output_root>/disparity_time_filtered_levels/level_L/<cam>/<frame>.{pfm,png}); the filter itself is
// the temporalKernel of libderp_b200.so (derp_temporal_filter).
// Reproduced verbatim (parity > elegance): weights are passed as (weight_b, weight_g, weight_b) and
// --weight_r is ignored (TemporalBilateralFilter.cpp:176-178).
//
// Frames live in DEVICE memory between filter calls: every frame of a GPU's block is decoded and uploaded once (the
// reference re-reads the +-time_radius window from disk for every frame, TemporalBilateralFilter.cpp:139-160), and with
// --gpus > 1 the halo frames of a block are copied from the neighbouring GPU's store over NVLink (derp_device_copy)
// after a rendezvous of the worker threads instead of being decoded a second time.
#include <map>
#include <thread>

#include "exchange.h"
#include "io.h"

static const int kTemporalSpaceRadiusMin = 1;
static const int kTemporalSpaceRadiusMax = 1;

const std::string kUsageMessage = R"(
  - Runs temporal filter across disparity frames using corresponding color frames as guides.

  - Example:
    ./TemporalBilateralFilter \
    --input_root=/path/to/ \
    --output_root=/path/to/output \
    --rig=/path/to/rigs/rig.json \
    --first=000000 \
    --last=000000
)";

DEFINE_string(color, "", "color directory");
DEFINE_string(cameras, "", "destination cameras");
DEFINE_string(disparity, "", "disparity directory");
DEFINE_string(first, "000000", "first frame to process (lexical)");
DEFINE_string(foreground_masks, "", "foreground masks directory");
DEFINE_string(input_root, "", "output root directory (required)");
DEFINE_string(last, "000000", "last frame to process (lexical)");
DEFINE_int32(level, 0, "pyramid level being processed");
DEFINE_string(output_formats, "", "saved formats, comma separated (exr, png, pfm supported)");
DEFINE_string(output_root, "", "output root directory (required)");
DEFINE_int32(resolution, 2048, "8192, 4096, 2048, 1024, 512, 256");
DEFINE_string(rig, "", "path to camera rig .json (required)");
DEFINE_double(sigma, 0.01, "spatio-temporal smoothing");
DEFINE_int32(space_radius, -1, "space filtering radius");
DEFINE_int32(threads, -1, "number of threads (-1 = auto, 0 = none)");
DEFINE_int32(time_radius, 2, "temporal filtering radius");
DEFINE_bool(use_foreground_masks, false, "use pre-computed foreground masks");
DEFINE_double(weight_b, 0.5, "Blue channel weight");
DEFINE_double(weight_g, 1.0, "Green channel weight");
DEFINE_double(weight_r, 1.0, "Red channel weight");
DEFINE_int32(gpu, 0, "first CUDA device to use");
DEFINE_int32(gpus, 1, "number of GPUs of this box to shard frames across");

#define DERP_CALL(expr)                                                 \
  do {                                                                  \
    const int rc_ = (expr);                                             \
    if (rc_ != 0) LOG(FATAL) << #expr << " failed: " << derp_last_error(); \
  } while (0)

// populateMinMaxFrame (TemporalBilateralFilter.cpp:96-119)
static void populateMinMaxFrame(const std::string& dir, int level, const std::string& camId, int cur, int& first,
                                int& last) {
  const std::string levelDir = io::levelDir(dir, level) + "/" + camId;
  const std::string ext = io::firstExtension(levelDir);
  int localFirst = INT32_MAX, localLast = 0;
  for (int f = cur - FLAGS_time_radius; f <= cur + FLAGS_time_radius; ++f)
    if (fs::exists(fs::path(levelDir) / (io::zeroPad(f) + ext))) {
      localFirst = std::min(f, localFirst);
      localLast = std::max(f, localLast);
    }
  first = std::max(localFirst, first);
  last = std::min(localLast, last);
}

// FOV mask of one destination at the level size, through the library (generateFovMasks)
static std::vector<std::vector<uint8_t>> fovMasks(const io::Rig& rig, const std::vector<int>& dst, int W, int H,
                                                  int device) {
  std::vector<int32_t> d2s(dst.begin(), dst.end());
  DerpCtx* ctx = nullptr;
  DERP_CALL(derp_create(rig.cams.data(), (int)rig.cams.size(), d2s.data(), (int)d2s.size(), device, &ctx));
  DerpLevelParams lp{};
  lp.width = W;
  lp.height = H;
  lp.num_levels = 1;
  lp.full_width = W;
  lp.full_height = H;
  DERP_CALL(derp_level_begin(ctx, &lp));
  std::vector<std::vector<uint8_t>> out(dst.size(), std::vector<uint8_t>((size_t)W * H));
  for (size_t d = 0; d < dst.size(); ++d) DERP_CALL(derp_get_fov_mask(ctx, (int)d, out[d].data()));
  derp_destroy(ctx);
  return out;
}

// One frame of the level on a GPU: per destination camera colour (u16 x 3), disparity (f32) and mask (u8) planes
struct DevFrame {
  int device = 0, W = 0, H = 0;
  std::vector<void*> color, disp, mask;
  bool valid = false;
};

// One GPU worker: its device, the FOV masks (level size only: computed once per worker) and its resident frames
struct Worker {
  int device = 0;
  std::vector<std::vector<uint8_t>> fov;
  std::map<int, DevFrame> frames;
};

static void freeFrame(DevFrame& f) {
  for (auto* v : {&f.color, &f.disp, &f.mask})
    for (void* p : *v) derp_device_free(f.device, p);
  f = DevFrame();
}

static bool frameOnDisk(int frame, const io::Rig& rig, const std::vector<int>& dst) {
  const std::string& ref = rig.ids[dst[0]];
  auto has = [&](const std::string& dir) {
    const std::string levelDir = io::levelDir(dir, FLAGS_level) + "/" + ref;
    return fs::exists(fs::path(levelDir) / (io::zeroPad(frame) + io::firstExtension(levelDir)));
  };
  return frame >= 0 && has(FLAGS_color) && has(FLAGS_disparity) && (!FLAGS_use_foreground_masks || has(FLAGS_foreground_masks));
}

// decode one frame of the level (all destination cameras) and upload it
static void loadFrame(int frame, const io::Rig& rig, const std::vector<int>& dst, Worker& wk) {
  DevFrame df;
  df.device = wk.device;
  const std::string name = io::zeroPad(frame);
  for (size_t ci = 0; ci < dst.size(); ++ci) {
    const std::string& id = rig.ids[dst[ci]];
    int w, h, W, H;
    const std::vector<uint16_t> color = io::loadColor16(io::imagePath(io::levelDir(FLAGS_color, FLAGS_level), id, name), &W, &H);
    const std::vector<float> disp = io::loadFloat(io::imagePath(io::levelDir(FLAGS_disparity, FLAGS_level), id, name), &w, &h);
    CHECK(w == W && h == H) << "colour / disparity size mismatch";
    if (wk.fov.empty()) wk.fov = fovMasks(rig, dst, W, H, wk.device);
    std::vector<uint8_t> mask;
    if (FLAGS_use_foreground_masks) {
      mask = io::loadMask(io::imagePath(io::levelDir(FLAGS_foreground_masks, FLAGS_level), id, name), &w, &h);
      CHECK(w == W && h == H) << "mask size mismatch";
      for (size_t i = 0; i < mask.size(); ++i) mask[i] = mask[i] & wk.fov[ci][i];
    } else {
      mask = wk.fov[ci];
    }
    df.W = W;
    df.H = H;
    const size_t n = (size_t)W * H;
    void *pc, *pd, *pm;
    DERP_CALL(derp_device_alloc(wk.device, n * 6, &pc));
    DERP_CALL(derp_device_alloc(wk.device, n * 4, &pd));
    DERP_CALL(derp_device_alloc(wk.device, n, &pm));
    DERP_CALL(derp_device_copy(wk.device, pc, color.data(), n * 6));
    DERP_CALL(derp_device_copy(wk.device, pd, disp.data(), n * 4));
    DERP_CALL(derp_device_copy(wk.device, pm, mask.data(), n));
    df.color.push_back(pc);
    df.disp.push_back(pd);
    df.mask.push_back(pm);
  }
  df.valid = true;
  wk.frames[frame] = df;
}

// copy a frame that is resident on a peer GPU (NVLink) instead of decoding it again
static void fetchFrame(int frame, const DevFrame& src, Worker& wk) {
  DevFrame df;
  df.device = wk.device;
  df.W = src.W;
  df.H = src.H;
  const size_t n = (size_t)src.W * src.H;
  for (size_t ci = 0; ci < src.color.size(); ++ci) {
    void *pc, *pd, *pm;
    DERP_CALL(derp_device_alloc(wk.device, n * 6, &pc));
    DERP_CALL(derp_device_alloc(wk.device, n * 4, &pd));
    DERP_CALL(derp_device_alloc(wk.device, n, &pm));
    DERP_CALL(derp_device_copy(wk.device, pc, src.color[ci], n * 6));
    DERP_CALL(derp_device_copy(wk.device, pd, src.disp[ci], n * 4));
    DERP_CALL(derp_device_copy(wk.device, pm, src.mask[ci], n));
    df.color.push_back(pc);
    df.disp.push_back(pd);
    df.mask.push_back(pm);
  }
  df.valid = true;
  wk.frames[frame] = df;
}

static void filterFrame(int cur, const io::Rig& rig, const std::vector<int>& dst, Worker& wk) {  // :121-184
  int first = 0, last = INT32_MAX;
  const std::string& ref = rig.ids[dst[0]];
  populateMinMaxFrame(FLAGS_color, FLAGS_level, ref, cur, first, last);
  populateMinMaxFrame(FLAGS_disparity, FLAGS_level, ref, cur, first, last);
  if (FLAGS_use_foreground_masks) populateMinMaxFrame(FLAGS_foreground_masks, FLAGS_level, ref, cur, first, last);
  CHECK_LE(first, last) << "no frames available around " << cur;
  const int T = last - first + 1;
  const float scale = (float)std::pow((double)0.9f, (double)FLAGS_level);
  const int spaceRadius = FLAGS_space_radius == -1
      ? (int)std::max(std::ceil(kTemporalSpaceRadiusMax * scale), float(kTemporalSpaceRadiusMin))
      : FLAGS_space_radius;
  for (int f = first; f <= last; ++f)
    if (!wk.frames.count(f)) loadFrame(f, rig, dst, wk);  // not resident (first use, or beyond what the peers hold)
  const int W = wk.frames[cur].W, H = wk.frames[cur].H;
  for (size_t ci = 0; ci < dst.size(); ++ci) {
    const std::string& id = rig.ids[dst[ci]];
    std::vector<const uint16_t*> g(T);
    std::vector<const float*> dp(T);
    std::vector<const uint8_t*> mp(T);
    for (int t = 0; t < T; ++t) {
      const DevFrame& df = wk.frames[first + t];
      CHECK(df.W == W && df.H == H) << "frame size mismatch";
      g[t] = static_cast<const uint16_t*>(df.color[ci]);
      dp[t] = static_cast<const float*>(df.disp[ci]);
      mp[t] = static_cast<const uint8_t*>(df.mask[ci]);
    }
    std::vector<float> out((size_t)W * H);
    DERP_CALL(derp_temporal_filter(wk.device, W, H, T, g.data(), dp.data(), mp.data(), cur - first, (float)FLAGS_sigma,
                                   spaceRadius, (float)FLAGS_weight_b, (float)FLAGS_weight_g, (float)FLAGS_weight_b,
                                   out.data()));
    // saveDisparity (:61-94): pfm always
    std::vector<std::string> formats = {"pfm"};
    std::stringstream ss(FLAGS_output_formats);
    std::string f;
    while (std::getline(ss, f, ','))
      if ((f == "png" || f == "exr") && std::find(formats.begin(), formats.end(), f) == formats.end()) formats.push_back(f);
    const fs::path stem = fs::path(io::levelDir(FLAGS_output_root + "/" + io::kDisparityTimeFilteredLevels, FLAGS_level)) /
        id / io::zeroPad(cur);
    for (const auto& ext : formats) io::saveDisparity(stem, ext, out.data(), W, H);
  }
  // frames before the next window are never needed again
  for (auto it = wk.frames.begin(); it != wk.frames.end();)
    if (it->first < cur + 1 - FLAGS_time_radius) {
      freeFrame(it->second);
      it = wk.frames.erase(it);
    } else {
      ++it;
    }
}

int main(int argc, char** argv) {
  flags::initDep(argc, argv, kUsageMessage);
  CHECK_NE(FLAGS_rig, "");
  CHECK_NE(FLAGS_input_root, "");
  CHECK_NE(FLAGS_output_root, "");
  if (FLAGS_color.empty()) FLAGS_color = FLAGS_input_root + "/" + io::kColorLevels;
  if (FLAGS_foreground_masks.empty()) FLAGS_foreground_masks = FLAGS_input_root + "/" + io::kForegroundMasksLevels;
  if (FLAGS_disparity.empty()) FLAGS_disparity = FLAGS_output_root + "/" + io::kDisparityLevels;
  const io::Rig rig = io::loadRig(FLAGS_rig);
  const std::vector<int> dst = io::filterDestinations(rig, FLAGS_cameras);
  CHECK_GT(dst.size(), 0u) << "no destination cameras!";
  // Contiguous frame blocks per GPU, one worker thread per GPU.  With several GPUs every worker first
  // makes its own block resident (as far as kMaxResident frames go), the workers meet, and each copies the halo frames
  // it needs from the neighbour that holds them; frames nobody holds are decoded from disk on first use.
  const int firstFrame = std::stoi(FLAGS_first), numFrames = std::stoi(FLAGS_last) - firstFrame + 1;
  CHECK_GT(numFrames, 0);
  const int G = std::max(1, std::min(FLAGS_gpus, numFrames));
  LOG(INFO) << "backend " << derp_backend() << ", " << G << " GPU(s)";
  const int per = (numFrames + G - 1) / G;
  const int kMaxResident = 2 * FLAGS_time_radius + 4;  // what a worker publishes to its neighbours
  std::vector<Worker> workers(G);
  Exchange meet(G, 0);
  std::vector<std::thread> threads;
  for (int g = 0; g < G; ++g)
    threads.emplace_back([&, g] {
      Worker& wk = workers[g];
      wk.device = FLAGS_gpu + g;
      const int b0 = firstFrame + g * per, b1 = firstFrame + std::min(numFrames, (g + 1) * per);
      if (G > 1) {
        // publish: the first and last frames of the block (the ones neighbours need), bounded
        for (int f = b0; f < b1; ++f)
          if ((f < b0 + FLAGS_time_radius || f >= b1 - FLAGS_time_radius) && (int)wk.frames.size() < kMaxResident &&
              frameOnDisk(f, rig, dst))
            loadFrame(f, rig, dst, wk);
        meet.arriveAndWait();  // every store is complete and read-only from here
        for (int f = b0 - FLAGS_time_radius; f < b1 + FLAGS_time_radius; ++f) {
          if (f >= b0 && f < b1) continue;
          const int owner = (f - firstFrame) / per;
          if (f < firstFrame || owner < 0 || owner >= G) continue;
          const auto it = workers[owner].frames.find(f);
          if (it != workers[owner].frames.end() && it->second.valid) fetchFrame(f, it->second, wk);
        }
        meet.arriveAndWait();  // all peer copies done: stores may evict from here
      }
      for (int f = b0; f < b1; ++f) {
        LOG(INFO) << "Filtering images... frame " << io::zeroPad(f);
        filterFrame(f, rig, dst, wk);
      }
      for (auto& kv : wk.frames) freeFrame(kv.second);
      wk.frames.clear();
    });
  for (auto& t : threads) t.join();
  return EXIT_SUCCESS;
}
