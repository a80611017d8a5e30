// UpsampleDisparity — drop-in for source/depth_estimation/UpsampleDisparity.cpp on B200.
// Same flags and directory contract; upsampling (UpsampleDisparityLib.cpp:98-182) and the optional
// colour-guided joint bilateral filter run in libderp_b200.so.  Colour images of another size than the output go through
// cv::resize(INTER_AREA) like cv_util::resizeImage does (UpsampleDisparity.cpp:117) on the [0, 1] float image the reference
// holds (PixelType = cv::Vec3f): a host stage (area_resize.h, bit-identical to cv2) before the guided filter on the GPU.
#include <thread>

#include "io.h"

const std::string kUsage = R"(
   - Reads a disparity image and upsamples to the given resolution.

   - Example:
     ./UpsampleDisparity \
     --rig=/path/to/rigs/rig.json \
     --disparity=/path/to/output/disparity \
     --output=/path/to/output/disparity_upsample \
     --resolution=4096 \
     --color=/path/to/video/color_full_size \
     --first=000000 \
     --last=000000
)";

DEFINE_string(background_disp, "", "background disparity directory (output resolution)");
DEFINE_string(background_frame, "000000", "background frame (lexical)");
DEFINE_string(cameras, "", "destination cameras");
DEFINE_string(color, "", "color directory (output resolution)");
DEFINE_string(disparity, "", "disparity directory (input resolution) (required)");
DEFINE_string(first, "000000", "first frame to process (lexical)");
DEFINE_string(foreground_masks_in, "", "(optional) masks directory (input resolution)");
DEFINE_string(foreground_masks_out, "", "(optional) masks directory (output resolution)");
DEFINE_int32(height, -1, "output image height (aspect ratio maintained if unspecified)");
DEFINE_string(last, "000000", "last frame to process (lexical)");
DEFINE_string(output, "", "output directory (required)");
DEFINE_string(output_formats, "", "saved formats, comma separated (exr, png, pfm supported)");
DEFINE_int32(resolution, -1, "output resolution width in pixels (required)");
DEFINE_string(rig, "", "path to camera rig .json");
DEFINE_double(sigma, 0.05, "bilateral filter color difference sigma");
DEFINE_int32(threads, -1, "number of threads (-1 = auto, 0 = none)");
DEFINE_double(weight_b, 0.5, "bilateral filter blue channel weight");
DEFINE_double(weight_g, 0.5, "bilateral filter green channel weight");
DEFINE_double(weight_r, 1.0, "bilateral filter red channel weight");
DEFINE_int32(gpu, 0, "first CUDA device to use");
DEFINE_int32(gpus, 1, "number of GPUs of this box to shard frames across");

#define DERP_CALL(expr)                                                 \
  do {                                                                  \
    const int rc_ = (expr);                                             \
    if (rc_ != 0) LOG(FATAL) << #expr << " failed: " << derp_last_error(); \
  } while (0)

// cv_util::loadImage<cv::Vec3f> (UpsampleDisparity.cpp:57 PixelType = Vec3f): integer samples * (1/max)
static std::vector<float> loadColorF32(const fs::path& p, int* w, int* h) {
  const io::Image img = io::loadUnchanged(p);
  *w = img.w;
  *h = img.h;
  const size_t n = (size_t)img.w * img.h;
  std::vector<float> out(n * 3);
  const float scale = img.bits == 16 ? 1.0f / 65535.0f : 1.0f / 255.0f;
  for (size_t i = 0; i < n; ++i)
    for (int c = 0; c < 3; ++c) {
      const size_t idx = i * img.channels + (img.channels >= 3 ? c : 0);
      out[i * 3 + c] = img.bits == 32 ? img.f[idx] : img.u[idx] * scale;
    }
  return out;
}

static void upsampleFrame(const io::Rig& rig, const std::vector<int>& dst, const std::string& frame, int device) {  // :65-144
  const std::string exts = FLAGS_output_formats.empty() ? "pfm" : FLAGS_output_formats;
  int height;
  const DerpCameraDesc& c0 = rig.cams[dst[0]];
  if (FLAGS_height == -1) {
    height = (int)std::round(float(c0.resolution[1]) / c0.resolution[0] * FLAGS_resolution);
    height += height % 2;  // force even height
  } else {
    height = FLAGS_height;
  }
  const int W = FLAGS_resolution, H = height;
  const bool useFg = !FLAGS_foreground_masks_in.empty();
  for (size_t i = 0; i < dst.size(); ++i) {
    const std::string& id = rig.ids[dst[i]];
    int cw, ch, w, h;
    const std::vector<float> disp = io::loadFloat(io::imagePath(FLAGS_disparity, id, frame), &cw, &ch);
    std::vector<float> bg;
    if (!FLAGS_background_disp.empty()) {
      bg = io::loadFloat(io::imagePath(FLAGS_background_disp, id, FLAGS_background_frame), &w, &h);
      CHECK(w == W && h == H) << "background disparity must have the output size";
    }
    std::vector<uint8_t> maskIn((size_t)cw * ch, 1), maskUp((size_t)W * H, 1);
    if (useFg) {
      maskIn = io::loadMask(io::imagePath(FLAGS_foreground_masks_in, id, frame), &w, &h);
      CHECK(w == cw && h == ch) << "input mask size mismatch";
    }
    if (!FLAGS_foreground_masks_out.empty()) {
      maskUp = io::loadMask(io::imagePath(FLAGS_foreground_masks_out, id, frame), &w, &h);
      CHECK(w == W && h == H) << "Desired resolution does not match mask resolution";
    }
    std::vector<float> up((size_t)W * H);
    DERP_CALL(derp_upsample_disparity(device, &rig.cams[dst[i]], disp.data(), cw, ch, bg.empty() ? nullptr : bg.data(),
                                      maskIn.data(), maskUp.data(), W, H, useFg ? 1 : 0, up.data()));
    if (!FLAGS_color.empty()) {
      const float scale = float(W) / float(cw);  // getRadius (UpsampleDisparityLib.cpp:93-96)
      const int radius = (int)(scale * scale + 1);
      LOG(INFO) << "Applying filter with radius " << radius << " to " << W << "x" << H << " disparity to " << id << "...";
      std::vector<float> color = loadColorF32(io::imagePath(FLAGS_color, id, frame), &w, &h);
      if (!(w == W && h == H)) {  // cv_util::resizeImage(colors[i], sizeUp): cv::resize INTER_AREA of the Vec3f image — the
        std::vector<float> resized((size_t)W * H * 3);  // reference's arithmetic (float area sums; the bilinear variant when
        io::area::resize(color.data(), w, h, 3, resized.data(), W, H);  // the image has to grow), pinned to cv2
        color.swap(resized);
      }
      std::vector<float> filtered((size_t)W * H);
      DERP_CALL(derp_joint_bilateral_f32(device, W, H, up.data(), color.data(), maskUp.data(), radius, (float)FLAGS_sigma,
                                         (float)FLAGS_weight_b, (float)FLAGS_weight_g, (float)FLAGS_weight_r,
                                         filtered.data()));
      up.swap(filtered);
    }
    LOG(INFO) << "Saving output images...";
    std::stringstream ss(exts);
    std::string ext;
    while (std::getline(ss, ext, ',')) {
      if (ext.empty()) continue;
      if (ext[0] == '.') ext = ext.substr(1);
      io::saveDisparity(fs::path(FLAGS_output) / id / frame, ext, up.data(), W, H);
    }
  }
}

int main(int argc, char** argv) {
  flags::initDep(argc, argv, kUsage);
  CHECK_NE(FLAGS_disparity, "");
  CHECK_NE(FLAGS_output, "");
  CHECK_NE(FLAGS_resolution, -1);
  const io::Rig rig = io::loadRig(FLAGS_rig);
  const std::vector<int> dst = io::filterDestinations(rig, FLAGS_cameras);
  CHECK_GT(dst.size(), 0u) << "no destination cameras!";
  // getFrameRange (ImageUtil.cpp:40-58): empty first/last = first/last file of the first camera
  std::string first = FLAGS_first, last = FLAGS_last;
  if (first.empty() || last.empty()) {
    const auto files = io::visibleFilesSorted(fs::path(FLAGS_disparity) / rig.ids[dst[0]]);
    CHECK_GT(files.size(), 0u) << "No files found in " << FLAGS_disparity;
    if (first.empty()) first = files.front().stem().string();
    if (last.empty()) last = files.back().stem().string();
  }
  // frames are independent: contiguous frame blocks per GPU, one worker thread per GPU
  const int firstFrame = std::stoi(first), numFrames = std::stoi(last) - firstFrame + 1;
  CHECK_GT(numFrames, 0);
  const int G = std::max(1, std::min(FLAGS_gpus, numFrames));
  LOG(INFO) << "backend " << derp_backend() << ", " << G << " GPU(s)";
  const int per = (numFrames + G - 1) / G;
  std::vector<std::thread> threads;
  for (int g = 0; g < G; ++g)
    threads.emplace_back([&, g] {
      for (int i = g * per; i < std::min(numFrames, (g + 1) * per); ++i)
        upsampleFrame(rig, dst, io::zeroPad(firstFrame + i), FLAGS_gpu + g);
    });
  for (auto& t : threads) t.join();
  return EXIT_SUCCESS;
}
