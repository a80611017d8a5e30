// LayerDisparities — drop-in for facebook360_dep's source/depth_estimation/LayerDisparities.cpp: composites a
// foreground disparity over a background disparity per camera and frame.  Same flags, same input/output tree, same
// 8-bit PNG output.  Pure file IO + one select per pixel (12 bytes of file data per pixel): there is nothing for the
// GPU to do here, so this app is host code only, like the reference's.
#include <cmath>

#include "io.h"

const std::string kUsageMessage = R"(
   - Layers foreground disparity atop background disparity assuming nans to correspond to locations
   without valid disparities.

   - Example:
     ./LayerDisparities \
     --rig=/path/to/rigs/rig.json \
     --background_disp=/path/to/background/disparity \
     --foreground_disp=/path/to/output/disparity \
     --output=/path/to/output \
     --first=000000 \
     --last=000000
 )";

DEFINE_string(background_disp, "", "path to background disparity directory (required)");
DEFINE_string(background_frame, "000000", "background frame to process (lexical)");
DEFINE_string(cameras, "", "destination cameras");
DEFINE_string(first, "000000", "first frame to process (lexical)");
DEFINE_string(foreground_disp, "", "path to foreground disparity directory (required)");
DEFINE_string(last, "000000", "last frame to process (lexical)");
DEFINE_string(output, "", "path to output disparity directory");
DEFINE_string(rig, "", "path to camera rig .json (required)");
DEFINE_int32(threads, -1, "number of threads (-1 = auto, 0 = none)");

// layerDisparities (LayerDisparities.cpp:46-56): mask = threshold(fg, 0, 1, THRESH_BINARY) (NaN > 0 is false);
// layer = fg .* mask + bg .* (1 - mask) evaluated in float exactly like the matrix expression (so a NaN foreground
// stays NaN: NaN * 0); written through cv::imwrite of a float matrix = convertTo(CV_8U) = saturate(cvRound(255 * v)),
// NaN / inf -> 0, as an 8-bit single-channel PNG.
static void layerDisparities(const std::vector<float>& fg, const std::vector<float>& bg, int w, int h,
                             const fs::path& outputPath) {
  std::vector<uint8_t> out(fg.size());
  for (size_t i = 0; i < fg.size(); ++i) {
    const float mask = fg[i] > 0.0f ? 1.0f : 0.0f;
    const float layer = fg[i] * mask + bg[i] * (1 - mask);
    const float v = layer * 255;
    // cvRound = cvtss2si: round half to even; NaN, +-inf and anything outside int32 give INT_MIN, which saturates to 0
    const bool inRange = v > -2147483648.0f && v < 2147483648.0f;
    const long r = inRange ? std::lrint(v) : 0;
    out[i] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
  }
  io::writePng8Gray(outputPath, out.data(), w, h);
}

int main(int argc, char* argv[]) {
  flags::initDep(argc, argv, kUsageMessage);
  CHECK_NE(FLAGS_rig, "");
  CHECK_NE(FLAGS_background_disp, "");
  CHECK_NE(FLAGS_foreground_disp, "");
  CHECK_LE(FLAGS_first, FLAGS_last);

  const io::Rig rig = io::loadRig(FLAGS_rig);
  const std::vector<int> dst = io::filterDestinations(rig, FLAGS_cameras);
  CHECK_GT(dst.size(), 0u) << "no destination cameras!";

  std::vector<std::vector<float>> bgs(dst.size());
  std::vector<std::pair<int, int>> bgSize(dst.size());
  for (size_t c = 0; c < dst.size(); ++c)
    bgs[c] = io::loadFloat(io::imagePath(FLAGS_background_disp, rig.ids[dst[c]], FLAGS_background_frame), &bgSize[c].first,
                           &bgSize[c].second);
  const int numFrames = std::stoi(FLAGS_last) - std::stoi(FLAGS_first) + 1;
  for (int iFrame = 0; iFrame < numFrames; ++iFrame) {
    const std::string frameName = io::zeroPad(iFrame + std::stoi(FLAGS_first));
    for (size_t c = 0; c < dst.size(); ++c) {
      const std::string& id = rig.ids[dst[c]];
      int w, h;
      const std::vector<float> fg = io::loadFloat(io::imagePath(FLAGS_foreground_disp, id, frameName), &w, &h);
      CHECK(w == bgSize[c].first && h == bgSize[c].second) << "Background and foreground images must be of the same size!";
      // getImageDir(output, ImageType::disparity, id) (DerpUtil.cpp:278-291)
      const fs::path outputDir = fs::path(FLAGS_output) / io::kDisparity / id;
      fs::create_directories(outputDir);
      layerDisparities(fg, bgs[c], w, h, outputDir / (frameName + ".png"));
    }
  }
  return EXIT_SUCCESS;
}
