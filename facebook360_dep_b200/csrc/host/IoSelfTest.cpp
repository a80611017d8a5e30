// Test helper (not a reference app): loads an image with the apps' loaders and dumps the converted
// samples, so tests can compare the PNG/PFM readers and writers with cv2.
#include <atomic>
#include <chrono>
#include <thread>

#include "area_resize.h"
#include "exchange.h"
#include "io.h"

DEFINE_string(in, "", "input image");
DEFINE_string(mode, "color", "color | rgba | float | mask | rig | exchange | inflate | raster | area | bc7surface | gauss");
DEFINE_int32(size, 0, "mode=inflate: number of bytes the zlib stream in --in decodes to");
DEFINE_string(faces, "", "mode=raster: .idx file (uint32 x 3 per face); --in is the .vtx file (float32 x 3 per vertex)");
DEFINE_int32(width, 0, "mode=raster: depth grid width");
DEFINE_int32(height, 0, "mode=raster: depth grid height");
DEFINE_double(resolution_x, 0, "mode=raster: camera resolution");
DEFINE_double(resolution_y, 0, "mode=raster: camera resolution");
DEFINE_int32(channels, 4, "mode=area: samples per pixel");
DEFINE_int32(dst_width, 0, "mode=area: output width");
DEFINE_int32(dst_height, 0, "mode=area: output height");
DEFINE_string(type, "f32", "mode=area: f32 | u8 samples in the raw file --in (--width x --height pixels)");
DEFINE_bool(simd4, false, "mode=area: float 2 x 2 formula of 4-channel images");
DEFINE_double(scale, 1, "mode=bc7surface / rgba: --color_scale");
DEFINE_double(gamma, 2.2 / 1.8, "mode=bc7surface: --gamma_correction");
DEFINE_string(out, "", "output file (raw samples, or .png/.pfm for mode=float)");

int main(int argc, char** argv) {
  flags::initDep(argc, argv, "IoSelfTest --in=<image> --mode=color|float|mask|rig --out=<file>");
  int w = 0, h = 0;
  std::ofstream o;
  if (FLAGS_mode == "color") {
    const auto v = io::loadColor16(FLAGS_in, &w, &h);
    o.open(FLAGS_out, std::ios::binary);
    o.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)v.size() * 2);
  } else if (FLAGS_mode == "float") {
    const auto v = io::loadFloat(FLAGS_in, &w, &h);
    const fs::path outp(FLAGS_out);
    const std::string ext = outp.extension().string();
    if (ext == ".png" || ext == ".pfm") {
      fs::path stem = outp;
      stem.replace_extension("");
      io::saveDisparity(stem, ext.substr(1), v.data(), w, h);
    } else {
      o.open(FLAGS_out, std::ios::binary);
      o.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)v.size() * 4);
    }
  } else if (FLAGS_mode == "rgba") {
    auto v = io::loadRgba8(FLAGS_in, &w, &h);
    if (FLAGS_scale < 1) v = io::scaleRgba8(v, &w, &h, FLAGS_scale);
    o.open(FLAGS_out, std::ios::binary);
    o.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)v.size());
  } else if (FLAGS_mode == "mask") {
    const auto v = io::loadMask(FLAGS_in, &w, &h);
    o.open(FLAGS_out, std::ios::binary);
    o.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)v.size());
  } else if (FLAGS_mode == "rig") {
    const io::Rig rig = io::loadRig(FLAGS_in);
    o.open(FLAGS_out, std::ios::binary);
    o.write(reinterpret_cast<const char*>(rig.cams.data()), (std::streamsize)(rig.cams.size() * sizeof(DerpCameraDesc)));
    for (const auto& id : rig.ids) std::printf("%s\n", id.c_str());
  } else if (FLAGS_mode == "raster") {
    // ConvertToBinary's "pfm" format: the mesh of --in (.vtx) / --faces (.idx) rasterised onto a --width x --height grid
    auto slurp = [](const std::string& path) {
      std::ifstream f(path, std::ios::binary);
      return std::vector<char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    };
    const std::vector<char> v = slurp(FLAGS_in), i = slurp(FLAGS_faces);
    const std::vector<float> r = io::rasterMesh(reinterpret_cast<const float*>(v.data()), reinterpret_cast<const uint32_t*>(i.data()),
                                                i.size() / 12, FLAGS_width, FLAGS_height, FLAGS_resolution_x, FLAGS_resolution_y);
    io::writePfm(FLAGS_out, r.data(), FLAGS_width, FLAGS_height);
    w = FLAGS_width;
    h = FLAGS_height;
  } else if (FLAGS_mode == "bc7surface") {
    // the RGBA8 surface ConvertToBinary feeds the BC7 encoder for --in with --color_scale = --scale (< 1)
    const std::vector<uint8_t> v = io::bc7SurfaceScaled(io::loadUnchanged(FLAGS_in), FLAGS_scale, (float)FLAGS_gamma, &w, &h);
    o.open(FLAGS_out, std::ios::binary);
    o.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)v.size());
  } else if (FLAGS_mode == "gauss") {
    // cv::GaussianBlur((2 r + 1)^2, sigma 0) of a raw u16 x 3 image (--width x --height), r = --size
    std::ifstream f(FLAGS_in, std::ios::binary);
    const std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::vector<uint16_t> img((size_t)FLAGS_width * FLAGS_height * 3);
    std::memcpy(img.data(), raw.data(), img.size() * 2);
    const std::vector<uint16_t> v = io::gaussianBlurU16C3(img, FLAGS_width, FLAGS_height, FLAGS_size);
    o.open(FLAGS_out, std::ios::binary);
    o.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)v.size() * 2);
    w = FLAGS_width;
    h = FLAGS_height;
  } else if (FLAGS_mode == "area") {
    // cv::resize(INTER_AREA) of a raw interleaved image (area_resize.h) for comparison with cv2
    std::ifstream f(FLAGS_in, std::ios::binary);
    const std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const size_t n = (size_t)FLAGS_dst_width * FLAGS_dst_height * FLAGS_channels;
    o.open(FLAGS_out, std::ios::binary);
    if (FLAGS_type == "f32") {
      std::vector<float> d(n);
      io::area::resize(reinterpret_cast<const float*>(raw.data()), FLAGS_width, FLAGS_height, FLAGS_channels, d.data(), FLAGS_dst_width,
                       FLAGS_dst_height, FLAGS_simd4);
      o.write(reinterpret_cast<const char*>(d.data()), (std::streamsize)n * 4);
    } else {
      std::vector<uint8_t> d(n);
      io::area::resize(reinterpret_cast<const uint8_t*>(raw.data()), FLAGS_width, FLAGS_height, FLAGS_channels, d.data(),
                       FLAGS_dst_width, FLAGS_dst_height);
      o.write(reinterpret_cast<const char*>(d.data()), (std::streamsize)n);
    }
    w = FLAGS_dst_width;
    h = FLAGS_dst_height;
  } else if (FLAGS_mode == "inflate") {
    // the PNG reader's own inflate (inflate.h) against zlib on a raw zlib stream: accepted streams must decode to zlib's
    // bytes; prints "same" / "declined" / "DIFFERENT" and both times
    std::ifstream f(FLAGS_in, std::ios::binary);
    const std::vector<uint8_t> z((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::vector<uint8_t> mine((size_t)FLAGS_size + 1, 0xCD), theirs((size_t)FLAGS_size + 1, 0xEF);
    auto t0 = std::chrono::steady_clock::now();
    const bool accepted = io::inflate::zlibDecode(z.data(), z.size(), mine.data(), (size_t)FLAGS_size);
    auto t1 = std::chrono::steady_clock::now();
    uLongf outLen = (uLongf)FLAGS_size;
    const int zr = uncompress(theirs.data(), &outLen, z.data(), (uLong)z.size());
    auto t2 = std::chrono::steady_clock::now();
    const bool zok = zr == Z_OK && outLen == (uLongf)FLAGS_size;
    const bool guard = mine[(size_t)FLAGS_size] == 0xCD;  // nothing written past the end
    const char* verdict = !guard ? "OVERRUN" : !accepted ? "declined" : (zok && std::memcmp(mine.data(), theirs.data(), (size_t)FLAGS_size) == 0) ? "same" : "DIFFERENT";
    std::printf("%s zlib_ok=%d own %.4f s zlib %.4f s\n", verdict, (int)zok, std::chrono::duration<double>(t1 - t0).count(),
                std::chrono::duration<double>(t2 - t1).count());
    return 0;
  } else if (FLAGS_mode == "exchange") {
    // 5 threads x 200 rounds through the workers' rendezvous: between two barriers every thread must observe the
    // values all threads published before the first one (the protocol of DerpCLI's camera-sharded mismatch levels)
    const int parties = 5, rounds = 200;
    Exchange ex(parties, parties);
    std::vector<float> slots(parties, 0.f);
    std::atomic<int> bad{0};
    std::vector<std::thread> th;
    for (int g = 0; g < parties; ++g)
      th.emplace_back([&, g] {
        for (int r = 1; r <= rounds; ++r) {
          slots[g] = (float)(r * 100 + g);
          ex.planes[g] = &slots[g];
          ex.arriveAndWait();  // everything published
          for (int k = 0; k < parties; ++k)
            if (ex.planes[k] != &slots[k] || *ex.planes[k] != (float)(r * 100 + k)) bad++;
          ex.arriveAndWait();  // everybody has read: values may change
        }
      });
    for (auto& t : th) t.join();
    std::printf("exchange %s\n", bad.load() == 0 ? "ok" : "FAILED");
    return bad.load() == 0 ? 0 : 1;
  } else {
    LOG(FATAL) << "bad mode";
  }
  std::printf("%d %d\n", w, h);
  return 0;
}
