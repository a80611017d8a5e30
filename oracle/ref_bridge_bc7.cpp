// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// The reference's OWN BC7 colour path, compiled where it lies under /root/reference (never copied) into
// oracle/_ref/libderp_ref.so:
//
//   /root/reference/source/conversion/BC7Util.h:41-98            bc7_util::gammaCorrect, compressBC7 (packing to RGBA8,
//                                                                GetProfile_veryfast, CompressBlocksBC7, file write)
//   .../thirdparty/bc7_compressor/ISPCTextureCompressor/ispc/ispc_texcomp/ispc_texcomp.cpp   profiles + C wrappers
//   .../ispc_texcomp/kernel.ispc                                 the encoder itself, compiled by the ispc 1.10.0 binary the
//                                                                reference vendors next to it (ispc/ispc_linux) with the
//                                                                flags of the reference's ISPC.cmake:4
//                                                                (-O2 --arch=x86-64 --target=sse2,avx --opt=fast-math --pic)
//
// Note on determinism: --opt=fast-math turns every x / y into x * rcp(y) and the kernel calls rsqrt(); ispc implements
// both with the CPU's RCPPS / RSQRTPS estimate + one Newton step, whose low bits differ between CPU vendors.  The
// reference's .bc7 bytes are therefore a property of (code, CPU); tests/test_bc7.py compares block identity rates and
// decoded error, not a golden file.
#include <cstdint>
#include <cstring>
#include <string>

#include "source/conversion/BC7Util.h"

#include <unistd.h>

#include "../include/derp_b200.h"

using namespace fb360_dep;

extern "C" {

// compressBC7(image, path, gamma, writeDDSHeader) on a BGR float image (h x w x 3, values in [0, 1]) -> file `path`.
int derp_ref_bc7_file(const float* bgr, int width, int height, float gamma, int dds_header, const char* path) {
  try {
    cv::Mat_<cv::Vec3f> image(height, width);
    std::memcpy(image.ptr(), bgr, sizeof(float) * 3 * (size_t)width * height);
    bc7_util::compressBC7(image, filesystem::path(path), gamma, dds_header != 0);
    return 0;
  } catch (...) {
    return -1;
  }
}

// CompressBlocksBC7 with GetProfile_veryfast on an RGBA8 surface (what compressBC7 hands to the encoder).
int derp_ref_bc7_blocks(const uint8_t* rgba, int width, int height, uint8_t* out) {
  rgba_surface surface;
  surface.width = width;
  surface.height = height;
  surface.stride = width * 4;
  surface.ptr = const_cast<uint8_t*>(rgba);
  bc7_enc_settings settings;
  GetProfile_veryfast(&settings);
  CompressBlocksBC7(&surface, out, &settings);
  return 0;
}

uint8_t derp_ref_gamma_correct(float v, float gamma) { return bc7_util::gammaCorrect(v, gamma); }

// ---- the two ABI entries of include/derp_b200.h, on the reference's code ----
int derp_bc7_compress(int, const uint8_t* rgba, int width, int height, uint8_t* blocks) {
  if (!rgba || !blocks || width < 1 || height < 1) return DERP_EINVAL;
  std::memset(blocks, 0, (size_t)width * height);  // compressBC7's std::vector<unsigned char> bc7data(w * h)
  return derp_ref_bc7_blocks(rgba, width, height, blocks);
}

// image as cv::imread(IMREAD_UNCHANGED) holds it -> cv_util::convertTo(CV_32F) (CvUtil.h:196-207) -> compressBC7 to a
// temporary file (no DDS header, ConvertToBinary.cpp:135-136) -> bytes.  An alpha channel is dropped here instead of by
// cv::cvtColor(BGRA2BGR) (a stub in refshim); compressBC7 ignores alpha anyway (BC7Util.h:50, 66).
int derp_bc7_compress_image(int, const void* pixels, int bits_per_channel, int channels, int width, int height, float gamma,
                            uint8_t* blocks) {
  if (!pixels || !blocks || width < 1 || height < 1 || (bits_per_channel != 8 && bits_per_channel != 16) ||
      (channels != 3 && channels != 4))
    return DERP_EINVAL;
  try {
    const int depth = bits_per_channel == 8 ? CV_8U : CV_16U;
    cv::Mat stored(height, width, CV_MAKETYPE(depth, 3));
    const size_t elem = bits_per_channel / 8;
    for (int y = 0; y < height; ++y)
      for (int x = 0; x < width; ++x)
        std::memcpy(stored.ptr(y) + (size_t)x * 3 * elem,
                    static_cast<const uint8_t*>(pixels) + ((size_t)y * width + x) * channels * elem, 3 * elem);
    const cv::Mat image = cv_util::convertTo(stored, CV_32F);
    char tmpl[] = "/tmp/derp_ref_bc7_XXXXXX";
    const int fd = mkstemp(tmpl);
    if (fd < 0) return DERP_EINVAL;
    close(fd);
    bc7_util::compressBC7(image, filesystem::path(tmpl), gamma, false);
    FILE* f = std::fopen(tmpl, "rb");
    const size_t got = f ? std::fread(blocks, 1, (size_t)width * height, f) : 0;
    if (f) std::fclose(f);
    std::remove(tmpl);
    return got == (size_t)width * height ? DERP_OK : DERP_EINVAL;
  } catch (...) {
    return DERP_EINVAL;
  }
}
}
