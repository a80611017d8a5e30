// area_resize.h — cv::resize(..., INTER_AREA) on interleaved host images of float or 8-bit samples: what cv_util::resizeImage
// applies to the colour guide of UpsampleDisparity (UpsampleDisparity.cpp:117, Vec3f images, either direction) and what
// cv_util::scaleImage / resizeImage (CvUtil.h:138-154) apply to the colour images of ConvertToBinary when --color_scale < 1
// (image_util::loadScaledImage<Vec4f> for the BC7 stream, <Vec4b> for the .rgba stream; ConvertToBinary.cpp:127-142).  The
// UI's export tab sets that flag whenever it exports below the full width (scripts/ui/export.py:311-318, 389).
// Restates OpenCV's three code paths (modules/imgproc/src/resize.cpp), pinned to cv2 in tests/test_apps.py:
//   * both ratios exactly 2 ("ResizeAreaFastVec"): 8-bit: (a + b + c + d + 2) >> 2; float with 1 or 4 channels (the SIMD
//     body): ((a + b) + (c + d)) * 0.25f; float with other channel counts: the generic sum below;
//   * other integer ratios ("resizeAreaFast_"): the ky * kx samples summed row by row (int for 8-bit; float, in OpenCV's
//     groups of four, for float), times 1.f / (kx * ky), rounded half-to-even and saturated for 8-bit;
//   * everything else ("ResizeArea_Invoker"): separable weighted sums in float from the taps of computeResizeAreaTab.
// The 16-bit 3-channel case of the depth path has its own CUDA kernel (derp_downscale_area); these are host loops because
// their callers are host stages (PNG decode, gamma with the host's powf) of an app whose GPU work is the BC7 encoder.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace io {
namespace area {

struct Tap {
  int si, di;
  float alpha;
};
inline void taps(int ssize, int dsize, double scale, std::vector<Tap>& tab) {  // computeResizeAreaTab
  tab.clear();
  for (int dx = 0; dx < dsize; ++dx) {
    const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
    const double cell = std::min(scale, ssize - fsx1);
    int sx1 = (int)std::ceil(fsx1), sx2 = (int)std::floor(fsx2);
    sx2 = std::min(sx2, ssize - 1);
    sx1 = std::min(sx1, sx2);
    if (sx1 - fsx1 > 1e-3) tab.push_back(Tap{sx1 - 1, dx, (float)((sx1 - fsx1) / cell)});
    for (int sx = sx1; sx < sx2; ++sx) tab.push_back(Tap{sx, dx, float(1.0 / cell)});
    if (fsx2 - sx2 > 1e-3) tab.push_back(Tap{sx2, dx, (float)(std::min(std::min(fsx2 - sx2, 1.), cell) / cell)});
  }
}
inline int roundHalfEven(float v) { return (int)std::lrintf(v); }  // cvRound (the default rounding mode)
inline float store(float v, float*) { return v; }
inline uint8_t store(float v, uint8_t*) {
  const int r = roundHalfEven(v);
  return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}

// INTER_AREA when an axis grows ("true area interpolation is only implemented for the case scale >= 1; in other cases it is
// emulated using some variant of bilinear", resize.cpp): both axes then take two taps at sx = floor(dx * scale) with
// fx = (dx + 1) - (sx + 1) / scale, fx <= 0 ? 0 : fx - floor(fx); float weights, rows first, no rounding for float images.
inline void enlargeAxis(int ssize, int dsize, std::vector<int>& at, std::vector<float>& w0, std::vector<float>& w1) {
  const double scale = (double)ssize / dsize, inv = (double)dsize / ssize;
  at.resize(dsize);
  w0.resize(dsize);
  w1.resize(dsize);
  for (int d = 0; d < dsize; ++d) {
    int s = (int)std::floor(d * scale);
    double f = (d + 1) - (s + 1) * inv;
    f = f <= 0 ? 0. : f - std::floor(f);
    if (s < 0) {
      s = 0;
      f = 0;
    }
    if (s >= ssize - 1) {
      s = ssize - 1;
      f = 0;
    }
    at[d] = s;
    w1[d] = (float)f;
    w0[d] = 1.f - w1[d];
  }
}
inline void enlarge(const float* src, int sw, int sh, int cn, float* dst, int dw, int dh) {
  std::vector<int> xat, yat;
  std::vector<float> xa, xb, ya, yb;
  enlargeAxis(sw, dw, xat, xa, xb);
  enlargeAxis(sh, dh, yat, ya, yb);
  const size_t n = (size_t)dw * cn;
  std::vector<float> rows((size_t)sh * n);
  for (int y = 0; y < sh; ++y)
    for (int x = 0; x < dw; ++x) {
      const float* S0 = src + ((size_t)y * sw + xat[x]) * cn;
      const float* S1 = src + ((size_t)y * sw + std::min(xat[x] + 1, sw - 1)) * cn;
      for (int c = 0; c < cn; ++c) rows[(size_t)y * n + (size_t)x * cn + c] = S0[c] * xa[x] + S1[c] * xb[x];
    }
  for (int y = 0; y < dh; ++y) {
    const float* R0 = rows.data() + (size_t)yat[y] * n;
    const float* R1 = rows.data() + (size_t)std::min(yat[y] + 1, sh - 1) * n;
    for (size_t i = 0; i < n; ++i) dst[(size_t)y * n + i] = R0[i] * ya[y] + R1[i] * yb[y];
  }
}
inline void enlarge(const uint8_t*, int, int, int, uint8_t*, int, int) {}  // 8-bit images are never enlarged here (fixed point in OpenCV)

// src: sh rows of sw pixels of cn samples; dst: dh x dw (float images may also grow: see enlarge()).  simdAsFourChannels: use the 2 x 2 float
// formula of images with 1 or 4 channels although cn differs (the caller dropped a channel the reference still carries).
template <typename T>
inline void resize(const T* src, int sw, int sh, int cn, T* dst, int dw, int dh, bool simdAsFourChannels = false) {
  if (dw > sw || dh > sh) {  // an axis grows: the bilinear variant, for both axes
    enlarge(src, sw, sh, cn, dst, dw, dh);
    return;
  }
  const double fx = (double)sw / dw, fy = (double)sh / dh;
  const int kx = (int)std::floor(fx + 0.5), ky = (int)std::floor(fy + 0.5);
  const bool integer = std::fabs(fx - kx) < 2.220446049250313e-16 && std::fabs(fy - ky) < 2.220446049250313e-16;
  const size_t row = (size_t)sw * cn;
  if (integer) {
    const float scale = 1.f / (kx * ky);
    const bool two = kx == 2 && ky == 2;
    const bool isFloat = sizeof(T) == sizeof(float);
    for (int y = 0; y < dh; ++y)
      for (int x = 0; x < dw; ++x)
        for (int c = 0; c < cn; ++c) {
          const T* S = src + (size_t)y * ky * row + (size_t)x * kx * cn + c;
          T* D = dst + ((size_t)y * dw + x) * cn + c;
          if (two && !isFloat) {
            *D = (T)(((int)S[0] + (int)S[cn] + (int)S[row] + (int)S[row + cn] + 2) >> 2);
          } else if (two && (cn == 1 || cn == 4 || simdAsFourChannels)) {
            *D = (T)((((float)S[0] + (float)S[cn]) + ((float)S[row] + (float)S[row + cn])) * 0.25f);
          } else if (isFloat) {
            // samples in row order, added in groups of four (CV_ENABLE_UNROLLED): sum += ((s0 + s1) + s2) + s3
            const int area = kx * ky;
            auto sample = [&](int k) { return (float)S[(size_t)(k / kx) * row + (size_t)(k % kx) * cn]; };
            float sum = 0;
            int k = 0;
            for (; k <= area - 4; k += 4) sum += ((sample(k) + sample(k + 1)) + sample(k + 2)) + sample(k + 3);
            for (; k < area; ++k) sum += sample(k);
            *D = store(sum * scale, D);
          } else {
            int sum = 0;
            for (int j = 0; j < ky; ++j)
              for (int i = 0; i < kx; ++i) sum += (int)S[(size_t)j * row + (size_t)i * cn];
            *D = store((float)sum * scale, D);
          }
        }
    return;
  }
  std::vector<Tap> xtab, ytab;
  taps(sw, dw, fx, xtab);
  taps(sh, dh, fy, ytab);
  const size_t n = (size_t)dw * cn;
  std::vector<float> buf(n), sum(n, 0.f);
  int prev = ytab.empty() ? 0 : ytab[0].di;
  for (const Tap& ty : ytab) {
    const T* S = src + (size_t)ty.si * row;
    std::fill(buf.begin(), buf.end(), 0.f);
    for (const Tap& tx : xtab)
      for (int c = 0; c < cn; ++c) buf[(size_t)tx.di * cn + c] = buf[(size_t)tx.di * cn + c] + (float)S[(size_t)tx.si * cn + c] * tx.alpha;
    if (ty.di != prev) {
      for (size_t i = 0; i < n; ++i) {
        dst[(size_t)prev * n + i] = store(sum[i], dst);
        sum[i] = ty.alpha * buf[i];
      }
      prev = ty.di;
    } else {
      for (size_t i = 0; i < n; ++i) sum[i] += ty.alpha * buf[i];
    }
  }
  for (size_t i = 0; i < n; ++i) dst[(size_t)prev * n + i] = store(sum[i], dst);
}

}  // namespace area
}  // namespace io
