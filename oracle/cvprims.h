// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the five OpenCV primitives the reference's depth path calls.  OpenCV is a
// third-party dependency that is NOT under /root/reference (CMakeLists.txt:67 asks for "4", the
// Dockerfile:120 builds 3.4.3 — no single pinned version).  We restate the published algorithms
// of modules/imgproc (imgwarp.cpp remapBicubic, box_filter ColumnSum, resize.cpp Lanczos4 /
// nearest) and pin them to outputs of the cv2 4.13.0 wheel present in this image:
// tests/golden/cv_*.npz, generator tests/golden/gen_cv_vectors.py.
//   remap INTER_CUBIC / BORDER_CONSTANT, u16x3   : bit-exact vs cv2 (tests/test_oracle_cv.py)
//   blur 3x3 u16x3 (BORDER_REFLECT_101)          : bit-exact vs cv2
//   computeImageVariance pipeline (float blur)   : <= 1e-6 relative vs cv2 (SIMD/FMA dispatch)
//   resize INTER_LANCZOS4 float                  : <= 2e-6 absolute vs cv2
//   resize INTER_NEAREST                         : bit-exact vs cv2
// Call sites in the reference: DerpUtil.cpp:203 (remap), CvUtil.h:314-323 (blur),
// DerpUtil.cpp:214-237 (variance), UpsampleDisparityLib.cpp:125,145 (resize).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace oracle {

inline int cvRoundF(float v) { return (int)lrintf(v); }   // cvRound: round-half-even (SSE cvtss2si)
inline int cvRoundD(double v) { return (int)lrint(v); }
inline int cvFloorD(double v) {
  int i = (int)v;
  return i - (i > v);
}
inline uint16_t saturateU16FromInt(int v) { return (uint16_t)(v < 0 ? 0 : v > 65535 ? 65535 : v); }

// ---- cv::remap INTER_CUBIC ------------------------------------------------------------------
// imgwarp.cpp: interpolateCubic (A = -0.75), initInterTab1D/2D (INTER_TAB_SIZE = 32, float table).
struct BicubicTab {
  float w[32 * 32][16];
  BicubicTab() {
    float tab1[32][4];
    const float scale = 1.f / 32;
    for (int i = 0; i < 32; ++i) {
      const float x = i * scale;
      const float A = -0.75f;
      tab1[i][0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
      tab1[i][1] = ((A + 2) * x - (A + 3)) * x * x + 1;
      tab1[i][2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
      tab1[i][3] = 1.f - tab1[i][0] - tab1[i][1] - tab1[i][2];
    }
    for (int i = 0; i < 32; ++i)
      for (int j = 0; j < 32; ++j)
        for (int k1 = 0; k1 < 4; ++k1) {
          const float vy = tab1[i][k1];
          for (int k2 = 0; k2 < 4; ++k2) w[i * 32 + j][k1 * 4 + k2] = vy * tab1[j][k2];
        }
  }
};
inline const BicubicTab& bicubicTab() {
  static const BicubicTab t;
  return t;
}

// src: u16 [sh][sw][3]; map: float [dh][dw][2] (x,y in OpenCV convention); dst: u16 [dh][dw][3].
// remap() converts the float map to Q5 fixed point (cvRound(x*32)), then remapBicubic.
inline void remapBicubicU16C3(const uint16_t* src, int sw, int sh, const float* map, int dw, int dh,
                              uint16_t* dst, int y0 = 0, int y1 = -1) {
  const BicubicTab& tab = bicubicTab();
  const int cn = 3;
  const size_t sstep = (size_t)sw * cn;
  const unsigned width1 = (unsigned)std::max(sw - 3, 0), height1 = (unsigned)std::max(sh - 3, 0);
  if (y1 < 0) y1 = dh;
  for (int dy = y0; dy < y1; ++dy) {
    for (int dx = 0; dx < dw; ++dx) {
      const float mx = map[((size_t)dy * dw + dx) * 2], my = map[((size_t)dy * dw + dx) * 2 + 1];
      // cvRound of NaN/huge -> INT_MIN on x86; >>5 then saturate_cast<short>
      int sxq = cvRoundF(mx * 32), syq = cvRoundF(my * 32);
      if (!(mx == mx) || mx * 32 >= 2147483648.f || mx * 32 < -2147483648.f) sxq = INT32_MIN;
      if (!(my == my) || my * 32 >= 2147483648.f || my * 32 < -2147483648.f) syq = INT32_MIN;
      const int fidx = (syq & 31) * 32 + (sxq & 31);
      int ix = sxq >> 5, iy = syq >> 5;
      ix = ix < -32768 ? -32768 : ix > 32767 ? 32767 : ix;
      iy = iy < -32768 ? -32768 : iy > 32767 ? 32767 : iy;
      const int sx = ix - 1, sy = iy - 1;
      const float* w = tab.w[fidx];
      uint16_t* D = dst + ((size_t)dy * dw + dx) * cn;
      if ((unsigned)sx < width1 && (unsigned)sy < height1) {
        const uint16_t* S = src + (size_t)sy * sstep + (size_t)sx * cn;
        for (int k = 0; k < cn; ++k) {
          float sum = S[0] * w[0] + S[cn] * w[1] + S[cn * 2] * w[2] + S[cn * 3] * w[3];
          S += sstep;
          sum += S[0] * w[4] + S[cn] * w[5] + S[cn * 2] * w[6] + S[cn * 3] * w[7];
          S += sstep;
          sum += S[0] * w[8] + S[cn] * w[9] + S[cn * 2] * w[10] + S[cn * 3] * w[11];
          S += sstep;
          sum += S[0] * w[12] + S[cn] * w[13] + S[cn * 2] * w[14] + S[cn * 3] * w[15];
          S += 1 - sstep * 3;
          D[k] = saturateU16FromInt(cvRoundF(sum));
        }
      } else {
        if (sx >= sw || sx + 4 <= 0 || sy >= sh || sy + 4 <= 0) {
          D[0] = D[1] = D[2] = 0;
          continue;
        }
        int x[4], y[4];
        for (int i = 0; i < 4; ++i) {
          x[i] = ((unsigned)(sx + i) < (unsigned)sw) ? (sx + i) * cn : -cn;
          y[i] = ((unsigned)(sy + i) < (unsigned)sh) ? (sy + i) : -1;
        }
        for (int k = 0; k < cn; ++k) {
          const float cv = 0.f;
          float sum = cv * 1;
          const float* ww = w;
          for (int i = 0; i < 4; ++i, ww += 4) {
            const int yi = y[i];
            if (yi < 0) continue;
            const uint16_t* S = src + (size_t)yi * sstep + k;
            if (x[0] >= 0) sum += (S[x[0]] - cv) * ww[0];
            if (x[1] >= 0) sum += (S[x[1]] - cv) * ww[1];
            if (x[2] >= 0) sum += (S[x[2]] - cv) * ww[2];
            if (x[3] >= 0) sum += (S[x[3]] - cv) * ww[3];
          }
          D[k] = saturateU16FromInt(cvRoundF(sum));
        }
      }
    }
  }
}

inline int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * len - 2 - p;
  }
  return p;
}

// ---- cv::blur 3x3 on u16x3 (box_filter: int sums, ColumnSum<int,ushort> = saturate(round(s/9))) ----
inline void blur3x3U16C3(const uint16_t* src, int w, int h, uint16_t* dst, int y0 = 0, int y1 = -1) {
  if (y1 < 0) y1 = h;
  const double scale = 1.0 / 9;
  for (int y = y0; y < y1; ++y) {
    const int ys[3] = {reflect101(y - 1, h), y, reflect101(y + 1, h)};
    for (int x = 0; x < w; ++x) {
      const int xs[3] = {reflect101(x - 1, w), x, reflect101(x + 1, w)};
      for (int c = 0; c < 3; ++c) {
        int s = 0;
        for (int j = 0; j < 3; ++j)
          for (int i = 0; i < 3; ++i) s += src[((size_t)ys[j] * w + xs[i]) * 3 + c];
        dst[((size_t)y * w + x) * 3 + c] = saturateU16FromInt(cvRoundD(s * scale));
      }
    }
  }
}

// ---- computeImageVariance (DerpUtil.cpp:214-237) ---------------------------------------------
// imageF = u16 * (1/65535.f) [convertTo, float math]; mean = blur(imageF); msq = blur(imageF^2);
// var_c = msq - mean*mean; var = var_B*w[2] + var_G*w[1] + var_R*w[0], w = {.3333,.3334,.3333}.
// cv::blur on CV_32F accumulates in double (sumType CV_64F) and scales by 1/9 in double.
inline void imageVarianceU16C3(const uint16_t* src, int w, int h, float* var, int y0 = 0, int y1 = -1) {
  if (y1 < 0) y1 = h;
  const float alpha = 1.0f / 65535.0f;
  const double scale = 1.0 / 9;
  const float kw[3] = {0.3333f, 0.3334f, 0.3333f};
  for (int y = y0; y < y1; ++y) {
    const int ys[3] = {reflect101(y - 1, h), y, reflect101(y + 1, h)};
    for (int x = 0; x < w; ++x) {
      const int xs[3] = {reflect101(x - 1, w), x, reflect101(x + 1, w)};
      float vc[3];
      for (int c = 0; c < 3; ++c) {
        double s = 0, s2 = 0;
        for (int j = 0; j < 3; ++j) {
          double rs = 0, rs2 = 0;  // RowSum then ColumnSum
          for (int i = 0; i < 3; ++i) {
            const float f = src[((size_t)ys[j] * w + xs[i]) * 3 + c] * alpha;
            const float f2 = f * f;
            rs += (double)f;
            rs2 += (double)f2;
          }
          s += rs;
          s2 += rs2;
        }
        const float mean = (float)(s * scale);
        const float msq = (float)(s2 * scale);
        vc[c] = msq - mean * mean;
      }
      var[(size_t)y * w + x] = vc[0] * kw[2] + vc[1] * kw[1] + vc[2] * kw[0];
    }
  }
}

// ---- cv::resize INTER_LANCZOS4, float 1ch (resize.cpp: interpolateLanczos4, HResizeLanczos4,
// VResizeLanczos4) ----------------------------------------------------------------------------
inline void lanczos4Coeffs(float x, float* coeffs) {
  static const double s45 = 0.70710678118654752440084436210485;
  static const double cs[][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45},
                                 {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
  if (x < 1.1920928955078125e-07f) {
    for (int i = 0; i < 8; ++i) coeffs[i] = 0;
    coeffs[3] = 1;
    return;
  }
  float sum = 0;
  const double y0 = -(x + 3) * M_PI * 0.25, s0 = std::sin(y0), c0 = std::cos(y0);
  for (int i = 0; i < 8; ++i) {
    const double y = -(x + 3 - i) * M_PI * 0.25;
    coeffs[i] = (float)((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
    sum += coeffs[i];
  }
  sum = 1.f / sum;
  for (int i = 0; i < 8; ++i) coeffs[i] *= sum;
}

inline void resizeLanczos4F32(const float* src, int sw, int sh, float* dst, int dw, int dh) {
  const double inv_x = (double)dw / sw, inv_y = (double)dh / sh;
  const double scale_x = 1. / inv_x, scale_y = 1. / inv_y;
  std::vector<int> xofs(dw), yofs(dh);
  std::vector<float> alpha((size_t)dw * 8), beta((size_t)dh * 8);
  for (int dx = 0; dx < dw; ++dx) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    const int sx = cvFloorD(fx);
    fx -= sx;
    xofs[dx] = sx;
    lanczos4Coeffs(fx, &alpha[(size_t)dx * 8]);
  }
  for (int dy = 0; dy < dh; ++dy) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    const int sy = cvFloorD(fy);
    fy -= sy;
    yofs[dy] = sy;
    lanczos4Coeffs(fy, &beta[(size_t)dy * 8]);
  }
  // horizontal pass of the 8 needed rows, then vertical
  std::vector<float> rows((size_t)sh * dw);
  for (int y = 0; y < sh; ++y) {
    const float* S = src + (size_t)y * sw;
    float* D = &rows[(size_t)y * dw];
    for (int dx = 0; dx < dw; ++dx) {
      const int sx = xofs[dx];
      const float* a = &alpha[(size_t)dx * 8];
      if (sx - 3 >= 0 && sx + 4 < sw) {
        D[dx] = S[sx - 3] * a[0] + S[sx - 2] * a[1] + S[sx - 1] * a[2] + S[sx] * a[3] +
            S[sx + 1] * a[4] + S[sx + 2] * a[5] + S[sx + 3] * a[6] + S[sx + 4] * a[7];
      } else {
        float v = 0;
        for (int j = 0; j < 8; ++j) {
          int sxj = sx + j - 3;
          sxj = sxj < 0 ? 0 : sxj >= sw ? sw - 1 : sxj;
          v += S[sxj] * a[j];
        }
        D[dx] = v;
      }
    }
  }
  for (int dy = 0; dy < dh; ++dy) {
    const float* b = &beta[(size_t)dy * 8];
    const float* R[8];
    for (int k = 0; k < 8; ++k) {
      int sy = yofs[dy] - 3 + k;
      sy = sy < 0 ? 0 : sy >= sh ? sh - 1 : sy;
      R[k] = &rows[(size_t)sy * dw];
    }
    float* D = dst + (size_t)dy * dw;
    for (int x = 0; x < dw; ++x)
      D[x] = R[0][x] * b[0] + R[1][x] * b[1] + R[2][x] * b[2] + R[3][x] * b[3] + R[4][x] * b[4] +
          R[5][x] * b[5] + R[6][x] * b[6] + R[7][x] * b[7];
  }
}

// ---- cv::resize INTER_NEAREST (resize.cpp resizeNN: sx = min(floor(x * ifx), sw-1)) -------------
template <typename T>
inline void resizeNearest(const T* src, int sw, int sh, T* dst, int dw, int dh) {
  const double inv_x = (double)dw / sw, inv_y = (double)dh / sh;
  const double ifx = 1. / inv_x, ify = 1. / inv_y;
  std::vector<int> xofs(dw);
  for (int x = 0; x < dw; ++x) xofs[x] = std::min(cvFloorD(x * ifx), sw - 1);
  for (int y = 0; y < dh; ++y) {
    const int sy = std::min(cvFloorD(y * ify), sh - 1);
    for (int x = 0; x < dw; ++x) dst[(size_t)y * dw + x] = src[(size_t)sy * sw + xofs[x]];
  }
}

// cv::resize(src, dst, Size(), fx, fy, INTER_NEAREST): dsize = cvRound(size * f), source index =
// min(floor(d * (1 / f)), size - 1) — the scale factors the caller gave, not the size ratio (resize.cpp).
// ConvertToBinary.cpp:153-156 shrinks the depth map with it.
inline void nearestScaledAxis(int sn, double f, std::vector<int>& ofs) {
  const int dn = cvRoundD(sn * f);
  const double ifx = 1. / f;
  ofs.resize(dn > 0 ? dn : 0);
  for (int d = 0; d < dn; ++d) ofs[d] = std::min(cvFloorD(d * ifx), sn - 1);
}
template <typename T>
inline void resizeNearestScaled(const T* src, int sw, int sh, double fx, double fy, std::vector<T>& dst, int* dw, int* dh) {
  std::vector<int> xo, yo;
  nearestScaledAxis(sw, fx, xo);
  nearestScaledAxis(sh, fy, yo);
  *dw = (int)xo.size();
  *dh = (int)yo.size();
  dst.resize(xo.size() * yo.size());
  for (size_t y = 0; y < yo.size(); ++y)
    for (size_t x = 0; x < xo.size(); ++x) dst[y * xo.size() + x] = src[(size_t)yo[y] * sw + xo[x]];
}

// ---- cv::resize INTER_AREA, u16 x 3 channels, shrinking (resize.cpp: resizeAreaFast_ for integer ratios,
// computeResizeAreaTab + ResizeArea_Invoker<ushort, float> otherwise) — scripts/render/resize.py:79 builds every
// pyramid level from the full-size image with it, UpsampleDisparity.cpp:117 (cv_util::resizeImage) shrinks colour.
struct AreaTap {
  int si, di;
  float alpha;
};
inline void resizeAreaTab(int ssize, int dsize, double scale, std::vector<AreaTap>& tab) {
  tab.clear();
  for (int dx = 0; dx < dsize; ++dx) {
    const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
    const double cellWidth = std::min(scale, ssize - fsx1);
    int sx1 = (int)std::ceil(fsx1), sx2 = (int)std::floor(fsx2);
    sx2 = std::min(sx2, ssize - 1);
    sx1 = std::min(sx1, sx2);
    if (sx1 - fsx1 > 1e-3) tab.push_back({sx1 - 1, dx, (float)((sx1 - fsx1) / cellWidth)});
    for (int sx = sx1; sx < sx2; ++sx) tab.push_back({sx, dx, float(1.0 / cellWidth)});
    if (fsx2 - sx2 > 1e-3) tab.push_back({sx2, dx, (float)(std::min(std::min(fsx2 - sx2, 1.), cellWidth) / cellWidth)});
  }
}
// true when both ratios are integers (resize.cpp: is_area_fast)
inline bool areaIsFast(int sw, int sh, int dw, int dh, int* kx, int* ky) {
  const double sx = (double)sw / dw, sy = (double)sh / dh;
  *kx = (int)std::floor(sx + 0.5);
  *ky = (int)std::floor(sy + 0.5);
  return std::fabs(sx - *kx) < 2.220446049250313e-16 && std::fabs(sy - *ky) < 2.220446049250313e-16;
}
inline bool resizeAreaU16C3(const uint16_t* src, int sw, int sh, uint16_t* dst, int dw, int dh) {
  if (dw > sw || dh > sh) return false;  // INTER_AREA enlarging is a different code path (not on the depth path)
  const int cn = 3;
  int kx, ky;
  if (areaIsFast(sw, sh, dw, dh, &kx, &ky)) {
    const int area = kx * ky;
    const float scale = 1.f / area;
    for (int y = 0; y < dh; ++y)
      for (int x = 0; x < dw; ++x)
        for (int c = 0; c < cn; ++c) {
          const uint16_t* S = src + ((size_t)y * ky * sw + (size_t)x * kx) * cn + c;
          if (kx == 2 && ky == 2) {  // ResizeAreaFastVec<ushort>: integer mean with rounding
            dst[((size_t)y * dw + x) * cn + c] = (uint16_t)((S[0] + S[cn] + S[(size_t)sw * cn] + S[(size_t)sw * cn + cn] + 2) >> 2);
          } else {
            float sum = 0;
            for (int j = 0; j < ky; ++j)
              for (int i = 0; i < kx; ++i) sum += S[((size_t)j * sw + i) * cn];
            dst[((size_t)y * dw + x) * cn + c] = saturateU16FromInt(cvRoundF(sum * scale));
          }
        }
    return true;
  }
  std::vector<AreaTap> xtab, ytab;
  resizeAreaTab(sw, dw, (double)sw / dw, xtab);
  resizeAreaTab(sh, dh, (double)sh / dh, ytab);
  std::vector<float> buf((size_t)dw * cn), sum((size_t)dw * cn, 0.f);
  int prev_dy = ytab.empty() ? 0 : ytab[0].di;
  for (size_t j = 0; j < ytab.size(); ++j) {
    const float beta = ytab[j].alpha;
    const int dy = ytab[j].di, sy = ytab[j].si;
    const uint16_t* S = src + (size_t)sy * sw * cn;
    std::fill(buf.begin(), buf.end(), 0.f);
    for (const AreaTap& t : xtab)
      for (int c = 0; c < cn; ++c) buf[(size_t)t.di * cn + c] = buf[(size_t)t.di * cn + c] + S[(size_t)t.si * cn + c] * t.alpha;
    if (dy != prev_dy) {
      for (size_t i = 0; i < sum.size(); ++i) {
        dst[(size_t)prev_dy * dw * cn + i] = saturateU16FromInt(cvRoundF(sum[i]));
        sum[i] = beta * buf[i];
      }
      prev_dy = dy;
    } else {
      for (size_t i = 0; i < sum.size(); ++i) sum[i] += beta * buf[i];
    }
  }
  for (size_t i = 0; i < sum.size(); ++i) dst[(size_t)prev_dy * dw * cn + i] = saturateU16FromInt(cvRoundF(sum[i]));
  return true;
}

// ---- GenerateForegroundMasks (source/render/BackgroundSubtractionUtil.h:20-59) -------------------------------------
// cv::GaussianBlur(3 x 3, sigma 0) on u16 x 3: OpenCV's fixed-point path with the table kernel (1/4, 1/2, 1/4), i.e.
// (sum of (1 2 1; 2 4 2; 1 2 1) weights + 8) >> 4, BORDER_REFLECT_101 (verified against cv2 4.13).
inline void gaussian3U16C3(const uint16_t* src, int w, int h, uint16_t* dst) {
  static const int k[3] = {1, 2, 1};
  for (int y = 0; y < h; ++y) {
    const int ys[3] = {reflect101(y - 1, h), y, reflect101(y + 1, h)};
    for (int x = 0; x < w; ++x) {
      const int xs[3] = {reflect101(x - 1, w), x, reflect101(x + 1, w)};
      for (int c = 0; c < 3; ++c) {
        unsigned s = 0;
        for (int j = 0; j < 3; ++j)
          for (int i = 0; i < 3; ++i) s += (unsigned)(k[j] * k[i]) * src[((size_t)ys[j] * w + xs[i]) * 3 + c];
        dst[((size_t)y * w + x) * 3 + c] = (uint16_t)((s + 8u) >> 4);
      }
    }
  }
}
// cv::dilate / cv::erode with a k x k rectangle, anchor at its centre (k / 2), constant border that never wins
inline void morphRectU8(const uint8_t* src, int w, int h, int k, bool dilate, uint8_t* dst) {
  const int a = k / 2;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      uint8_t v = dilate ? 0 : 255;
      for (int j = -a; j < k - a; ++j)
        for (int i = -a; i < k - a; ++i) {
          const int yy = y + j, xx = x + i;
          if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
          const uint8_t t = src[(size_t)yy * w + xx];
          v = dilate ? std::max(v, t) : std::min(v, t);
        }
      dst[(size_t)y * w + x] = v;
    }
}
// generateForegroundMask<Vec3w, Vec3f>: blur both images, convert to float in [0, 1], mask = ||template - frame||_2 >
// threshold (cv::norm accumulates in double), morphological closing with a `closing` x `closing` rectangle.
inline bool foregroundMaskU16C3(const uint16_t* templ, const uint16_t* frame, int w, int h, int blurRadius, float threshold,
                                int closing, uint8_t* mask) {
  if (blurRadius < 0 || blurRadius > 1 || closing < 0) return false;  // radius 1 = the reference's default; larger kernels are not restated
  const size_t n = (size_t)w * h;
  std::vector<uint16_t> tb, fb;
  if (blurRadius == 1) {
    tb.resize(n * 3);
    fb.resize(n * 3);
    gaussian3U16C3(templ, w, h, tb.data());
    gaussian3U16C3(frame, w, h, fb.data());
    templ = tb.data();
    frame = fb.data();
  }
  const float alpha = 1.0f / 65535.0f;
  std::vector<uint8_t> m(n);
  for (size_t i = 0; i < n; ++i) {
    double s = 0;
    for (int c = 0; c < 3; ++c) {
      const float d = std::fabs(templ[i * 3 + c] * alpha - frame[i * 3 + c] * alpha);
      s += (double)d * (double)d;
    }
    m[i] = std::sqrt(s) > (double)threshold ? 1 : 0;
  }
  if (closing > 0) {
    std::vector<uint8_t> t(n);
    morphRectU8(m.data(), w, h, closing, true, t.data());
    morphRectU8(t.data(), w, h, closing, false, m.data());
  }
  std::memcpy(mask, m.data(), n);
  return true;
}

}  // namespace oracle
