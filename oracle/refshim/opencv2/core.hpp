// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// Minimal stand-in for the part of OpenCV's C++ API that the reference's depth path uses, so that
// the reference's OWN sources (source/depth_estimation/{Derp,DerpUtil,UpsampleDisparityLib}.cpp,
// source/util/{Camera,CvUtil,ImageUtil}.cpp and the headers they include) compile unmodified from
// /root/reference into oracle/_ref/libderp_ref.so (recipe: oracle/Makefile, target `ref`).
// OpenCV itself is a third-party dependency that is not under /root/reference and not installed in
// this image as a C++ library.  What is restated here:
//   * cv::Mat / cv::Mat_<T> / cv::Vec / cv::Size / cv::Point containers with OpenCV's conversion rules
//     (saturate_cast on element conversion, Vec(v0) sets channel 0 only, comparisons yield 0/255);
//   * the five numeric primitives of the path (remap INTER_CUBIC, blur, resize LANCZOS4 / NEAREST, float
//     box filter, dilate) through oracle/cvprims.h, which is pinned to cv2 4.13 outputs
//     (tests/test_oracle_cv.py);
//   * everything else the headers merely mention (imread, cvtColor, GaussianBlur, ...) as stubs that throw.
// Nothing here is reference code; nothing here is used by the product.
#pragma once

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../cvprims.h"

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_MAT_DEPTH(t) ((t)&7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 63) + 1)
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_16UC3 CV_MAKETYPE(CV_16U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_32FC4 CV_MAKETYPE(CV_32F, 4)

namespace cv {

[[noreturn]] inline void shimUnsupported(const char* what) {
  throw std::runtime_error(std::string("refshim: cv::") + what + " is not part of the depth path and is not implemented");
}

typedef unsigned char uchar;
typedef unsigned short ushort;

enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3, INTER_LANCZOS4 = 4 };
enum BorderTypes { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };
enum ImreadModes { IMREAD_UNCHANGED = -1, IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1 };
enum ColorConversionCodes {
  COLOR_BGR2BGRA = 0,
  COLOR_BGRA2BGR = 1,
  COLOR_BGR2RGBA = 2,
  COLOR_BGRA2RGBA = 5,
  COLOR_BGR2GRAY = 6,
  COLOR_GRAY2BGR = 8,
  COLOR_GRAY2BGRA = 9,
  COLOR_BGRA2GRAY = 10,
  COLOR_BGR2Lab = 44
};
enum ThresholdTypes { THRESH_BINARY = 0 };
enum NormTypes { NORM_L2 = 4 };
enum MorphShapes { MORPH_RECT = 0 };

// ---- saturate_cast (core/saturate.hpp): float -> integer rounds to nearest even (cvRound), then clamps ----------
template <class T>
inline T saturate_cast(uchar v) { return T(v); }
template <class T>
inline T saturate_cast(ushort v) { return T(v); }
template <class T>
inline T saturate_cast(int v) { return T(v); }
template <class T>
inline T saturate_cast(unsigned v) { return T(v); }
template <class T>
inline T saturate_cast(float v) { return T(v); }
template <class T>
inline T saturate_cast(double v) { return T(v); }
template <class T>
inline T saturate_cast(bool v) { return T(v); }
template <>
inline uchar saturate_cast<uchar>(int v) { return (uchar)((unsigned)v <= 255 ? v : v > 0 ? 255 : 0); }
template <>
inline uchar saturate_cast<uchar>(ushort v) { return (uchar)std::min<unsigned>(v, 255u); }
template <>
inline uchar saturate_cast<uchar>(float v) { return saturate_cast<uchar>((int)lrintf(v)); }
template <>
inline uchar saturate_cast<uchar>(double v) { return saturate_cast<uchar>((int)lrint(v)); }
template <>
inline ushort saturate_cast<ushort>(int v) { return (ushort)((unsigned)v <= 65535u ? v : v > 0 ? 65535 : 0); }
template <>
inline ushort saturate_cast<ushort>(float v) { return saturate_cast<ushort>((int)lrintf(v)); }
template <>
inline ushort saturate_cast<ushort>(double v) { return saturate_cast<ushort>((int)lrint(v)); }
template <>
inline int saturate_cast<int>(float v) { return (int)lrintf(v); }
template <>
inline int saturate_cast<int>(double v) { return (int)lrint(v); }
template <>
inline bool saturate_cast<bool>(float v) { return v != 0; }
template <>
inline bool saturate_cast<bool>(double v) { return v != 0; }

// ---- Size / Point ------------------------------------------------------------------------------------------
template <class T>
struct Size_ {
  T width, height;
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
  bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
  bool operator!=(const Size_& o) const { return !(*this == o); }
  T area() const { return width * height; }
};
typedef Size_<int> Size;
template <class T>
inline std::ostream& operator<<(std::ostream& os, const Size_<T>& s) {
  return os << "[" << s.width << " x " << s.height << "]";
}
template <class T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T xx, T yy) : x(xx), y(yy) {}
};
typedef Point_<int> Point;

// ---- Vec (core/matx.hpp) -----------------------------------------------------------------------------------
template <class T, int N>
struct Vec {
  typedef T value_type;
  enum { channels = N, rows = N, cols = 1 };
  T val[N];
  Vec() {
    for (int i = 0; i < N; ++i) val[i] = T(0);
  }
  Vec(T v0) : Vec() { val[0] = v0; }  // Matx(_Tp v0): the remaining channels stay 0
  Vec(T v0, T v1) : Vec() {
    static_assert(N >= 2, "");
    val[0] = v0;
    val[1] = v1;
  }
  Vec(T v0, T v1, T v2) : Vec() {
    static_assert(N >= 3, "");
    val[0] = v0;
    val[1] = v1;
    val[2] = v2;
  }
  Vec(T v0, T v1, T v2, T v3) : Vec() {
    static_assert(N >= 4, "");
    val[0] = v0;
    val[1] = v1;
    val[2] = v2;
    val[3] = v3;
  }
  template <class U>
  Vec(const Vec<U, N>& o) {  // Matx::operator Matx<T2,m,n>() : saturate_cast per element
    for (int i = 0; i < N; ++i) val[i] = saturate_cast<T>(o.val[i]);
  }
  const T& operator[](int i) const { return val[i]; }
  T& operator[](int i) { return val[i]; }
  const T& operator()(int i) const { return val[i]; }
  T& operator()(int i) { return val[i]; }
  T dot(const Vec& o) const {  // Matx::dot: s = 0; s += a[i]*b[i]
    T s = 0;
    for (int i = 0; i < N; ++i) s += val[i] * o.val[i];
    return s;
  }
  static Vec all(T v) {
    Vec r;
    for (int i = 0; i < N; ++i) r.val[i] = v;
    return r;
  }
};
template <class T, int N>
inline Vec<T, N> operator+(const Vec<T, N>& a, const Vec<T, N>& b) {
  Vec<T, N> r;
  for (int i = 0; i < N; ++i) r.val[i] = saturate_cast<T>(a.val[i] + b.val[i]);
  return r;
}
template <class T, int N>
inline Vec<T, N> operator-(const Vec<T, N>& a, const Vec<T, N>& b) {
  Vec<T, N> r;
  for (int i = 0; i < N; ++i) r.val[i] = saturate_cast<T>(a.val[i] - b.val[i]);
  return r;
}
template <class T, int N>
inline Vec<T, N>& operator+=(Vec<T, N>& a, const Vec<T, N>& b) {
  for (int i = 0; i < N; ++i) a.val[i] = saturate_cast<T>(a.val[i] + b.val[i]);
  return a;
}
#define REFSHIM_VEC_SCALE(S)                                                           \
  template <class T, int N>                                                            \
  inline Vec<T, N> operator*(const Vec<T, N>& a, S s) {                                \
    Vec<T, N> r;                                                                       \
    for (int i = 0; i < N; ++i) r.val[i] = saturate_cast<T>(a.val[i] * s);             \
    return r;                                                                          \
  }                                                                                    \
  template <class T, int N>                                                            \
  inline Vec<T, N> operator*(S s, const Vec<T, N>& a) { return a * s; }                \
  template <class T, int N>                                                            \
  inline Vec<T, N>& operator*=(Vec<T, N>& a, S s) {                                    \
    for (int i = 0; i < N; ++i) a.val[i] = saturate_cast<T>(a.val[i] * s);             \
    return a;                                                                          \
  }                                                                                    \
  template <class T, int N>                                                            \
  inline Vec<T, N>& operator/=(Vec<T, N>& a, S s) {                                    \
    const S ia = 1 / s; /* matx.hpp: multiplies by the reciprocal */                   \
    for (int i = 0; i < N; ++i) a.val[i] = saturate_cast<T>(a.val[i] * ia);            \
    return a;                                                                          \
  }
REFSHIM_VEC_SCALE(int)
REFSHIM_VEC_SCALE(float)
REFSHIM_VEC_SCALE(double)
#undef REFSHIM_VEC_SCALE
template <class T, int N>
inline bool operator==(const Vec<T, N>& a, const Vec<T, N>& b) {
  for (int i = 0; i < N; ++i)
    if (!(a.val[i] == b.val[i])) return false;
  return true;
}
template <class T, int N>
inline bool operator!=(const Vec<T, N>& a, const Vec<T, N>& b) { return !(a == b); }

typedef Vec<uchar, 3> Vec3b;
typedef Vec<uchar, 4> Vec4b;
typedef Vec<ushort, 3> Vec3w;
typedef Vec<ushort, 4> Vec4w;
typedef Vec<float, 2> Vec2f;
typedef Vec<float, 3> Vec3f;
typedef Vec<float, 4> Vec4f;
typedef Vec<double, 2> Vec2d;
typedef Vec<double, 3> Vec3d;

template <class T, int N>
inline double norm(const Vec<T, N>& a, const Vec<T, N>& b, int /*normType*/) {
  double s = 0;
  for (int i = 0; i < N; ++i) {
    const double d = (double)a.val[i] - (double)b.val[i];
    s += d * d;
  }
  return std::sqrt(s);
}

// ---- element type traits (core/traits.hpp) -----------------------------------------------------------------
template <class T>
struct DataType;
#define REFSHIM_DT(T, D)                            \
  template <>                                       \
  struct DataType<T> {                              \
    typedef T channel_type;                         \
    enum { depth = D, channels = 1 };               \
  };
REFSHIM_DT(bool, CV_8U)
REFSHIM_DT(uchar, CV_8U)
REFSHIM_DT(signed char, CV_8S)
REFSHIM_DT(ushort, CV_16U)
REFSHIM_DT(short, CV_16S)
REFSHIM_DT(int, CV_32S)
REFSHIM_DT(float, CV_32F)
REFSHIM_DT(double, CV_64F)
#undef REFSHIM_DT
template <class T, int N>
struct DataType<Vec<T, N>> {
  typedef T channel_type;
  enum { depth = DataType<T>::depth, channels = N };
};
inline size_t depthBytes(int depth) {
  static const size_t b[7] = {1, 1, 2, 2, 4, 4, 8};
  return b[depth];
}

// ---- Mat ---------------------------------------------------------------------------------------------------
template <class T>
class Mat_;
class Mat {
 public:
  int rows = 0, cols = 0;
  uchar* data = nullptr;
  struct Step {
    size_t p[2] = {0, 0};
    size_t operator[](int i) const { return p[i]; }
    operator size_t() const { return p[0]; }
  } step;

  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(Size s, int type) { create(s.height, s.width, type); }

  int type() const { return type_; }
  int depth() const { return CV_MAT_DEPTH(type_); }
  int channels() const { return CV_MAT_CN(type_); }
  size_t elemSize() const { return depthBytes(depth()) * channels(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  Size size() const { return Size(cols, rows); }
  size_t total() const { return (size_t)rows * cols; }
  bool isContinuous() const { return true; }

  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == type_) return;
    type_ = type;
    rows = r;
    cols = c;
    const size_t bytes = (size_t)r * c * elemSize();
    store_ = std::shared_ptr<uchar>(new uchar[bytes ? bytes : 1], std::default_delete<uchar[]>());
    data = store_.get();
    step.p[0] = (size_t)c * elemSize();
    step.p[1] = elemSize();
  }
  void create(Size s, int type) { create(s.height, s.width, type); }
  void release() {
    store_.reset();
    data = nullptr;
    rows = cols = 0;
    step = Step();
  }
  Mat clone() const {
    Mat m;
    if (empty()) {
      m.type_ = type_;
      return m;
    }
    m.create(rows, cols, type_);
    std::memcpy(m.data, data, (size_t)rows * step.p[0]);
    return m;
  }
  void copyTo(Mat& dst) const {
    if (empty()) {
      dst.release();
      dst.type_ = type_;
      return;
    }
    if (dst.data == data) return;
    dst.create(rows, cols, type_);
    std::memcpy(dst.data, data, (size_t)rows * step.p[0]);
  }
  // copyTo with a mask: elements where mask != 0; an unallocated / mismatching destination is created and zeroed
  void copyTo(Mat& dst, const Mat& mask) const {
    if (mask.empty()) return copyTo(dst);
    if (mask.rows != rows || mask.cols != cols || mask.depth() != CV_8U || mask.channels() != 1)
      throw std::runtime_error("refshim: copyTo mask mismatch");
    if (dst.rows != rows || dst.cols != cols || dst.type_ != type_ || !dst.data) {
      dst.create(rows, cols, type_);
      std::memset(dst.data, 0, (size_t)rows * step.p[0]);
    }
    const size_t es = elemSize();
    for (size_t i = 0, n = total(); i < n; ++i)
      if (mask.data[i]) std::memcpy(dst.data + i * es, data + i * es, es);
  }
  // convertTo (core/convert_scale): float arithmetic for <= 32-bit inputs, saturate_cast to the target
  void convertTo(Mat& dst, int rtype, double alpha = 1, double beta = 0) const;
  Mat mul(const Mat& o) const;

  template <class T>
  T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step.p[0]); }
  template <class T>
  const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step.p[0]); }
  uchar* ptr(int r = 0) { return data + (size_t)r * step.p[0]; }
  const uchar* ptr(int r = 0) const { return data + (size_t)r * step.p[0]; }
  template <class T>
  T& at(int r, int c) { return ptr<T>(r)[c]; }
  template <class T>
  const T& at(int r, int c) const { return ptr<T>(r)[c]; }

 protected:
  int type_ = 0;
  std::shared_ptr<uchar> store_;
  template <class T>
  friend class Mat_;
};

template <class S, class D>
inline void convertElems(const S* s, D* d, size_t n, double alpha, double beta) {
  const bool scale = !(alpha == 1 && beta == 0);
  if (std::is_same<S, double>::value || std::is_same<D, double>::value) {
    for (size_t i = 0; i < n; ++i) d[i] = scale ? saturate_cast<D>((double)s[i] * alpha + beta) : saturate_cast<D>(s[i]);
  } else {
    const float a = (float)alpha, b = (float)beta;
    for (size_t i = 0; i < n; ++i) d[i] = scale ? saturate_cast<D>((float)s[i] * a + b) : saturate_cast<D>(s[i]);
  }
}
template <class S>
inline void convertFrom(const S* s, Mat& dst, int ddepth, size_t n, double alpha, double beta) {
  switch (ddepth) {
    case CV_8U: return convertElems(s, dst.ptr<uchar>(), n, alpha, beta);
    case CV_16U: return convertElems(s, dst.ptr<ushort>(), n, alpha, beta);
    case CV_32S: return convertElems(s, dst.ptr<int>(), n, alpha, beta);
    case CV_32F: return convertElems(s, dst.ptr<float>(), n, alpha, beta);
    case CV_64F: return convertElems(s, dst.ptr<double>(), n, alpha, beta);
    default: shimUnsupported("convertTo to this depth");
  }
}
inline void Mat::convertTo(Mat& dst, int rtype, double alpha, double beta) const {
  const int ddepth = rtype < 0 ? depth() : CV_MAT_DEPTH(rtype);
  if (empty()) {
    dst.release();
    dst.type_ = CV_MAKETYPE(ddepth, channels());
    return;
  }
  Mat out(rows, cols, CV_MAKETYPE(ddepth, channels()));
  const size_t n = total() * channels();
  switch (depth()) {
    case CV_8U: convertFrom(ptr<uchar>(), out, ddepth, n, alpha, beta); break;
    case CV_16U: convertFrom(ptr<ushort>(), out, ddepth, n, alpha, beta); break;
    case CV_32S: convertFrom(ptr<int>(), out, ddepth, n, alpha, beta); break;
    case CV_32F: convertFrom(ptr<float>(), out, ddepth, n, alpha, beta); break;
    case CV_64F: convertFrom(ptr<double>(), out, ddepth, n, alpha, beta); break;
    default: shimUnsupported("convertTo from this depth");
  }
  dst = out;
}

// element-wise float arithmetic of the untyped Mat expressions the path writes (computeRgbVariance, saveDstImage)
inline void requireF32(const Mat& m, const char* what) {
  if (m.depth() != CV_32F) shimUnsupported(what);
}
inline Mat Mat::mul(const Mat& o) const {
  requireF32(*this, "Mat::mul on a non-float Mat");
  Mat r(rows, cols, type_);
  const float *a = ptr<float>(), *b = o.ptr<float>();
  float* d = r.ptr<float>();
  for (size_t i = 0, n = total() * channels(); i < n; ++i) d[i] = a[i] * b[i];
  return r;
}
inline Mat operator-(const Mat& x, const Mat& y) {
  requireF32(x, "Mat - Mat on non-float Mats");
  Mat r(x.rows, x.cols, x.type());
  const float *a = x.ptr<float>(), *b = y.ptr<float>();
  float* d = r.ptr<float>();
  for (size_t i = 0, n = x.total() * x.channels(); i < n; ++i) d[i] = a[i] - b[i];
  return r;
}
inline Mat operator+(const Mat& x, const Mat& y) {
  requireF32(x, "Mat + Mat on non-float Mats");
  Mat r(x.rows, x.cols, x.type());
  const float *a = x.ptr<float>(), *b = y.ptr<float>();
  float* d = r.ptr<float>();
  for (size_t i = 0, n = x.total() * x.channels(); i < n; ++i) d[i] = a[i] + b[i];
  return r;
}
inline Mat operator*(const Mat& x, double s) {
  requireF32(x, "Mat * scalar on a non-float Mat");
  Mat r(x.rows, x.cols, x.type());
  const float* a = x.ptr<float>();
  float* d = r.ptr<float>();
  const float f = (float)s;
  for (size_t i = 0, n = x.total() * x.channels(); i < n; ++i) d[i] = a[i] * f;
  return r;
}
inline Mat operator*(double s, const Mat& x) { return x * s; }

// ---- Mat_<T> -----------------------------------------------------------------------------------------------
template <class T>
class Mat_ : public Mat {
 public:
  typedef T value_type;
  enum { kType = CV_MAKETYPE(DataType<T>::depth, DataType<T>::channels) };
  Mat_() { type_ = kType; }
  Mat_(int r, int c) { Mat::create(r, c, kType); }
  Mat_(int r, int c, const T& v) {
    Mat::create(r, c, kType);
    fill(v);
  }
  explicit Mat_(Size s) { Mat::create(s.height, s.width, kType); }
  Mat_(Size s, const T& v) {
    Mat::create(s.height, s.width, kType);
    fill(v);
  }
  Mat_(const Mat& m) { assignFrom(m); }
  Mat_(const Mat_& m) = default;
  Mat_& operator=(const Mat_& m) = default;
  Mat_& operator=(const Mat& m) {
    assignFrom(m);
    return *this;
  }
  void create(int r, int c) { Mat::create(r, c, kType); }
  void create(Size s) { Mat::create(s.height, s.width, kType); }
  Mat_ clone() const { return Mat_(Mat::clone()); }
  int channels() const { return DataType<T>::channels; }
  int depth() const { return DataType<T>::depth; }
  int type() const { return kType; }

  T& operator()(int r, int c) { return reinterpret_cast<T*>(data)[(size_t)r * cols + c]; }
  const T& operator()(int r, int c) const { return reinterpret_cast<const T*>(data)[(size_t)r * cols + c]; }
  T& operator()(Point p) { return (*this)(p.y, p.x); }
  const T& operator()(Point p) const { return (*this)(p.y, p.x); }

  Mat_& setTo(const T& v) {
    fill(v);
    return *this;
  }
  Mat_& setTo(const T& v, const Mat& mask) {
    if (mask.empty()) return setTo(v);
    T* p = reinterpret_cast<T*>(data);
    for (size_t i = 0, n = total(); i < n; ++i)
      if (mask.data[i]) p[i] = v;
    return *this;
  }
  using Mat::copyTo;

 private:
  void fill(const T& v) {
    T* p = reinterpret_cast<T*>(data);
    for (size_t i = 0, n = total(); i < n; ++i) p[i] = v;
  }
  void assignFrom(const Mat& m) {
    if (m.empty() && m.type() == 0 && m.data == nullptr) {  // default-constructed Mat
      Mat::release();
      type_ = kType;
      return;
    }
    if (m.type() == kType) {
      Mat::operator=(m);
      return;
    }
    if (m.channels() == DataType<T>::channels) {  // Mat_<T>(const Mat&): converts the depth
      Mat t;
      m.convertTo(t, DataType<T>::depth);
      Mat::operator=(t);
      type_ = kType;
      return;
    }
    throw std::runtime_error("refshim: Mat_ from a Mat with a different channel count");
  }
};

// comparisons / logic on masks and float planes.  cv::compare writes 255 for true; the reference then reads those
// bytes through Mat_<bool>::operator() (e.g. `!closerMask(y, x)`, Derp.cpp:250), which is only well defined for
// 0 / 1, so the stand-in writes 1: every use on the path is a zero / non-zero test.
template <class T, class F>
inline Mat_<bool> cmpScalar(const Mat_<T>& a, F f) {
  Mat_<bool> r(a.size());
  const T* p = reinterpret_cast<const T*>(a.data);
  for (size_t i = 0, n = a.total(); i < n; ++i) r.data[i] = f(p[i]) ? 1 : 0;
  return r;
}
inline Mat_<bool> operator<(const Mat_<float>& a, double s) {
  return cmpScalar(a, [s](float v) { return (double)v < s; });
}
inline Mat_<bool> operator>(const Mat_<float>& a, double s) {
  return cmpScalar(a, [s](float v) { return (double)v > s; });
}
inline Mat_<bool> operator==(const Mat_<bool>& a, int s) {
  Mat_<bool> r(a.size());
  for (size_t i = 0, n = a.total(); i < n; ++i) r.data[i] = ((int)a.data[i] == s) ? 1 : 0;
  return r;
}
inline Mat_<bool> operator!=(const Mat_<float>& a, const Mat_<float>& b) {
  Mat_<bool> r(a.size());
  const float *p = a.ptr<float>(), *q = b.ptr<float>();
  for (size_t i = 0, n = a.total(); i < n; ++i) r.data[i] = (p[i] != q[i]) ? 1 : 0;
  return r;
}
// MatOp_Bin::assign: when the right operand has no data the expression is evaluated as `a & Scalar()`, i.e. zeros
// (DerpCLI.cpp:277-296 passes empty foreground masks when masks are not used).
inline Mat_<bool> operator&(const Mat_<bool>& a, const Mat_<bool>& b) {
  Mat_<bool> r(a.size());
  if (b.empty()) return r.setTo(false);
  for (size_t i = 0, n = a.total(); i < n; ++i) r.data[i] = a.data[i] & b.data[i];
  return r;
}
inline Mat_<bool> operator-(int s, const Mat_<bool>& a) {  // `1 - mask`: saturating u8 arithmetic
  Mat_<bool> r(a.size());
  for (size_t i = 0, n = a.total(); i < n; ++i) r.data[i] = saturate_cast<uchar>(s - (int)a.data[i]);
  return r;
}
// Mat_<float> * scalar and sums of such products (computeImageVariance): eager float arithmetic
inline Mat_<float> operator*(const Mat_<float>& a, float s) {
  Mat_<float> r(a.size());
  const float* p = a.ptr<float>();
  float* d = r.ptr<float>();
  for (size_t i = 0, n = a.total(); i < n; ++i) d[i] = p[i] * s;
  return r;
}
inline Mat_<float> operator+(const Mat_<float>& a, const Mat_<float>& b) {
  Mat_<float> r(a.size());
  const float *p = a.ptr<float>(), *q = b.ptr<float>();
  float* d = r.ptr<float>();
  for (size_t i = 0, n = a.total(); i < n; ++i) d[i] = p[i] + q[i];
  return r;
}

inline int countNonZero(const Mat& m) {
  int c = 0;
  for (size_t i = 0, n = m.total() * m.elemSize(); i < n; i += m.elemSize()) {
    bool nz = false;
    for (size_t k = 0; k < m.elemSize(); ++k) nz |= m.data[i + k] != 0;
    c += nz;
  }
  return c;
}
inline void findNonZero(const Mat& m, std::vector<Point>& out) {
  out.clear();
  for (int y = 0; y < m.rows; ++y)
    for (int x = 0; x < m.cols; ++x)
      if (m.data[(size_t)y * m.cols + x]) out.emplace_back(x, y);
}
inline void split(const Mat& m, Mat* mv) {
  const int cn = m.channels();
  const size_t eb = depthBytes(m.depth());
  for (int c = 0; c < cn; ++c) {
    Mat plane(m.rows, m.cols, CV_MAKETYPE(m.depth(), 1));
    for (size_t i = 0, n = m.total(); i < n; ++i) std::memcpy(plane.data + i * eb, m.data + (i * cn + c) * eb, eb);
    mv[c] = plane;
  }
}
inline void extractChannel(const Mat& m, Mat& dst, int coi) {
  const int cn = m.channels();
  const size_t eb = depthBytes(m.depth());
  Mat plane(m.rows, m.cols, CV_MAKETYPE(m.depth(), 1));
  for (size_t i = 0, n = m.total(); i < n; ++i) std::memcpy(plane.data + i * eb, m.data + (i * cn + coi) * eb, eb);
  dst = plane;
}

// ---- imgproc primitives of the path (implementations: oracle/cvprims.h, pinned to cv2 4.13) ----------------
inline void remap(const Mat& src, Mat& dst, const Mat& map1, const Mat& map2, int interpolation, int borderMode = BORDER_CONSTANT) {
  if (src.type() != CV_16UC3 || map1.type() != CV_32FC2 || !map2.empty() || interpolation != INTER_CUBIC || borderMode != BORDER_CONSTANT)
    shimUnsupported("remap other than (u16x3, Vec2f map, INTER_CUBIC, BORDER_CONSTANT)");
  Mat out(map1.rows, map1.cols, CV_16UC3);
  oracle::remapBicubicU16C3(src.ptr<uint16_t>(), src.cols, src.rows, map1.ptr<float>(), map1.cols, map1.rows, out.ptr<uint16_t>());
  dst = out;
}
inline void blur(const Mat& src, Mat& dst, Size ksize) {
  if (ksize.width != 3 || ksize.height != 3) shimUnsupported("blur with a kernel other than 3x3");
  Mat out(src.rows, src.cols, src.type());
  if (src.type() == CV_16UC3) {
    oracle::blur3x3U16C3(src.ptr<uint16_t>(), src.cols, src.rows, out.ptr<uint16_t>());
  } else if (src.depth() == CV_32F) {
    // box_filter: RowSum<float,double> then ColumnSum<double,float> scaled by 1/9 in double
    const int cn = src.channels(), w = src.cols, h = src.rows;
    const float* s = src.ptr<float>();
    float* d = out.ptr<float>();
    const double scale = 1.0 / 9;
    for (int y = 0; y < h; ++y) {
      const int ys[3] = {oracle::reflect101(y - 1, h), y, oracle::reflect101(y + 1, h)};
      for (int x = 0; x < w; ++x) {
        const int xs[3] = {oracle::reflect101(x - 1, w), x, oracle::reflect101(x + 1, w)};
        for (int c = 0; c < cn; ++c) {
          double sum = 0;
          for (int j = 0; j < 3; ++j) {
            double rs = 0;
            for (int i = 0; i < 3; ++i) rs += (double)s[((size_t)ys[j] * w + xs[i]) * cn + c];
            sum += rs;
          }
          d[((size_t)y * w + x) * cn + c] = (float)(sum * scale);
        }
      }
    }
  } else {
    shimUnsupported("blur on this element type");
  }
  dst = out;
}
inline void resize(const Mat& src, Mat& dst, Size dsize, double fx, double fy, int interpolation) {
  if (dsize.width == 0 && dsize.height == 0) {  // scale factors instead of a size (ConvertToBinary.cpp:155)
    if (src.type() != CV_32FC1 || interpolation != INTER_NEAREST) shimUnsupported("resize by factors other than float NEAREST");
    std::vector<float> small;
    int dw = 0, dh = 0;
    oracle::resizeNearestScaled<float>(src.ptr<float>(), src.cols, src.rows, fx, fy, small, &dw, &dh);
    Mat scaled(dh, dw, src.type());
    std::memcpy(scaled.data, small.data(), small.size() * sizeof(float));
    dst = scaled;
    return;
  }
  Mat out(dsize.height, dsize.width, src.type());
  if (src.type() == CV_32FC1 && interpolation == INTER_LANCZOS4) {
    oracle::resizeLanczos4F32(src.ptr<float>(), src.cols, src.rows, out.ptr<float>(), dsize.width, dsize.height);
  } else if (src.type() == CV_32FC1 && interpolation == INTER_NEAREST) {
    oracle::resizeNearest(src.ptr<float>(), src.cols, src.rows, out.ptr<float>(), dsize.width, dsize.height);
  } else if (src.type() == CV_8UC1 && interpolation == INTER_NEAREST) {
    oracle::resizeNearest(src.ptr<uchar>(), src.cols, src.rows, out.ptr<uchar>(), dsize.width, dsize.height);
  } else {
    shimUnsupported("resize other than float LANCZOS4 / NEAREST");
  }
  dst = out;
}
// UpsampleDisparityLib.cpp:131 passes a const Mat_ as the OUTPUT of cv::resize (a reference bug that only compiles
// because OutputArray binds to const Mat&); OpenCV would throw at run time, so does this overload.
inline void resize(const Mat&, const Mat&, Size, double, double, int) { shimUnsupported("resize into a const Mat"); }
inline Mat getStructuringElement(int /*shape*/, Size ksize) {
  Mat m(ksize.height, ksize.width, CV_8UC1);
  std::memset(m.data, 1, m.total());
  return m;
}
inline void dilate(const Mat& src, Mat& dst, const Mat& kernel) {
  if (src.type() != CV_8UC1 || kernel.rows != 3 || kernel.cols != 3) shimUnsupported("dilate other than 3x3 on u8");
  Mat out(src.rows, src.cols, src.type());
  for (int y = 0; y < src.rows; ++y)
    for (int x = 0; x < src.cols; ++x) {
      uchar m = 0;  // BORDER_CONSTANT with the morphology default border value: outside taps do not contribute
      for (int j = -1; j <= 1; ++j)
        for (int i = -1; i <= 1; ++i) {
          const int yy = y + j, xx = x + i;
          if (yy < 0 || yy >= src.rows || xx < 0 || xx >= src.cols) continue;
          m = std::max(m, src.data[(size_t)yy * src.cols + xx]);
        }
      out.data[(size_t)y * src.cols + x] = m;
    }
  dst = out;
}

// ---- mentioned by the headers, never reached on the depth path ----------------------------------------------
inline void cvtColor(const Mat&, Mat&, int) { shimUnsupported("cvtColor"); }
inline double threshold(const Mat& src, Mat& dst, double thresh, double maxval, int /*type*/) {
  if (src.depth() != CV_8U) shimUnsupported("threshold on a non-u8 Mat");
  Mat out(src.rows, src.cols, src.type());
  for (size_t i = 0, n = src.total() * src.channels(); i < n; ++i) out.data[i] = src.data[i] > thresh ? (uchar)maxval : 0;
  dst = out;
  return thresh;
}
inline void GaussianBlur(const Mat&, Mat&, Size, double, double = 0) { shimUnsupported("GaussianBlur"); }
inline void hconcat(const Mat&, const Mat&, Mat&) { shimUnsupported("hconcat"); }
inline void vconcat(const Mat&, const Mat&, Mat&) { shimUnsupported("vconcat"); }
inline Mat imread(const std::string&, int = IMREAD_COLOR) { shimUnsupported("imread"); }
inline bool imwrite(const std::string&, const Mat&, const std::vector<int>& = std::vector<int>()) { shimUnsupported("imwrite"); }

}  // namespace cv
