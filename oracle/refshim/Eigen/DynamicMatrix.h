// ORACLE — TEST INFRASTRUCTURE ONLY.   -*- C++ -*-
//
// Second part of the Eigen stand-in (see Geometry): the run-time sized matrices source/render/MeshUtil.h works on
// (MatrixXd vertexes, MatrixXi faces), their row views, 2 x 2 matrices and the typedef names.  Only element access,
// row copies and casts are on the path the tests call (getVertexesEquiError, getFaces, applyMaskToVertexesAndFaces,
// writeDepth's float / uint32 casts); the remaining members exist so that the header's other inline functions compile.
// Storage is row-major whatever the Options parameter says; data() is only offered for RowMajor types.
#pragma once

#include <cstdint>
#include <memory>

namespace Eigen {

template <class S>
struct RowRef {
  S* p;
  Index n;
  Index size() const { return n; }
  S& operator()(Index i) const { return p[i]; }
  S& operator[](Index i) const { return p[i]; }
  S& x() const { return p[0]; }
  S& y() const { return p[1]; }
  S& z() const { return p[2]; }
  S norm() const {
    S s = p[0] * p[0];
    for (Index i = 1; i < n; ++i) s = s + p[i] * p[i];
    return std::sqrt(s);
  }
  template <int K>
  operator Matrix<S, K, 1>() const {
    Matrix<S, K, 1> r;
    for (int i = 0; i < K; ++i) r.v[i] = p[i];
    return r;
  }
  const RowRef& operator=(const RowRef& o) const {  // copies the coefficients, like Eigen's block assignment
    for (Index i = 0; i < n; ++i) p[i] = o.p[i];
    return *this;
  }
  template <int K>
  const RowRef& operator=(const Matrix<S, K, 1>& o) const {
    for (int i = 0; i < K; ++i) p[i] = o.v[i];
    return *this;
  }
  template <int K>
  typename Matrix<S, K, 1>::template Head<K> head() const {
    return typename Matrix<S, K, 1>::template Head<K>{p};
  }
};

template <class S>
inline typename Matrix<S, 3, 3>::Row& Matrix<S, 3, 3>::Row::operator=(const RowRef<S>& o) {
  p[0] = o.p[0];
  p[1] = o.p[1];
  p[2] = o.p[2];
  return *this;
}

template <class S, int O>
class Matrix<S, Dynamic, Dynamic, O> {
 public:
  Matrix() : r_(0), c_(0) {}
  Matrix(Index rows, Index cols) : v_((size_t)(rows * cols)), r_(rows), c_(cols) {}
  template <int O2>
  Matrix(const Matrix<S, Dynamic, Dynamic, O2>& o) : v_(o.raw()), r_(o.rows()), c_(o.cols()) {}
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Index size() const { return r_ * c_; }
  S& operator()(Index i, Index j) { return v_[(size_t)(i * c_ + j)]; }
  const S& operator()(Index i, Index j) const { return v_[(size_t)(i * c_ + j)]; }
  RowRef<S> row(Index i) { return RowRef<S>{v_.data() + i * c_, c_}; }
  RowRef<S> row(Index i) const { return RowRef<S>{const_cast<S*>(v_.data()) + i * c_, c_}; }
  Matrix topRows(Index n) const {
    Matrix r(n, c_);
    std::copy(v_.begin(), v_.begin() + n * c_, r.v_.begin());
    return r;
  }
  template <class T>
  Matrix<T, Dynamic, Dynamic, O> cast() const {
    Matrix<T, Dynamic, Dynamic, O> r(r_, c_);
    for (size_t i = 0; i < v_.size(); ++i) r.raw()[i] = static_cast<T>(v_[i]);
    return r;
  }
  S* data() {
    static_assert(O == RowMajor, "stand-in stores row-major");
    return v_.data();
  }
  void conservativeResize(NoChange_t, Index cols) {
    std::vector<S> w((size_t)(r_ * cols));
    for (Index i = 0; i < r_; ++i)
      for (Index j = 0; j < std::min(cols, c_); ++j) w[(size_t)(i * cols + j)] = v_[(size_t)(i * c_ + j)];
    v_.swap(w);
    c_ = cols;
  }
  std::vector<S>& raw() { return v_; }
  const std::vector<S>& raw() const { return v_; }

 private:
  std::vector<S> v_;
  Index r_, c_;
};

// ---- 2 x 2 (calcBarycentrics; not on a tested path) ----------------------------------------------------------
template <class S>
class Matrix<S, 2, 2> {
 public:
  typedef Matrix<S, 2, 1> V2;
  S m[2][2];
  S& operator()(Index i, Index j) { return m[i][j]; }
  const S& operator()(Index i, Index j) const { return m[i][j]; }
  struct Row {
    S* p;
    Row& operator-=(const V2& o) {
      p[0] = p[0] - o.v[0];
      p[1] = p[1] - o.v[1];
      return *this;
    }
  };
  Row row(Index i) { return Row{m[i]}; }
  Matrix transpose() const {
    Matrix r;
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j) r.m[i][j] = m[j][i];
    return r;
  }
  struct Solver {  // plain 2 x 2 solve with partial pivoting (Eigen's QR differs in the last bits)
    Matrix a;
    V2 solve(const V2& b) const {
      S A[2][3] = {{a.m[0][0], a.m[0][1], b.v[0]}, {a.m[1][0], a.m[1][1], b.v[1]}};
      if (std::abs(A[1][0]) > std::abs(A[0][0]))
        for (int j = 0; j < 3; ++j) std::swap(A[0][j], A[1][j]);
      const S f = A[1][0] / A[0][0];
      const S y = (A[1][2] - f * A[0][2]) / (A[1][1] - f * A[0][1]);
      const S x = (A[0][2] - A[0][1] * y) / A[0][0];
      return V2(x, y);
    }
  };
  Solver colPivHouseholderQr() const { return Solver{*this}; }
};

// ---- 4 x 4 (the quadrics of source/render/MeshSimplifier.cpp) ---------------------------------------------------
template <class S>
class Matrix<S, 4, 4> {
 public:
  S m[4][4];
  static Matrix Zero() {
    Matrix r;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) r.m[i][j] = 0;
    return r;
  }
  S& operator()(Index i, Index j) { return m[i][j]; }
  const S& operator()(Index i, Index j) const { return m[i][j]; }
  Matrix operator+(const Matrix& o) const {
    Matrix r;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) r.m[i][j] = m[i][j] + o.m[i][j];
    return r;
  }
  Matrix& operator+=(const Matrix& o) {
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) m[i][j] = m[i][j] + o.m[i][j];
    return *this;
  }
  template <int R2, int C2>
  Matrix<S, R2, C2> block(Index i0, Index j0) const {
    Matrix<S, R2, C2> r;
    for (int i = 0; i < R2; ++i)
      for (int j = 0; j < C2; ++j) r(i, j) = m[i0 + i][j0 + j];
    return r;
  }
};
// q * q.transpose(): outer product of a 4-vector with itself (exact products, no sums)
template <class S>
inline Matrix<S, 4, 4> operator*(const Matrix<S, 4, 1>& a, const typename Matrix<S, 4, 1>::Transposed& b) {
  Matrix<S, 4, 4> r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) r.m[i][j] = a.v[i] * b.m->v[j];
  return r;
}
template <class T>
using aligned_allocator = std::allocator<T>;

inline Matrix<double, 3, 1> operator*(float s, const Matrix<double, 3, 1>& m) { return m * (double)s; }

typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<int, 3, 1> Vector3i;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<int, Dynamic, Dynamic> MatrixXi;

}  // namespace Eigen
