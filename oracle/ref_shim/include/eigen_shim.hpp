// TEST INFRASTRUCTURE (oracle/): a minimal stand-in for the part of the Eigen 3 API that the reference's hot-path
// translation units use, so that those UNMODIFIED sources (under /root/reference, never copied) compile here, where
// Eigen is not installed, into oracle/_ref/ (see oracle/ref_shim/Makefile).  It is NOT Eigen and shares no code with
// it: one dense, column-major, run-time sized double matrix with EAGER evaluation (every expression returns a
// matrix), mutable block views, and the few decompositions the reference calls (LLT, LDLT, inverse, Householder QR,
// a Givens rotation, a full-U "SVD" whose null-space columns come from a Householder QR, and a dense QR behind the
// SPQR interface).  Speed is irrelevant: this library is the parity pin of the oracle, never a timing baseline.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <initializer_list>
#include <iostream>
#include <limits>
#include <stdexcept>
#include <type_traits>
#include <utility>
#include <vector>

namespace Eigen {

constexpr int Dynamic = -1;
using Index = long;
enum { ComputeFullU = 1, ComputeThinU = 2, ComputeFullV = 4, ComputeThinV = 8 };

class MatX;
class TransposedX;
class BlockRef;
class LLTShim;
class LDLTShim;
class HouseholderQRShim;
struct SparseShim;
template <typename S> class JacobiRotation;

// ---- comma initialiser:  m << a, b, c;  (row-major fill; scalars or whole matrices side by side) ----
class CommaInit {
 public:
  CommaInit(MatX& m) : m_(m) {}
  CommaInit& operator,(double v);
  CommaInit& operator,(const MatX& b);
  void push(double v);
  void push_block(const MatX& b);

 private:
  MatX& m_;
  long r_ = 0, c_ = 0, blk_rows_ = 1;
};

class MatX {
 public:
  MatX() = default;
  MatX(long r, long c) : r_(r), c_(c), d_((size_t)(r * c), 0.0) {}
  MatX(const MatX&) = default;
  MatX(MatX&&) = default;
  MatX(const BlockRef& b);
  MatX& operator=(const MatX&) = default;
  MatX& operator=(MatX&&) = default;
  MatX& operator=(const BlockRef& b);

  long rows() const { return r_; }
  long cols() const { return c_; }
  long size() const { return r_ * c_; }
  double* data() { return d_.data(); }
  const double* data() const { return d_.data(); }
  double& operator()(long i, long j) { return d_[(size_t)(i + j * r_)]; }
  const double& operator()(long i, long j) const { return d_[(size_t)(i + j * r_)]; }
  double& operator()(long i) { return d_[(size_t)i]; }
  const double& operator()(long i) const { return d_[(size_t)i]; }
  double& operator[](long i) { return d_[(size_t)i]; }
  const double& operator[](long i) const { return d_[(size_t)i]; }
  double& x() { return d_[0]; }
  double& y() { return d_[1]; }
  double& z() { return d_[2]; }
  double& w() { return d_[3]; }
  const double& x() const { return d_[0]; }
  const double& y() const { return d_[1]; }
  const double& z() const { return d_[2]; }
  const double& w() const { return d_[3]; }
  double value() const { return d_[0]; }

  void resize(long r, long c) { r_ = r; c_ = c; d_.assign((size_t)(r * c), 0.0); }
  void resize(long n) { resize(n, c_ == 1 || c_ == 0 ? 1 : c_); if (c_ != 1) { r_ = n; c_ = 1; d_.assign((size_t)n, 0.0); } }
  void conservativeResize(long r, long c) {
    MatX t(r, c);
    for (long j = 0; j < std::min(c, c_); ++j)
      for (long i = 0; i < std::min(r, r_); ++i) t(i, j) = (*this)(i, j);
    *this = std::move(t);
  }
  void conservativeResize(long n) { conservativeResize(n, 1); }

  MatX& setZero() { std::fill(d_.begin(), d_.end(), 0.0); return *this; }
  MatX& setZero(long r, long c) { resize(r, c); return *this; }
  MatX& setZero(long n) { resize(n, 1); return *this; }
  MatX& setOnes() { std::fill(d_.begin(), d_.end(), 1.0); return *this; }
  MatX& setConstant(double v) { std::fill(d_.begin(), d_.end(), v); return *this; }
  MatX& setIdentity() { setZero(); for (long i = 0; i < std::min(r_, c_); ++i) (*this)(i, i) = 1.0; return *this; }
  MatX& setIdentity(long r, long c) { resize(r, c); return setIdentity(); }
  MatX& setRandom() { for (auto& v : d_) v = 2.0 * std::rand() / RAND_MAX - 1.0; return *this; }
  MatX& noalias() { return *this; }
  const MatX& eval() const { return *this; }

  TransposedX transpose() const;
  TransposedX adjoint() const;
  double norm() const { return std::sqrt(squaredNorm()); }
  double squaredNorm() const { double s = 0; for (double v : d_) s += v * v; return s; }
  double sum() const { double s = 0; for (double v : d_) s += v; return s; }
  double trace() const { double s = 0; for (long i = 0; i < std::min(r_, c_); ++i) s += (*this)(i, i); return s; }
  double maxCoeff() const { return *std::max_element(d_.begin(), d_.end()); }
  double minCoeff() const { return *std::min_element(d_.begin(), d_.end()); }
  bool hasNaN() const { for (double v : d_) if (std::isnan(v)) return true; return false; }
  bool allFinite() const { for (double v : d_) if (!std::isfinite(v)) return false; return true; }
  MatX normalized() const { MatX t(*this); t.normalize(); return t; }
  void normalize() { const double n = norm(); if (n > 0) for (auto& v : d_) v /= n; }
  double dot(const MatX& o) const { double s = 0; for (long i = 0; i < size(); ++i) s += d_[(size_t)i] * o.d_[(size_t)i]; return s; }
  MatX cross(const MatX& o) const {
    MatX t(3, 1);
    t(0) = y() * o.z() - z() * o.y(); t(1) = z() * o.x() - x() * o.z(); t(2) = x() * o.y() - y() * o.x();
    return t;
  }
  MatX cwiseAbs() const { MatX t(*this); for (auto& v : t.d_) v = std::fabs(v); return t; }
  MatX cwiseProduct(const MatX& o) const { MatX t(*this); for (long i = 0; i < size(); ++i) t.d_[(size_t)i] *= o.d_[(size_t)i]; return t; }
  MatX diagonal() const { const long n = std::min(r_, c_); MatX t(n, 1); for (long i = 0; i < n; ++i) t(i) = (*this)(i, i); return t; }
  MatX asDiagonal() const { MatX t(size(), size()); for (long i = 0; i < size(); ++i) t(i, i) = d_[(size_t)i]; return t; }
  double determinant() const;
  MatX inverse() const;
  LLTShim llt() const;
  LDLTShim ldlt() const;
  HouseholderQRShim householderQr() const;
  SparseShim sparseView() const;
  bool isApprox(const MatX& o, double prec = 1e-12) const {
    if (r_ != o.r_ || c_ != o.c_) return false;
    double dn = 0, a = 0, b = 0;
    for (long i = 0; i < size(); ++i) { const double e = d_[(size_t)i] - o.d_[(size_t)i]; dn += e * e; a += d_[(size_t)i] * d_[(size_t)i]; b += o.d_[(size_t)i] * o.d_[(size_t)i]; }
    return dn <= prec * prec * std::min(a, b);
  }

  // ---- views (mutable on non-const matrices, copies on const ones) ----
  BlockRef block(long i, long j, long r, long c);
  MatX block(long i, long j, long r, long c) const {
    assert(i >= 0 && j >= 0 && i + r <= r_ && j + c <= c_);
    MatX t(r, c);
    for (long b = 0; b < c; ++b) for (long a = 0; a < r; ++a) t(a, b) = (*this)(i + a, j + b);
    return t;
  }
  template <int R, int C> BlockRef block(long i, long j);
  template <int R, int C> MatX block(long i, long j) const { return block(i, j, R, C); }
  BlockRef segment(long i, long n);
  MatX segment(long i, long n) const { return block(i, 0, n, 1); }
  template <int N> BlockRef segment(long i);
  template <int N> MatX segment(long i) const { return block(i, 0, N, 1); }
  BlockRef head(long n);
  MatX head(long n) const { return block(0, 0, n, 1); }
  template <int N> BlockRef head();
  template <int N> MatX head() const { return block(0, 0, N, 1); }
  BlockRef tail(long n);
  MatX tail(long n) const { return block(r_ - n, 0, n, 1); }
  template <int N> BlockRef tail();
  template <int N> MatX tail() const { return block(r_ - N, 0, N, 1); }
  BlockRef topRows(long n);
  MatX topRows(long n) const { return block(0, 0, n, c_); }
  BlockRef bottomRows(long n);
  MatX bottomRows(long n) const { return block(r_ - n, 0, n, c_); }
  BlockRef leftCols(long n);
  MatX leftCols(long n) const { return block(0, 0, r_, n); }
  template <int N> BlockRef leftCols();
  template <int N> MatX leftCols() const { return block(0, 0, r_, N); }
  BlockRef rightCols(long n);
  MatX rightCols(long n) const { return block(0, c_ - n, r_, n); }
  template <int N> BlockRef rightCols();
  template <int N> MatX rightCols() const { return block(0, c_ - N, r_, N); }
  BlockRef middleCols(long j, long n);
  MatX middleCols(long j, long n) const { return block(0, j, r_, n); }
  template <int N> BlockRef middleCols(long j);
  template <int N> MatX middleCols(long j) const { return block(0, j, r_, N); }
  BlockRef middleRows(long i, long n);
  MatX middleRows(long i, long n) const { return block(i, 0, n, c_); }
  BlockRef row(long i);
  MatX row(long i) const { return block(i, 0, 1, c_); }
  BlockRef col(long j);
  MatX col(long j) const { return block(0, j, r_, 1); }
  BlockRef topLeftCorner(long r, long c);
  MatX topLeftCorner(long r, long c) const { return block(0, 0, r, c); }
  template <int R, int C> BlockRef topLeftCorner();
  template <int R, int C> MatX topLeftCorner() const { return block(0, 0, R, C); }
  template <int N> BlockRef topRows();
  template <int N> MatX topRows() const { return block(0, 0, N, c_); }
  template <int N> BlockRef bottomRows();
  template <int N> MatX bottomRows() const { return block(r_ - N, 0, N, c_); }

  CommaInit operator<<(double v) { CommaInit ci(*this); ci.push(v); return ci; }
  CommaInit operator<<(const MatX& b) { CommaInit ci(*this); ci.push_block(b); return ci; }

  MatX& operator+=(const MatX& o) { assert(r_ == o.r_ && c_ == o.c_); for (long i = 0; i < size(); ++i) d_[(size_t)i] += o.d_[(size_t)i]; return *this; }
  MatX& operator-=(const MatX& o) { assert(r_ == o.r_ && c_ == o.c_); for (long i = 0; i < size(); ++i) d_[(size_t)i] -= o.d_[(size_t)i]; return *this; }
  MatX& operator*=(double s) { for (auto& v : d_) v *= s; return *this; }
  MatX& operator/=(double s) { for (auto& v : d_) v /= s; return *this; }
  MatX& operator*=(const MatX& o);
  MatX operator-() const { MatX t(*this); for (auto& v : t.d_) v = -v; return t; }

  template <typename S> void applyOnTheLeft(long p, long q, const JacobiRotation<S>& g);

 protected:
  long r_ = 0, c_ = 0;
  std::vector<double> d_;
};

// x.transpose() is its own type only so that (row vector) * (vector) can ALSO be read as a scalar, the one place where
// Eigen converts a 1 x 1 product implicitly (Update.cpp:55, :78: `return res.transpose()*S.ldlt().solve(res);`).
class TransposedX : public MatX {
 public:
  TransposedX(long r, long c) : MatX(r, c) {}
  TransposedX operator-() const { TransposedX t(*this); for (auto& v : t.d_) v = -v; return t; }
};
class ScalarProduct : public TransposedX {   // a product with a transposed factor: still "transposed" for the next product
 public:
  ScalarProduct(const MatX& m) : TransposedX(m.rows(), m.cols()) { MatX::operator=(m); }
  operator double() const { assert(r_ == 1 && c_ == 1); return d_[0]; }
};
inline TransposedX MatX::transpose() const {
  TransposedX t(c_, r_);
  for (long j = 0; j < c_; ++j) for (long i = 0; i < r_; ++i) t(j, i) = (*this)(i, j);
  return t;
}
inline TransposedX MatX::adjoint() const { return transpose(); }

inline MatX operator+(const MatX& a, const MatX& b) { MatX t(a); t += b; return t; }
inline MatX operator-(const MatX& a, const MatX& b) { MatX t(a); t -= b; return t; }
inline MatX operator*(const MatX& a, const MatX& b) {
  assert(a.cols() == b.rows());
  MatX t(a.rows(), b.cols());
  for (long j = 0; j < b.cols(); ++j)
    for (long k = 0; k < a.cols(); ++k) {
      const double bkj = b(k, j);
      if (bkj == 0.0) continue;
      for (long i = 0; i < a.rows(); ++i) t(i, j) += a(i, k) * bkj;
    }
  return t;
}
inline ScalarProduct operator*(const TransposedX& a, const MatX& b) {
  return ScalarProduct(static_cast<const MatX&>(a) * b);
}
inline ScalarProduct operator*(const MatX& a, const TransposedX& b) {
  return ScalarProduct(a * static_cast<const MatX&>(b));
}
inline ScalarProduct operator*(const TransposedX& a, const TransposedX& b) {
  return ScalarProduct(static_cast<const MatX&>(a) * static_cast<const MatX&>(b));
}
inline MatX operator*(const MatX& a, double s) { MatX t(a); t *= s; return t; }
inline MatX operator*(double s, const MatX& a) { MatX t(a); t *= s; return t; }
inline MatX operator/(const MatX& a, double s) { MatX t(a); t /= s; return t; }
inline MatX& MatX::operator*=(const MatX& o) { *this = (*this) * o; return *this; }
inline std::ostream& operator<<(std::ostream& os, const MatX& m) {
  for (long i = 0; i < m.rows(); ++i) { for (long j = 0; j < m.cols(); ++j) os << (j ? " " : "") << m(i, j); if (i + 1 < m.rows()) os << "\n"; }
  return os;
}

inline void CommaInit::push(double v) {
  if (c_ >= m_.cols()) { r_ += blk_rows_; c_ = 0; blk_rows_ = 1; }
  m_(r_, c_) = v; ++c_;
}
inline void CommaInit::push_block(const MatX& b) {
  if (c_ >= m_.cols()) { r_ += blk_rows_; c_ = 0; blk_rows_ = 1; }
  for (long j = 0; j < b.cols(); ++j) for (long i = 0; i < b.rows(); ++i) m_(r_ + i, c_ + j) = b(i, j);
  c_ += b.cols(); blk_rows_ = b.rows();
}
inline CommaInit& CommaInit::operator,(double v) { push(v); return *this; }
inline CommaInit& CommaInit::operator,(const MatX& b) { push_block(b); return *this; }

// ---- mutable view of a rectangular part of a MatX ----
class BlockRef {
 public:
  BlockRef(MatX& m, long i, long j, long r, long c) : m_(&m), i_(i), j_(j), r_(r), c_(c) {
    assert(i >= 0 && j >= 0 && r >= 0 && c >= 0 && i + r <= m.rows() && j + c <= m.cols());
  }
  long rows() const { return r_; }
  long cols() const { return c_; }
  long size() const { return r_ * c_; }
  double& operator()(long a, long b) const { return (*m_)(i_ + a, j_ + b); }
  double& operator()(long a) const { return c_ == 1 ? (*m_)(i_ + a, j_) : (*m_)(i_, j_ + a); }
  double& operator[](long a) const { return (*this)(a); }
  double& x() const { return (*this)(0); }
  double& y() const { return (*this)(1); }
  double& z() const { return (*this)(2); }
  MatX eval() const { MatX t(r_, c_); for (long b = 0; b < c_; ++b) for (long a = 0; a < r_; ++a) t(a, b) = (*this)(a, b); return t; }
  BlockRef& assign(const MatX& o) {
    assert(o.rows() == r_ && o.cols() == c_);
    for (long b = 0; b < c_; ++b) for (long a = 0; a < r_; ++a) (*this)(a, b) = o(a, b);
    return *this;
  }
  BlockRef& operator=(const MatX& o) { return assign(o); }
  BlockRef& operator=(const BlockRef& o) { return assign(o.eval()); }   // eval first: the views may overlap
  BlockRef& operator+=(const MatX& o) { return assign(eval() + o); }
  BlockRef& operator-=(const MatX& o) { return assign(eval() - o); }
  BlockRef& operator*=(double s) { return assign(eval() * s); }
  BlockRef& operator/=(double s) { return assign(eval() / s); }
  BlockRef& noalias() { return *this; }
  BlockRef& setZero() { for (long b = 0; b < c_; ++b) for (long a = 0; a < r_; ++a) (*this)(a, b) = 0.0; return *this; }
  BlockRef& setOnes() { for (long b = 0; b < c_; ++b) for (long a = 0; a < r_; ++a) (*this)(a, b) = 1.0; return *this; }
  BlockRef& setIdentity() { setZero(); for (long a = 0; a < std::min(r_, c_); ++a) (*this)(a, a) = 1.0; return *this; }
  BlockRef& setRandom() { MatX t(r_, c_); t.setRandom(); return assign(t); }
  MatX transpose() const { return eval().transpose(); }
  MatX adjoint() const { return transpose(); }
  MatX inverse() const { return eval().inverse(); }
  double norm() const { return eval().norm(); }
  double squaredNorm() const { return eval().squaredNorm(); }
  double sum() const { return eval().sum(); }
  double trace() const { return eval().trace(); }
  bool hasNaN() const { return eval().hasNaN(); }
  MatX normalized() const { return eval().normalized(); }
  double dot(const MatX& o) const { return eval().dot(o); }
  MatX cross(const MatX& o) const { return eval().cross(o); }
  MatX diagonal() const { return eval().diagonal(); }
  MatX operator-() const { return -eval(); }
  BlockRef block(long i, long j, long r, long c) const { return BlockRef(*m_, i_ + i, j_ + j, r, c); }
  template <int R, int C> BlockRef block(long i, long j) const { return block(i, j, R, C); }
  BlockRef head(long n) const { return block(0, 0, n, 1); }
  BlockRef tail(long n) const { return block(r_ - n, 0, n, 1); }
  BlockRef segment(long i, long n) const { return block(i, 0, n, 1); }
  BlockRef topRows(long n) const { return block(0, 0, n, c_); }
  BlockRef bottomRows(long n) const { return block(r_ - n, 0, n, c_); }
  BlockRef leftCols(long n) const { return block(0, 0, r_, n); }
  BlockRef rightCols(long n) const { return block(0, c_ - n, r_, n); }
  BlockRef row(long i) const { return block(i, 0, 1, c_); }
  BlockRef col(long j) const { return block(0, j, r_, 1); }
  CommaInit operator<<(double v);
  template <typename S> void applyOnTheLeft(long p, long q, const JacobiRotation<S>& g) const;

 private:
  MatX* m_;
  long i_, j_, r_, c_;
};

inline MatX::MatX(const BlockRef& b) { *this = b.eval(); }
inline MatX& MatX::operator=(const BlockRef& b) { MatX t = b.eval(); *this = std::move(t); return *this; }
inline BlockRef MatX::block(long i, long j, long r, long c) { return BlockRef(*this, i, j, r, c); }
template <int R, int C> inline BlockRef MatX::block(long i, long j) { return BlockRef(*this, i, j, R, C); }
inline BlockRef MatX::segment(long i, long n) { return BlockRef(*this, i, 0, n, 1); }
template <int N> inline BlockRef MatX::segment(long i) { return BlockRef(*this, i, 0, N, 1); }
inline BlockRef MatX::head(long n) { return BlockRef(*this, 0, 0, n, 1); }
template <int N> inline BlockRef MatX::head() { return BlockRef(*this, 0, 0, N, 1); }
inline BlockRef MatX::tail(long n) { return BlockRef(*this, r_ - n, 0, n, 1); }
template <int N> inline BlockRef MatX::tail() { return BlockRef(*this, r_ - N, 0, N, 1); }
inline BlockRef MatX::topRows(long n) { return BlockRef(*this, 0, 0, n, c_); }
inline BlockRef MatX::bottomRows(long n) { return BlockRef(*this, r_ - n, 0, n, c_); }
inline BlockRef MatX::leftCols(long n) { return BlockRef(*this, 0, 0, r_, n); }
template <int N> inline BlockRef MatX::leftCols() { return BlockRef(*this, 0, 0, r_, N); }
inline BlockRef MatX::rightCols(long n) { return BlockRef(*this, 0, c_ - n, r_, n); }
template <int N> inline BlockRef MatX::rightCols() { return BlockRef(*this, 0, c_ - N, r_, N); }
inline BlockRef MatX::middleCols(long j, long n) { return BlockRef(*this, 0, j, r_, n); }
template <int N> inline BlockRef MatX::middleCols(long j) { return BlockRef(*this, 0, j, r_, N); }
inline BlockRef MatX::middleRows(long i, long n) { return BlockRef(*this, i, 0, n, c_); }
inline BlockRef MatX::row(long i) { return BlockRef(*this, i, 0, 1, c_); }
inline BlockRef MatX::col(long j) { return BlockRef(*this, 0, j, r_, 1); }
inline BlockRef MatX::topLeftCorner(long r, long c) { return BlockRef(*this, 0, 0, r, c); }
template <int R, int C> inline BlockRef MatX::topLeftCorner() { return BlockRef(*this, 0, 0, R, C); }
template <int N> inline BlockRef MatX::topRows() { return BlockRef(*this, 0, 0, N, c_); }
template <int N> inline BlockRef MatX::bottomRows() { return BlockRef(*this, r_ - N, 0, N, c_); }

// a comma initialiser writing into a view goes through a temporary of the view's shape
struct BlockComma {
  BlockRef ref; MatX tmp; CommaInit ci;
  BlockComma(const BlockRef& r) : ref(r), tmp(r.rows(), r.cols()), ci(tmp) {}
};

// ---- fixed-size front: Matrix<double, R, C> is a MatX that starts with the stated shape ----
template <typename S, int R, int C, int Opt = 0, int MR = R, int MC = C>
class Matrix : public MatX {
 public:
  using Scalar = S;
  Matrix() : MatX(R == Dynamic ? 0 : R, C == Dynamic ? 0 : C) {}
  Matrix(const MatX& o) : MatX(o) {}
  Matrix(MatX&& o) : MatX(std::move(o)) {}
  Matrix(const BlockRef& b) : MatX(b) {}
  // VectorXd v(n) / MatrixXd m(r, c) / Vector2d(x, y) share these signatures: decide by the static shape
  template <typename A, typename = typename std::enable_if<std::is_integral<A>::value>::type>
  explicit Matrix(A n) : MatX(R == Dynamic ? (long)n : R, C == Dynamic ? (R == Dynamic ? 1 : (long)n) : C) {
    static_assert(R == Dynamic || C == Dynamic, "size constructor on a fixed-size matrix");
  }
  template <typename A, typename B,
            typename = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>::type>
  Matrix(A a, B b) : MatX() { init2((double)a, (double)b); }
  // from a rotation object (AngleAxisd, Quaterniond)
  template <typename T, typename = decltype(std::declval<const T&>().toRotationMatrix())>
  Matrix(const T& rot) : MatX(rot.toRotationMatrix()) {}
  Matrix(double a, double b, double c) : MatX(3, 1) { d_[0] = a; d_[1] = b; d_[2] = c; }
  Matrix(double a, double b, double c, double d) : MatX(4, 1) { d_[0] = a; d_[1] = b; d_[2] = c; d_[3] = d; }
  Matrix& operator=(const MatX& o) { MatX::operator=(o); return *this; }
  Matrix& operator=(const BlockRef& b) { MatX::operator=(b); return *this; }

  static Matrix Zero() { return Matrix(); }
  static Matrix Zero(long n) { Matrix m; m.MatX::resize(R == Dynamic ? n : R, C == Dynamic ? (R == Dynamic ? 1 : n) : C); return m; }
  static Matrix Zero(long r, long c) { Matrix m; m.MatX::resize(r, c); return m; }
  static Matrix Ones() { Matrix m; m.setOnes(); return m; }
  static Matrix Ones(long n) { Matrix m = Zero(n); m.setOnes(); return m; }
  static Matrix Ones(long r, long c) { Matrix m = Zero(r, c); m.setOnes(); return m; }
  static Matrix Constant(double v) { Matrix m; m.setConstant(v); return m; }
  static Matrix Constant(long r, long c, double v) { Matrix m = Zero(r, c); m.setConstant(v); return m; }
  static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
  static Matrix Identity(long r, long c) { Matrix m = Zero(r, c); m.setIdentity(); return m; }
  static Matrix Random() { Matrix m; m.setRandom(); return m; }
  static Matrix Random(long n) { Matrix m = Zero(n); m.setRandom(); return m; }
  static Matrix Random(long r, long c) { Matrix m = Zero(r, c); m.setRandom(); return m; }
  static Matrix UnitX() { Matrix m; m(0) = 1.0; return m; }
  static Matrix UnitY() { Matrix m; m(1) = 1.0; return m; }
  static Matrix UnitZ() { Matrix m; m(2) = 1.0; return m; }

 private:
  void init2(double a, double b) {
    if (R == Dynamic && C == Dynamic) { MatX::resize((long)a, (long)b); }       // MatrixXd(rows, cols)
    else if (R == Dynamic || C == Dynamic) { MatX::resize((long)a, (long)b); }   // Matrix<double, Dynamic, 3>(rows, 3)
    else { MatX::resize(R, C); d_[0] = a; d_[1] = b; }                           // Vector2d(x, y)
  }
};

using MatrixXd = Matrix<double, Dynamic, Dynamic>;
using VectorXd = Matrix<double, Dynamic, 1>;
using RowVectorXd = Matrix<double, 1, Dynamic>;
using Matrix2d = Matrix<double, 2, 2>;
using Matrix3d = Matrix<double, 3, 3>;
using Matrix4d = Matrix<double, 4, 4>;
using Vector2d = Matrix<double, 2, 1>;
using Vector3d = Matrix<double, 3, 1>;
using Vector4d = Matrix<double, 4, 1>;

inline CommaInit BlockRef::operator<<(double) {
  throw std::logic_error("eigen_shim: comma initialiser on a block view is not supported");
}

// ---- decompositions -------------------------------------------------------------------------------------------
// Cholesky A = L L^T (lower); solve by two triangular substitutions.
class LLTShim {
 public:
  LLTShim() = default;
  explicit LLTShim(const MatX& a) { compute(a); }
  void compute(const MatX& a) {
    const long n = a.rows();
    L_ = MatX(n, n); ok_ = true;
    for (long j = 0; j < n; ++j) {
      double d = a(j, j);
      for (long k = 0; k < j; ++k) d -= L_(j, k) * L_(j, k);
      if (!(d > 0.0)) { ok_ = false; d = std::numeric_limits<double>::quiet_NaN(); }
      const double l = std::sqrt(d);
      L_(j, j) = l;
      for (long i = j + 1; i < n; ++i) {
        double s = a(i, j);
        for (long k = 0; k < j; ++k) s -= L_(i, k) * L_(j, k);
        L_(i, j) = s / l;
      }
    }
  }
  MatX matrixL() const { return L_; }
  MatX matrixU() const { return L_.transpose(); }
  int info() const { return ok_ ? 0 : 1; }
  MatX solve(const MatX& b) const {
    const long n = L_.rows();
    MatX x(b);
    for (long c = 0; c < x.cols(); ++c) {
      for (long i = 0; i < n; ++i) { double s = x(i, c); for (long k = 0; k < i; ++k) s -= L_(i, k) * x(k, c); x(i, c) = s / L_(i, i); }
      for (long i = n - 1; i >= 0; --i) { double s = x(i, c); for (long k = i + 1; k < n; ++k) s -= L_(k, i) * x(k, c); x(i, c) = s / L_(i, i); }
    }
    return x;
  }
 private:
  MatX L_; bool ok_ = false;
};

// A = L D L^T without pivoting (the reference only factors S = H P H^T + R, symmetric positive definite).
class LDLTShim {
 public:
  explicit LDLTShim(const MatX& a) {
    const long n = a.rows();
    L_ = MatX(n, n); D_ = MatX(n, 1);
    for (long j = 0; j < n; ++j) {
      double d = a(j, j);
      for (long k = 0; k < j; ++k) d -= L_(j, k) * L_(j, k) * D_(k);
      D_(j) = d; L_(j, j) = 1.0;
      for (long i = j + 1; i < n; ++i) {
        double s = a(i, j);
        for (long k = 0; k < j; ++k) s -= L_(i, k) * L_(j, k) * D_(k);
        L_(i, j) = s / d;
      }
    }
  }
  MatX solve(const MatX& b) const {
    const long n = L_.rows();
    MatX x(b);
    for (long c = 0; c < x.cols(); ++c) {
      for (long i = 0; i < n; ++i) { double s = x(i, c); for (long k = 0; k < i; ++k) s -= L_(i, k) * x(k, c); x(i, c) = s; }
      for (long i = 0; i < n; ++i) x(i, c) /= D_(i);
      for (long i = n - 1; i >= 0; --i) { double s = x(i, c); for (long k = i + 1; k < n; ++k) s -= L_(k, i) * x(k, c); x(i, c) = s; }
    }
    return x;
  }
  bool isPositive() const { for (long i = 0; i < D_.rows(); ++i) if (!(D_(i) > 0)) return false; return true; }
 private:
  MatX L_, D_;
};

// Householder QR, A (m x n) = Q R with Q m x m explicit.
class HouseholderQRShim {
 public:
  HouseholderQRShim() = default;
  explicit HouseholderQRShim(const MatX& a) { compute(a); }
  void compute(const MatX& a) {
    const long m = a.rows(), n = a.cols();
    R_ = a; Q_ = MatX(m, m); Q_.setIdentity();
    for (long j = 0; j < std::min(m - 1, n); ++j) {
      double s = 0;
      for (long i = j; i < m; ++i) s += R_(i, j) * R_(i, j);
      const double nx = std::sqrt(s);
      if (nx == 0.0) continue;
      const double alpha = R_(j, j) > 0 ? -nx : nx;
      std::vector<double> v((size_t)m, 0.0);
      v[(size_t)j] = R_(j, j) - alpha;
      for (long i = j + 1; i < m; ++i) v[(size_t)i] = R_(i, j);
      double vn = 0; for (long i = j; i < m; ++i) vn += v[(size_t)i] * v[(size_t)i];
      if (vn == 0.0) continue;
      for (long c = 0; c < n; ++c) {   // R <- (I - 2 v v^T / v^T v) R
        double d = 0; for (long i = j; i < m; ++i) d += v[(size_t)i] * R_(i, c);
        d *= 2.0 / vn;
        for (long i = j; i < m; ++i) R_(i, c) -= d * v[(size_t)i];
      }
      for (long r = 0; r < m; ++r) {   // Q <- Q (I - 2 v v^T / v^T v)
        double d = 0; for (long i = j; i < m; ++i) d += Q_(r, i) * v[(size_t)i];
        d *= 2.0 / vn;
        for (long i = j; i < m; ++i) Q_(r, i) -= d * v[(size_t)i];
      }
    }
  }
  MatX householderQ() const { return Q_; }
  MatX matrixQR() const { return R_; }
  const MatX& Q() const { return Q_; }
  const MatX& R() const { return R_; }
 private:
  MatX Q_, R_;
};
template <typename M> class HouseholderQR : public HouseholderQRShim {
 public:
  HouseholderQR() = default;
  explicit HouseholderQR(const MatX& a) : HouseholderQRShim(a) {}
};
template <typename M> using LLT = LLTShim;

inline LLTShim MatX::llt() const { return LLTShim(*this); }
inline LDLTShim MatX::ldlt() const { return LDLTShim(*this); }
inline HouseholderQRShim MatX::householderQr() const { return HouseholderQRShim(*this); }

// LU with partial pivoting: determinant and inverse.
inline bool lu_factor(MatX& a, std::vector<long>& piv, int& sign) {
  const long n = a.rows();
  piv.resize((size_t)n); sign = 1;
  for (long k = 0; k < n; ++k) {
    long p = k; double best = std::fabs(a(k, k));
    for (long i = k + 1; i < n; ++i) if (std::fabs(a(i, k)) > best) { best = std::fabs(a(i, k)); p = i; }
    piv[(size_t)k] = p;
    if (p != k) { for (long j = 0; j < n; ++j) std::swap(a(k, j), a(p, j)); sign = -sign; }
    if (a(k, k) == 0.0) return false;
    for (long i = k + 1; i < n; ++i) {
      a(i, k) /= a(k, k);
      for (long j = k + 1; j < n; ++j) a(i, j) -= a(i, k) * a(k, j);
    }
  }
  return true;
}
inline double MatX::determinant() const {
  MatX a(*this); std::vector<long> piv; int sign;
  if (!lu_factor(a, piv, sign)) return 0.0;
  double d = sign; for (long i = 0; i < a.rows(); ++i) d *= a(i, i);
  return d;
}
inline MatX MatX::inverse() const {
  const long n = r_;
  MatX a(*this); std::vector<long> piv; int sign;
  MatX x(n, n); x.setIdentity();
  if (!lu_factor(a, piv, sign)) { x.setConstant(std::numeric_limits<double>::quiet_NaN()); return x; }
  for (long k = 0; k < n; ++k) if (piv[(size_t)k] != k) for (long j = 0; j < n; ++j) std::swap(x(k, j), x(piv[(size_t)k], j));
  for (long c = 0; c < n; ++c) {
    for (long i = 0; i < n; ++i) { double s = x(i, c); for (long k = 0; k < i; ++k) s -= a(i, k) * x(k, c); x(i, c) = s; }
    for (long i = n - 1; i >= 0; --i) { double s = x(i, c); for (long k = i + 1; k < n; ++k) s -= a(i, k) * x(k, c); x(i, c) = s / a(i, i); }
  }
  return x;
}

// Givens rotation with Eigen's conventions: makeGivens(p, q) gives G = [c s; -s c] with G^T [p; q] = [r; 0];
// applyOnTheLeft(i, j, G.adjoint()) therefore maps rows (p, q) to (r, 0): x <- c x - s y, y <- s x + c y.
template <typename S> class JacobiRotation {
 public:
  JacobiRotation() : c_(1), s_(0) {}
  JacobiRotation(S c, S s) : c_(c), s_(s) {}
  S c() const { return c_; }
  S s() const { return s_; }
  void makeGivens(S p, S q, S* r = nullptr) {
    if (q == S(0)) { c_ = p < S(0) ? S(-1) : S(1); s_ = S(0); if (r) *r = std::fabs(p); }
    else if (p == S(0)) { c_ = S(0); s_ = q < S(0) ? S(1) : S(-1); if (r) *r = std::fabs(q); }
    else if (std::fabs(p) > std::fabs(q)) {
      S t = q / p, u = std::sqrt(S(1) + t * t); if (p < S(0)) u = -u;
      c_ = S(1) / u; s_ = -t * c_; if (r) *r = p * u;
    } else {
      S t = p / q, u = std::sqrt(S(1) + t * t); if (q < S(0)) u = -u;
      s_ = -S(1) / u; c_ = -t * s_; if (r) *r = q * u;
    }
  }
  JacobiRotation adjoint() const { return JacobiRotation(c_, -s_); }
  JacobiRotation transpose() const { return JacobiRotation(c_, -s_); }
 private:
  S c_, s_;
};
template <typename S> inline void BlockRef::applyOnTheLeft(long p, long q, const JacobiRotation<S>& g) const {
  // Eigen: applyOnTheLeft(p, q, j) hands j itself to apply_rotation_in_the_plane(x = row p, y = row q, j), which does
  // x <- c x + s y, y <- -s x + c y with j's own (c, s); the reference passes tmpG.adjoint() = (c, -s).
  const double c = g.c(), s = g.s();
  for (long k = 0; k < c_; ++k) {
    const double xi = (*this)(p, k), yi = (*this)(q, k);
    (*this)(p, k) = c * xi + s * yi;
    (*this)(q, k) = -s * xi + c * yi;
  }
}
template <typename S> inline void MatX::applyOnTheLeft(long p, long q, const JacobiRotation<S>& g) {
  BlockRef(*this, 0, 0, r_, c_).applyOnTheLeft(p, q, g);
}

// "JacobiSVD" with the one product the reference reads: matrixU() of a tall m x n matrix with ComputeFullU.  The
// columns n.. of U are an orthonormal basis of the left null space (what RemoveLostUpdate.cpp:268-272 uses); Eigen
// obtains them from the Householder QR preconditioner of the Jacobi sweeps, here they are the trailing columns of a
// Householder Q. The first n columns span range(A) (not the singular vectors: nothing in the reference reads them).
template <typename M> class JacobiSVD {
 public:
  JacobiSVD() = default;
  JacobiSVD(const MatX& a, unsigned = 0) { compute(a); }
  JacobiSVD& compute(const MatX& a, unsigned = 0) { qr_.compute(a); return *this; }
  const MatX& matrixU() const { return qr_.Q(); }
 private:
  HouseholderQRShim qr_;
};

// ---- SparseCore / SPQRSupport stand-ins: a "sparse" matrix is a dense copy, SPQR is a dense Householder QR with the
// natural column order (SPQR_ORDERING_NATURAL, RemoveLostUpdate.cpp:141-151) ----
struct SparseShim { MatX dense; };
template <typename S> class SparseMatrix {
 public:
  SparseMatrix() = default;
  SparseMatrix(const SparseShim& s) : dense_(s.dense) {}
  SparseMatrix& operator=(const SparseShim& s) { dense_ = s.dense; return *this; }
  long rows() const { return dense_.rows(); }
  long cols() const { return dense_.cols(); }
  const MatX& dense() const { return dense_; }
 private:
  MatX dense_;
};
inline SparseShim MatX::sparseView() const { return SparseShim{*this}; }

struct EvalToProduct {
  MatX value;
  void evalTo(MatX& dst) const { dst = value; }
  operator MatX() const { return value; }
};
struct SPQRQt {
  const MatX* Q;
  EvalToProduct operator*(const MatX& rhs) const { return EvalToProduct{Q->transpose() * rhs}; }
};
struct SPQRQ {
  const MatX* Q;
  SPQRQt transpose() const { return SPQRQt{Q}; }
  SPQRQt adjoint() const { return SPQRQt{Q}; }
  EvalToProduct operator*(const MatX& rhs) const { return EvalToProduct{(*Q) * rhs}; }
};
template <typename SM> class SPQR {
 public:
  SPQR() = default;
  void setSPQROrdering(int) {}
  void setPivotThreshold(double) {}
  void compute(const SM& a) { qr_.compute(a.dense()); }
  SPQRQ matrixQ() const { return SPQRQ{&qr_.Q()}; }
  MatX matrixR() const { return qr_.R(); }
  long rank() const { return std::min(qr_.R().rows(), qr_.R().cols()); }
  int info() const { return 0; }
 private:
  HouseholderQRShim qr_;
};

// ---- Geometry: AngleAxis, Quaternion (w, x, y, z), Isometry3d ----
template <typename S> class Quaternion;
template <typename S> class AngleAxis {
 public:
  AngleAxis() : angle_(0), axis_(1.0, 0.0, 0.0) {}
  AngleAxis(S angle, const MatX& axis) : angle_(angle), axis_(axis) {}
  explicit AngleAxis(const Quaternion<S>& q);
  S angle() const { return angle_; }
  const Vector3d& axis() const { return axis_; }
  Matrix3d toRotationMatrix() const {
    const double c = std::cos(angle_), s = std::sin(angle_), t = 1 - c;
    const double x = axis_(0), y = axis_(1), z = axis_(2);
    Matrix3d R;
    R(0, 0) = t * x * x + c;     R(0, 1) = t * x * y - s * z; R(0, 2) = t * x * z + s * y;
    R(1, 0) = t * x * y + s * z; R(1, 1) = t * y * y + c;     R(1, 2) = t * y * z - s * x;
    R(2, 0) = t * x * z - s * y; R(2, 1) = t * y * z + s * x; R(2, 2) = t * z * z + c;
    return R;
  }
  Matrix3d matrix() const { return toRotationMatrix(); }
 private:
  S angle_; Vector3d axis_;
};
using AngleAxisd = AngleAxis<double>;

template <typename S> class Quaternion {
 public:
  Quaternion() : w_(1), x_(0), y_(0), z_(0) {}
  Quaternion(S w, S x, S y, S z) : w_(w), x_(x), y_(y), z_(z) {}
  Quaternion(const MatX& m) { if (m.rows() == 3 && m.cols() == 3) fromMatrix(m); else { x_ = m(0); y_ = m(1); z_ = m(2); w_ = m(3); } }
  Quaternion(const AngleAxis<S>& aa) {
    const double h = 0.5 * aa.angle(), s = std::sin(h);
    w_ = std::cos(h); x_ = s * aa.axis()(0); y_ = s * aa.axis()(1); z_ = s * aa.axis()(2);
  }
  Quaternion& operator=(const MatX& m) { fromMatrix(m); return *this; }
  static Quaternion Identity() { return Quaternion(); }
  static Quaternion UnitRandom() {
    Quaternion q(2.0 * std::rand() / RAND_MAX - 1, 2.0 * std::rand() / RAND_MAX - 1, 2.0 * std::rand() / RAND_MAX - 1, 2.0 * std::rand() / RAND_MAX - 1);
    q.normalize(); return q;
  }
  static Quaternion FromTwoVectors(const MatX& a, const MatX& b) { Quaternion q; q.setFromTwoVectors(a, b); return q; }
  Quaternion& setFromTwoVectors(const MatX& a, const MatX& b) {
    Vector3d v0 = a.normalized(), v1 = b.normalized();
    const double c = v0.dot(v1);
    if (c < -1.0 + 1e-12) {   // opposite vectors: any axis orthogonal to v0
      Vector3d ax = std::fabs(v0(0)) < 0.9 ? Vector3d(1, 0, 0) : Vector3d(0, 1, 0);
      Vector3d ort = v0.cross(ax).normalized();
      w_ = 0; x_ = ort(0); y_ = ort(1); z_ = ort(2);
      return *this;
    }
    Vector3d ax = v0.cross(v1);
    const double s = std::sqrt((1 + c) * 2), inv = 1 / s;
    x_ = ax(0) * inv; y_ = ax(1) * inv; z_ = ax(2) * inv; w_ = s * 0.5;
    return *this;
  }
  S& w() { return w_; } S& x() { return x_; } S& y() { return y_; } S& z() { return z_; }
  const S& w() const { return w_; } const S& x() const { return x_; } const S& y() const { return y_; } const S& z() const { return z_; }
  Vector3d vec() const { return Vector3d(x_, y_, z_); }
  Vector4d coeffs() const { return Vector4d(x_, y_, z_, w_); }
  Quaternion& setIdentity() { w_ = 1; x_ = y_ = z_ = 0; return *this; }
  S norm() const { return std::sqrt(w_ * w_ + x_ * x_ + y_ * y_ + z_ * z_); }
  S squaredNorm() const { return w_ * w_ + x_ * x_ + y_ * y_ + z_ * z_; }
  void normalize() { const S n = norm(); w_ /= n; x_ /= n; y_ /= n; z_ /= n; }
  Quaternion normalized() const { Quaternion q(*this); q.normalize(); return q; }
  Quaternion conjugate() const { return Quaternion(w_, -x_, -y_, -z_); }
  Quaternion inverse() const { const S n2 = squaredNorm(); return Quaternion(w_ / n2, -x_ / n2, -y_ / n2, -z_ / n2); }
  Quaternion operator*(const Quaternion& b) const {
    return Quaternion(w_ * b.w_ - x_ * b.x_ - y_ * b.y_ - z_ * b.z_, w_ * b.x_ + x_ * b.w_ + y_ * b.z_ - z_ * b.y_,
                      w_ * b.y_ + y_ * b.w_ + z_ * b.x_ - x_ * b.z_, w_ * b.z_ + z_ * b.w_ + x_ * b.y_ - y_ * b.x_);
  }
  Vector3d operator*(const MatX& v) const { return toRotationMatrix() * v; }
  Vector3d _transformVector(const MatX& v) const { return toRotationMatrix() * v; }
  S dot(const Quaternion& o) const { return w_ * o.w_ + x_ * o.x_ + y_ * o.y_ + z_ * o.z_; }
  S angularDistance(const Quaternion& o) const {
    Quaternion d = (*this) * o.conjugate();
    return 2 * std::atan2(d.vec().norm(), std::fabs(d.w()));
  }
  Matrix3d toRotationMatrix() const {
    const S tx = 2 * x_, ty = 2 * y_, tz = 2 * z_;
    const S twx = tx * w_, twy = ty * w_, twz = tz * w_, txx = tx * x_, txy = ty * x_, txz = tz * x_;
    const S tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
    Matrix3d R;
    R(0, 0) = 1 - (tyy + tzz); R(0, 1) = txy - twz;       R(0, 2) = txz + twy;
    R(1, 0) = txy + twz;       R(1, 1) = 1 - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy;       R(2, 1) = tyz + twx;       R(2, 2) = 1 - (txx + tyy);
    return R;
  }
  Matrix3d matrix() const { return toRotationMatrix(); }
  bool isApprox(const Quaternion& o, S prec = 1e-12) const { return coeffs().isApprox(o.coeffs(), prec); }

 private:
  void fromMatrix(const MatX& m) {   // Shepperd's branch on the trace (w >= 0 on the first branch)
    const S t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0) {
      S s = std::sqrt(t + 1.0); w_ = 0.5 * s; s = 0.5 / s;
      x_ = (m(2, 1) - m(1, 2)) * s; y_ = (m(0, 2) - m(2, 0)) * s; z_ = (m(1, 0) - m(0, 1)) * s;
    } else {
      int i = 0; if (m(1, 1) > m(0, 0)) i = 1; if (m(2, 2) > m(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      S s = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
      S q[3]; q[i] = 0.5 * s; s = 0.5 / s;
      w_ = (m(k, j) - m(j, k)) * s; q[j] = (m(j, i) + m(i, j)) * s; q[k] = (m(k, i) + m(i, k)) * s;
      x_ = q[0]; y_ = q[1]; z_ = q[2];
    }
  }
  S w_, x_, y_, z_;
};
using Quaterniond = Quaternion<double>;
template <typename S> inline AngleAxis<S>::AngleAxis(const Quaternion<S>& q) {
  double n = q.vec().norm();
  if (q.w() < 0) n = -n;
  if (n != 0) { angle_ = 2 * std::atan2(n, std::fabs(q.w())); axis_ = q.vec() / n; }
  else { angle_ = 0; axis_ = Vector3d(1, 0, 0); }
}
inline Matrix3d operator*(const AngleAxisd& a, const AngleAxisd& b) { return Matrix3d(a.toRotationMatrix() * b.toRotationMatrix()); }

enum TransformMode { Isometry = 1, Affine = 2 };
template <typename S, int Dim, int Mode> class Transform {
 public:
  Transform() { R_.setIdentity(); t_.setZero(); }
  static Transform Identity() { return Transform(); }
  Transform& setIdentity() { R_.setIdentity(); t_.setZero(); return *this; }
  Matrix3d& linear() { return R_; }
  const Matrix3d& linear() const { return R_; }
  Matrix3d rotation() const { return R_; }
  Vector3d& translation() { return t_; }
  const Vector3d& translation() const { return t_; }
  Matrix4d matrix() const {
    Matrix4d M; M.setIdentity();
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) M(i, j) = R_(i, j); M(i, 3) = t_(i); }
    return M;
  }
  Transform inverse() const { Transform T; T.R_ = R_.transpose(); T.t_ = -(T.R_ * t_); return T; }
  Transform operator*(const Transform& o) const { Transform T; T.R_ = R_ * o.R_; T.t_ = R_ * o.t_ + t_; return T; }
  Vector3d operator*(const MatX& v) const { return Vector3d(R_ * v + t_); }
  bool isApprox(const Transform& o, double prec = 1e-12) const { return matrix().isApprox(o.matrix(), prec); }
 private:
  Matrix3d R_; Vector3d t_;
};
using Isometry3d = Transform<double, 3, Isometry>;

}  // namespace Eigen

#ifndef SPQR_ORDERING_NATURAL
#define SPQR_ORDERING_NATURAL 3
#endif
