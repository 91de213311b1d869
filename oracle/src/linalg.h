// ORACLE — TEST INFRASTRUCTURE ONLY.
// Minimal dense linear algebra for the CPU oracle (stand-in for the Eigen calls in the reference:
// matrix products, fullPivLu().kernel()/solve() at qm_wbc/src/HoQp.cpp:129 and in OCS2's
// luConstraintProjection, LLT in the Riccati recursion).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <stdexcept>
#include <type_traits>
#include <vector>

namespace orc {

struct Mat {
  int r = 0, c = 0;
  std::vector<double> a;  // row-major
  Mat() {}
  Mat(int rows, int cols, double fill = 0.0) : r(rows), c(cols), a((size_t)rows * cols, fill) {}
  double& operator()(int i, int j) { return a[(size_t)i * c + j]; }
  double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
  static Mat identity(int n) { Mat m(n, n); for (int i = 0; i < n; ++i) m(i, i) = 1.0; return m; }
  Mat T() const { Mat t(c, r); for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) t(j, i) = (*this)(i, j); return t; }
  Mat block(int i0, int j0, int nr, int nc) const {
    Mat b(nr, nc); for (int i = 0; i < nr; ++i) for (int j = 0; j < nc; ++j) b(i, j) = (*this)(i0 + i, j0 + j); return b; }
  void set_block(int i0, int j0, const Mat& b) {
    for (int i = 0; i < b.r; ++i) for (int j = 0; j < b.c; ++j) (*this)(i0 + i, j0 + j) = b(i, j); }
  void add_block(int i0, int j0, const Mat& b, double s = 1.0) {
    for (int i = 0; i < b.r; ++i) for (int j = 0; j < b.c; ++j) (*this)(i0 + i, j0 + j) += s * b(i, j); }
};
using Vec = std::vector<double>;

inline Mat operator*(const Mat& A, const Mat& B) {
  assert(A.c == B.r);
  Mat C(A.r, B.c);
  for (int i = 0; i < A.r; ++i)
    for (int k = 0; k < A.c; ++k) {
      double aik = A(i, k);
      if (aik == 0.0) continue;
      const double* b = &B.a[(size_t)k * B.c];
      double* cc = &C.a[(size_t)i * C.c];
      for (int j = 0; j < B.c; ++j) cc[j] += aik * b[j];
    }
  return C;
}
inline Mat operator+(const Mat& A, const Mat& B) { assert(A.r == B.r && A.c == B.c); Mat C = A; for (size_t i = 0; i < C.a.size(); ++i) C.a[i] += B.a[i]; return C; }
inline Mat operator-(const Mat& A, const Mat& B) { assert(A.r == B.r && A.c == B.c); Mat C = A; for (size_t i = 0; i < C.a.size(); ++i) C.a[i] -= B.a[i]; return C; }
inline Mat operator*(double s, const Mat& A) { Mat C = A; for (auto& x : C.a) x *= s; return C; }
inline Vec operator*(const Mat& A, const Vec& x) {
  assert(A.c == (int)x.size()); Vec y(A.r, 0.0);
  for (int i = 0; i < A.r; ++i) { double s = 0; for (int j = 0; j < A.c; ++j) s += A(i, j) * x[j]; y[i] = s; } return y; }
inline Vec tmul(const Mat& A, const Vec& x) {  // A^T x
  assert(A.r == (int)x.size()); Vec y(A.c, 0.0);
  for (int i = 0; i < A.r; ++i) for (int j = 0; j < A.c; ++j) y[j] += A(i, j) * x[i]; return y; }
inline Vec operator+(const Vec& a, const Vec& b) { assert(a.size() == b.size()); Vec c = a; for (size_t i = 0; i < c.size(); ++i) c[i] += b[i]; return c; }
inline Vec operator-(const Vec& a, const Vec& b) { assert(a.size() == b.size()); Vec c = a; for (size_t i = 0; i < c.size(); ++i) c[i] -= b[i]; return c; }
inline Vec operator*(double s, const Vec& a) { Vec c = a; for (auto& x : c) x *= s; return c; }
inline double dot(const Vec& a, const Vec& b) { double s = 0; for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i]; return s; }
inline double norm2(const Vec& a) { return dot(a, a); }
inline Mat vstack(const Mat& A, const Mat& B) {
  if (A.r == 0 && A.c == 0) return B; if (B.r == 0 && B.c == 0) return A;
  if (A.r == 0) return B; if (B.r == 0) return A;
  assert(A.c == B.c); Mat C(A.r + B.r, A.c); C.set_block(0, 0, A); C.set_block(A.r, 0, B); return C; }
inline Vec vcat(const Vec& a, const Vec& b) { Vec c = a; c.insert(c.end(), b.begin(), b.end()); return c; }
inline Vec seg(const Vec& a, int i0, int n) { return Vec(a.begin() + i0, a.begin() + i0 + n); }

// Full-pivot LU (Eigen::FullPivLU semantics: P A Q = L U, rank by threshold eps*max(r,c)*|maxpivot|).
struct FullPivLU {
  Mat lu; std::vector<int> rowperm, colperm; int rank = 0; int rows, cols;
  explicit FullPivLU(const Mat& A) : lu(A), rows(A.r), cols(A.c) {
    rowperm.resize(rows); colperm.resize(cols);
    for (int i = 0; i < rows; ++i) rowperm[i] = i;
    for (int j = 0; j < cols; ++j) colperm[j] = j;
    const int n = std::min(rows, cols);
    double maxpivot = 0.0; std::vector<double> piv(n, 0.0);
    int k = 0;
    for (; k < n; ++k) {
      int pi = k, pj = k; double best = 0.0;
      for (int i = k; i < rows; ++i) for (int j = k; j < cols; ++j) { double v = std::fabs(lu(i, j)); if (v > best) { best = v; pi = i; pj = j; } }
      if (best == 0.0) break;
      if (pi != k) { for (int j = 0; j < cols; ++j) std::swap(lu(pi, j), lu(k, j)); std::swap(rowperm[pi], rowperm[k]); }
      if (pj != k) { for (int i = 0; i < rows; ++i) std::swap(lu(i, pj), lu(i, k)); std::swap(colperm[pj], colperm[k]); }
      piv[k] = std::fabs(lu(k, k)); maxpivot = std::max(maxpivot, piv[k]);
      for (int i = k + 1; i < rows; ++i) {
        double f = lu(i, k) / lu(k, k); lu(i, k) = f;
        if (f != 0.0) for (int j = k + 1; j < cols; ++j) lu(i, j) -= f * lu(k, j);
      }
    }
    const double thr = 2.220446049250313e-16 * std::max(rows, cols) * maxpivot;
    rank = 0; for (int i = 0; i < k; ++i) if (piv[i] > thr) ++rank;
  }
  // Basis of the null space, cols x (cols-rank); Eigen returns one zero column when the kernel is {0}.
  Mat kernel() const {
    const int dimker = cols - rank;
    if (dimker == 0) return Mat(cols, 1, 0.0);
    // U = [U11 U12; 0 0] in pivoted coordinates; kernel (pivoted) = [-U11^{-1} U12; I]
    Mat K(cols, dimker);
    for (int j = 0; j < dimker; ++j) {
      Vec y(rank, 0.0);
      for (int i = rank - 1; i >= 0; --i) {
        double s = -lu(i, rank + j);
        for (int l = i + 1; l < rank; ++l) s -= lu(i, l) * y[l];
        y[i] = s / lu(i, i);
      }
      for (int i = 0; i < rank; ++i) K(colperm[i], j) = y[i];
      K(colperm[rank + j], j) = 1.0;
    }
    return K;
  }
  // A particular solution of A X = B (free variables set to zero), as Eigen::FullPivLU::solve.
  Mat solve(const Mat& B) const {
    assert(B.r == rows); Mat X(cols, B.c, 0.0);
    for (int col = 0; col < B.c; ++col) {
      Vec c(rows);
      for (int i = 0; i < rows; ++i) c[i] = B(rowperm[i], col);
      const int n = std::min(rows, cols);
      for (int i = 0; i < rows; ++i) { const int lim = std::min(i, n); for (int l = 0; l < lim; ++l) c[i] -= lu(i, l) * c[l]; }
      Vec y(rank, 0.0);
      for (int i = rank - 1; i >= 0; --i) { double s = c[i]; for (int l = i + 1; l < rank; ++l) s -= lu(i, l) * y[l]; y[i] = s / lu(i, i); }
      for (int i = 0; i < rank; ++i) X(colperm[i], col) = y[i];
    }
    return X;
  }
};

// Cholesky A = L L^T (lower). Returns false if not PD.
inline bool cholesky(const Mat& A, Mat& L) {
  const int n = A.r; L = Mat(n, n);
  for (int j = 0; j < n; ++j) {
    double s = A(j, j); for (int k = 0; k < j; ++k) s -= L(j, k) * L(j, k);
    if (!(s > 0.0)) return false;
    L(j, j) = std::sqrt(s);
    for (int i = j + 1; i < n; ++i) { double t = A(i, j); for (int k = 0; k < j; ++k) t -= L(i, k) * L(j, k); L(i, j) = t / L(j, j); }
  }
  return true;
}
inline Mat chol_solve(const Mat& L, const Mat& B) {  // solves (L L^T) X = B
  const int n = L.r; Mat X = B;
  for (int c = 0; c < B.c; ++c) {
    for (int i = 0; i < n; ++i) { double s = X(i, c); for (int k = 0; k < i; ++k) s -= L(i, k) * X(k, c); X(i, c) = s / L(i, i); }
    for (int i = n - 1; i >= 0; --i) { double s = X(i, c); for (int k = i + 1; k < n; ++k) s -= L(k, i) * X(k, c); X(i, c) = s / L(i, i); }
  }
  return X;
}
inline Mat col(const Vec& v) { Mat m((int)v.size(), 1); for (size_t i = 0; i < v.size(); ++i) m((int)i, 0) = v[i]; return m; }
inline Vec colvec(const Mat& m, int j = 0) { Vec v(m.r); for (int i = 0; i < m.r; ++i) v[i] = m(i, j); return v; }

// Householder QR of A (m x n, m >= n not required): returns Q (m x m) and R (m x n).
inline void householder_qr(const Mat& A, Mat& Q, Mat& R) {
  const int m = A.r, n = A.c; R = A; Q = Mat::identity(m);
  for (int k = 0; k < std::min(m - 1, n); ++k) {
    double nrm = 0; for (int i = k; i < m; ++i) nrm += R(i, k) * R(i, k); nrm = std::sqrt(nrm);
    if (nrm == 0.0) continue;
    double alpha = R(k, k) > 0 ? -nrm : nrm;
    Vec v(m, 0.0); for (int i = k; i < m; ++i) v[i] = R(i, k); v[k] -= alpha;
    double vn = 0; for (int i = k; i < m; ++i) vn += v[i] * v[i];
    if (vn == 0.0) continue;
    for (int j = 0; j < n; ++j) { double s = 0; for (int i = k; i < m; ++i) s += v[i] * R(i, j); s *= 2.0 / vn; for (int i = k; i < m; ++i) R(i, j) -= s * v[i]; }
    for (int j = 0; j < m; ++j) { double s = 0; for (int i = k; i < m; ++i) s += v[i] * Q(j, i); s *= 2.0 / vn; for (int i = k; i < m; ++i) Q(j, i) -= s * v[i]; }
  }
}

// Householder QR with column pivoting: A P = Q [R; 0].  Returns the numerical rank (pivots below tol * largest pivot are treated as zero).
// Q is m x m, R is m x n (upper trapezoidal in its first `rank` rows), perm[j] = original index of the column now at position j.
inline int householder_qrcp(const Mat& A, Mat& Q, Mat& R, std::vector<int>& perm, double tol = 1e-11) {
  const int m = A.r, n = A.c; R = A; Q = Mat::identity(m); perm.resize(n); for (int j = 0; j < n; ++j) perm[j] = j;
  int rank = 0; double first = 0.0;
  for (int k = 0; k < std::min(m, n); ++k) {
    int piv = k; double best = -1.0;
    for (int j = k; j < n; ++j) { double s = 0; for (int i = k; i < m; ++i) s += R(i, j) * R(i, j); if (s > best) { best = s; piv = j; } }
    const double nrm = std::sqrt(std::max(best, 0.0)); if (k == 0) first = nrm;
    if (!(nrm > tol * first) || nrm == 0.0) break;
    if (piv != k) { for (int i = 0; i < m; ++i) std::swap(R(i, k), R(i, piv)); std::swap(perm[k], perm[piv]); }
    const double alpha = R(k, k) > 0 ? -nrm : nrm;
    Vec v(m, 0.0); for (int i = k; i < m; ++i) v[i] = R(i, k); v[k] -= alpha;
    double vn = 0; for (int i = k; i < m; ++i) vn += v[i] * v[i];
    if (vn > 0.0) {
      for (int j = k; j < n; ++j) { double s = 0; for (int i = k; i < m; ++i) s += v[i] * R(i, j); s *= 2.0 / vn; for (int i = k; i < m; ++i) R(i, j) -= s * v[i]; }
      for (int j = 0; j < m; ++j) { double s = 0; for (int i = k; i < m; ++i) s += v[i] * Q(j, i); s *= 2.0 / vn; for (int i = k; i < m; ++i) Q(j, i) -= s * v[i]; }
    }
    ++rank;
  }
  return rank;
}
// Equality-constrained least squares  min |A x - b|  s.t.  E x = e  (E may have dependent rows; the system is assumed consistent) by the
// null-space method with orthogonal factorisations only (no normal equations): x = Q1 y1 + Q2 y2, E' P = Q [R; 0].
// In flat directions of the reduced problem the minimum-norm solution is taken.
inline Vec constrained_lstsq(const Mat& A, const Vec& b, const Mat& E, const Vec& e) {
  const int n = A.c; Mat Q = Mat::identity(n); int k = 0; Vec y1;
  if (E.r > 0) { Mat R; std::vector<int> perm; k = householder_qrcp(E.T(), Q, R, perm);
    y1.assign(k, 0.0);   // R11' y1 = (P' e)[0:k]  (lower triangular, forward substitution)
    for (int i = 0; i < k; ++i) { double s = e[perm[i]]; for (int l = 0; l < i; ++l) s -= R(l, i) * y1[l]; y1[i] = s / R(i, i); } }
  Vec x(n, 0.0); for (int i = 0; i < n; ++i) for (int l = 0; l < k; ++l) x[i] += Q(i, l) * y1[l];
  const int nf = n - k; if (nf == 0 || A.r == 0) return x;
  Mat Q2 = Q.block(0, k, n, nf); Mat AQ = A * Q2; Vec r = b - A * x;
  Mat Qa, Ra; std::vector<int> pa; const int ka = householder_qrcp(AQ, Qa, Ra, pa);
  Vec c = tmul(Qa, r); Vec z(nf, 0.0);
  if (ka == nf) { for (int i = ka - 1; i >= 0; --i) { double s = c[i]; for (int l = i + 1; l < ka; ++l) s -= Ra(i, l) * z[pa[l]]; z[pa[i]] = s / Ra(i, i); } }
  else if (ka > 0) {   // rank deficient: minimum-norm solution of [R11 R12] w = c1 through the QR of the transposed trapezoid (complete orthogonal decomposition)
    Mat Tt(nf, ka); for (int i = 0; i < ka; ++i) for (int j = 0; j < nf; ++j) Tt(j, i) = Ra(i, j);
    Mat Qb, Rb; householder_qr(Tt, Qb, Rb); Vec u(ka, 0.0);
    for (int i = 0; i < ka; ++i) { double s = c[i]; for (int l = 0; l < i; ++l) s -= Rb(l, i) * u[l]; u[i] = s / Rb(i, i); }
    for (int j = 0; j < nf; ++j) { double w = 0; for (int l = 0; l < ka; ++l) w += Qb(j, l) * u[l]; z[pa[j]] = w; }
  }
  for (int i = 0; i < n; ++i) for (int l = 0; l < nf; ++l) x[i] += Q2(i, l) * z[l];
  return x;
}

// ---- small fixed-size templated types for kinematics ----
template <class T> struct V3 { T x, y, z; V3() : x(T(0.0)), y(T(0.0)), z(T(0.0)) {} V3(T a, T b, T c) : x(a), y(b), z(c) {}
  T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); } const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); } };
template <class T> inline V3<T> operator+(const V3<T>& a, const V3<T>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class T> inline V3<T> operator-(const V3<T>& a, const V3<T>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class T> inline V3<T> operator*(const T& s, const V3<T>& a) { return {s * a.x, s * a.y, s * a.z}; }
template <class T> inline V3<T> cross(const V3<T>& a, const V3<T>& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
template <class T> inline T dot(const V3<T>& a, const V3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> struct M3 { T m[3][3]; M3() { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[i][j] = T(0.0); }
  static M3 identity() { M3 r; for (int i = 0; i < 3; ++i) r.m[i][i] = T(1.0); return r; }
  T& operator()(int i, int j) { return m[i][j]; } const T& operator()(int i, int j) const { return m[i][j]; } };
template <class T> inline M3<T> operator*(const M3<T>& A, const M3<T>& B) { M3<T> C; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { T s = T(0.0); for (int k = 0; k < 3; ++k) s = s + A.m[i][k] * B.m[k][j]; C.m[i][j] = s; } return C; }
template <class T> inline V3<T> operator*(const M3<T>& A, const V3<T>& v) { V3<T> r; for (int i = 0; i < 3; ++i) r[i] = A.m[i][0] * v.x + A.m[i][1] * v.y + A.m[i][2] * v.z; return r; }
template <class T> inline M3<T> transpose(const M3<T>& A) { M3<T> C; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C.m[i][j] = A.m[j][i]; return C; }
template <class T> inline M3<T> operator+(const M3<T>& A, const M3<T>& B) { M3<T> C; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C.m[i][j] = A.m[i][j] + B.m[i][j]; return C; }
template <class T> inline M3<T> skew(const V3<T>& v) { M3<T> S; S.m[0][1] = -v.z; S.m[0][2] = v.y; S.m[1][0] = v.z; S.m[1][2] = -v.x; S.m[2][0] = -v.y; S.m[2][1] = v.x; return S; }
template <class T, class U> inline M3<T> cast3(const M3<U>& A) { M3<T> C; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C.m[i][j] = T(A.m[i][j]); return C; }
template <class T, class U> inline V3<T> cast3(const V3<U>& a) { return V3<T>(T(a.x), T(a.y), T(a.z)); }
inline M3<double> inverse3(const M3<double>& A) {
  M3<double> B; const auto& m = A.m;
  double det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) + m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
  double id = 1.0 / det;
  B.m[0][0] = (m[1][1] * m[2][2] - m[1][2] * m[2][1]) * id; B.m[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * id; B.m[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * id;
  B.m[1][0] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]) * id; B.m[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * id; B.m[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id;
  B.m[2][0] = (m[1][0] * m[2][1] - m[1][1] * m[2][0]) * id; B.m[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * id; B.m[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id;
  return B;
}

}  // namespace orc
