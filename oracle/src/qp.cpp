// ORACLE — TEST INFRASTRUCTURE ONLY.  See qp.h.
#include "qp.h"

namespace orc {
namespace {

// Solve H x = b for symmetric PSD H by diagonally pivoted Cholesky, truncating pivots below
// tol*max_diag (basic solution: components outside the numerical range are zero).
Vec solve_psd(const Mat& H, const Vec& b, double tol = 1e-11) {
  const int n = H.r; Mat A = H; std::vector<int> perm(n); for (int i = 0; i < n; ++i) perm[i] = i;
  double maxdiag = 0; for (int i = 0; i < n; ++i) maxdiag = std::max(maxdiag, std::fabs(A(i, i)));
  Mat L(n, n); int rank = 0;
  for (int k = 0; k < n; ++k) {
    int piv = k; double best = A(k, k); for (int i = k + 1; i < n; ++i) if (A(i, i) > best) { best = A(i, i); piv = i; }
    if (!(best > tol * maxdiag) || best <= 0.0) break;
    if (piv != k) { for (int j = 0; j < n; ++j) std::swap(A(k, j), A(piv, j)); for (int i = 0; i < n; ++i) std::swap(A(i, k), A(i, piv));
      for (int j = 0; j < k; ++j) std::swap(L(k, j), L(piv, j)); std::swap(perm[k], perm[piv]); }
    const double d = std::sqrt(A(k, k)); L(k, k) = d;
    for (int i = k + 1; i < n; ++i) L(i, k) = A(i, k) / d;
    for (int i = k + 1; i < n; ++i) for (int j = k + 1; j < n; ++j) A(i, j) -= L(i, k) * L(j, k);
    ++rank;
  }
  Vec y(rank), x(n, 0.0);
  for (int i = 0; i < rank; ++i) { double s = b[perm[i]]; for (int k = 0; k < i; ++k) s -= L(i, k) * y[k]; y[i] = s / L(i, i); }
  for (int i = rank - 1; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < rank; ++k) s -= L(k, i) * y[k]; y[i] = s / L(i, i); }
  for (int i = 0; i < rank; ++i) x[perm[i]] = y[i];
  return x;
}

}  // namespace

QpResult solve_qp_active_set(const Mat& H, const Vec& c, const Mat& A, const Vec& ub, const Vec& z0, int max_iter) {
  const int n = H.r, m = A.r;
  QpResult res; Vec z = z0; std::vector<int> W; std::vector<char> inW(m, 0);
  for (int it = 0; it < max_iter; ++it) {
    res.iterations = it + 1;
    Vec g = H * z + c;
    const int nw = (int)W.size();
    Mat N, Y, R;
    if (nw > 0) {
      Mat AWt(n, nw); for (int k = 0; k < nw; ++k) for (int j = 0; j < n; ++j) AWt(j, k) = A(W[k], j);
      Mat Q; householder_qr(AWt, Q, R);
      Y = Q.block(0, 0, n, nw); N = (nw < n) ? Q.block(0, nw, n, n - nw) : Mat(n, 0);
    } else { N = Mat::identity(n); }
    Vec p(n, 0.0);
    if (N.c > 0) { Mat Hr = N.T() * (H * N); Vec gr = tmul(N, g); for (auto& x : gr) x = -x; Vec pn = solve_psd(Hr, gr); p = N * pn; }
    double pinf = 0, zinf = 0; for (int i = 0; i < n; ++i) { pinf = std::max(pinf, std::fabs(p[i])); zinf = std::max(zinf, std::fabs(z[i])); }
    if (pinf <= 1e-11 * (1.0 + zinf)) {
      if (nw == 0) break;
      // multipliers: A_W' lambda = -g  ->  R11 lambda = -Y'g
      Vec rhs = tmul(Y, g); Vec lam(nw, 0.0);
      for (int i = nw - 1; i >= 0; --i) { double s = -rhs[i]; for (int k = i + 1; k < nw; ++k) s -= R(i, k) * lam[k]; lam[i] = s / R(i, i); }
      int worst = -1; double lmin = -1e-9; double gscale = 0; for (double x : g) gscale = std::max(gscale, std::fabs(x));
      lmin *= (1.0 + gscale);
      for (int k = 0; k < nw; ++k) if (lam[k] < lmin) { lmin = lam[k]; worst = k; }
      if (worst < 0) break;
      inW[W[worst]] = 0; W.erase(W.begin() + worst);
    } else {
      double alpha = 1.0; int block = -1;
      for (int i = 0; i < m; ++i) { if (inW[i]) continue;
        double ap = 0, az = 0; for (int j = 0; j < n; ++j) { ap += A(i, j) * p[j]; az += A(i, j) * z[j]; }
        if (ap > 1e-13 * (1.0 + pinf)) { double ratio = (ub[i] - az) / ap; if (ratio < 0) ratio = 0; if (ratio < alpha) { alpha = ratio; block = i; } } }
      for (int j = 0; j < n; ++j) z[j] += alpha * p[j];
      if (block >= 0) { W.push_back(block); inW[block] = 1; }
    }
    if (it == max_iter - 1) res.status = 1;
  }
  res.z = z; res.active = W; return res;
}

}  // namespace orc
