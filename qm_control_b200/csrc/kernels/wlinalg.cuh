// Warp-level dense linear algebra on small shared-memory matrices (n <= 36): Householder QR with column
// pivoting (rank revealing), application of Q / Q^T, triangular solves, Cholesky.  These replace the Eigen
// calls on the reference WBC path (fullPivLu().kernel() at qm_wbc/src/HoQp.cpp:129, the dense products in
// HoQp::buildHMatrix/buildCVector HoQp.cpp:60-90) and qpOASES' internal factorizations.
#pragma once
#include "dev_common.cuh"

// column norms of the pivot search in w_qrcp: 0 = re-summed in every step, 1 = LAPACK-style downdating, 2 = re-summed for free inside the trailing update (default;
// measured on B200 at 8192 robots: profiles/r02_ab_wbc.jsonl)
#ifndef QMB_QRCP_NORMS
#define QMB_QRCP_NORMS 2
#endif

namespace qmb {


// QR with column pivoting of W (n x r, column-major, leading dimension ld, r <= 32).
// On exit: R in the upper triangle, Householder vectors below the diagonal (unit leading entry implied),
// tau[0..k), perm[c] = original index of the column now at position c.  Returns the numerical rank k
// (columns whose remaining norm is <= tol_rel * largest initial column norm are treated as zero).
__device__ __forceinline__ int w_qrcp(double* W, int n, int r, int ld, double* tau, int* perm, double tol_rel, int lane) {
  if (lane < r) perm[lane] = lane;
  // vn = squared norm of the not yet factored part of this lane's column.  Mode 1 DOWNDATES it after every reflector (vn -= R[j][c]^2) with vref = its value
  // at the last exact evaluation and the safeguard of LAPACK's dgeqp3: once cancellation has eaten half of the digits (vn <= sqrt(eps) * vref) the lane re-sums
  // its column, so every rank decision (columns near zero) is taken on an exact value.  Mode 2 forms the exact sum inside the trailing update.
  double vn = 0.0;
  if (lane < r) for (int i = 0; i < n; ++i) { const double x = W[i + lane * ld]; vn += x * x; }
  double vref = vn;
  const double thresh = tol_rel * tol_rel * warp_max(vn);
  const int kmax = n < r ? n : r; int rank = kmax;
  __syncwarp();
  for (int j = 0; j < kmax; ++j) {
#if QMB_QRCP_NORMS == 0
    if (lane >= j && lane < r) { vn = 0.0; for (int i = j; i < n; ++i) { const double x = W[i + lane * ld]; vn += x * x; } }
#endif
    double nr = vn; int pc = lane;
    warp_argmax_nonneg(nr, pc, lane >= j && lane < r);
    if (!(nr > thresh)) { rank = j; break; }
    if (pc != j) {
      for (int i = lane; i < n; i += 32) { const double t = W[i + j * ld]; W[i + j * ld] = W[i + pc * ld]; W[i + pc * ld] = t; }
      if (lane == 0) { const int t = perm[j]; perm[j] = perm[pc]; perm[pc] = t; }
      const double vj = __shfl_sync(FULL, vn, j), rj = __shfl_sync(FULL, vref, j);
      if (lane == pc) { vn = vj; vref = rj; }
    }
    __syncwarp();
#if QMB_QRCP_NORMS == 1
    { // exact norm of the pivot column (the reflector must annihilate it to rounding): rows over lanes
      double part = 0.0; for (int i = j + lane; i < n; i += 32) { const double x = W[i + j * ld]; part += x * x; }
      nr = warp_sum(part);
      if (!(nr > thresh)) { rank = j; break; }
    }
#endif
    // one reciprocal for both tau = (beta - x0) / beta and the scale 1 / (x0 - beta) of the reflector's tail
    const double x0 = W[j + j * ld]; const double beta = (x0 >= 0.0) ? -sqrt(nr) : sqrt(nr);
    const double d = x0 - beta, inv = 1.0 / (beta * d);
    const double tj = -d * d * inv; const double scal = beta * inv;
    __syncwarp();
    for (int i = j + 1 + lane; i < n; i += 32) W[i + j * ld] *= scal;
    if (lane == 0) { W[j + j * ld] = beta; tau[j] = tj; }
    __syncwarp();
    if (lane > j && lane < r) {
      double* col = W + lane * ld; const double* v = W + j * ld;
      double w = col[j]; for (int i = j + 1; i < n; ++i) w += v[i] * col[i];
      w *= tj; const double cj = col[j] - w; col[j] = cj;
#if QMB_QRCP_NORMS == 2
      double nn = 0.0; for (int i = j + 1; i < n; ++i) { const double c = fma(-w, v[i], col[i]); col[i] = c; nn = fma(c, c, nn); }
      vn = nn;   // the exact squared norm of what is left of this column, formed for free
#else
      for (int i = j + 1; i < n; ++i) col[i] -= w * v[i];
#endif
#if QMB_QRCP_NORMS == 1
      vn = fmax(vn - cj * cj, 0.0);
      if (vref > 0.0 && vn <= 1.4901161193847656e-8 * vref) {
        vn = 0.0; for (int i = j + 1; i < n; ++i) { const double x = col[i]; vn += x * x; }
        if (!(vn > thresh)) vn = 0.0;   // the remainder of a column only shrinks: below the rank threshold once = never a pivot, never re-summed again
        vref = vn;
      }
#endif
    }
    __syncwarp();
  }
  return rank;
}

// g <- Q^T g  (g has n entries in shared memory)
__device__ __forceinline__ void w_apply_qt(const double* W, int n, int k, int ld, const double* tau, double* g, int lane) {
  for (int j = 0; j < k; ++j) {
    const double* v = W + j * ld; double part = 0.0;
    for (int i = j + lane; i < n; i += 32) part += (i == j ? 1.0 : v[i]) * g[i];
    const double w = tau[j] * warp_sum(part);
    for (int i = j + lane; i < n; i += 32) g[i] -= w * (i == j ? 1.0 : v[i]);
    __syncwarp();
  }
}
// g <- Q g
__device__ __forceinline__ void w_apply_q(const double* W, int n, int k, int ld, const double* tau, double* g, int lane) {
  for (int j = k - 1; j >= 0; --j) {
    const double* v = W + j * ld; double part = 0.0;
    for (int i = j + lane; i < n; i += 32) part += (i == j ? 1.0 : v[i]) * g[i];
    const double w = tau[j] * warp_sum(part);
    for (int i = j + lane; i < n; i += 32) g[i] -= w * (i == j ? 1.0 : v[i]);
    __syncwarp();
  }
}
// C[:, c] <- Q^T C[:, c] for `ncols` columns stored column-major (leading dimension ldc): one lane per column, no reductions
__device__ __forceinline__ void w_apply_qt_cols(const double* W, int n, int k, int ld, const double* tau, double* C, int ncols, int ldc, int lane) {
  if (lane < ncols) {
    double* col = C + lane * ldc;
    for (int j = 0; j < k; ++j) {
      const double* v = W + j * ld; double w = col[j]; for (int i = j + 1; i < n; ++i) w = fma(v[i], col[i], w);
      w *= tau[j]; col[j] -= w; for (int i = j + 1; i < n; ++i) col[i] = fma(-w, v[i], col[i]);
    }
  }
  __syncwarp();
}
// Z[:, off:off+n] <- Z[:, off:off+n] * Q   (Z has `rows` rows, row-major, leading dimension ldz); lanes over rows
__device__ __forceinline__ void w_apply_q_right(const double* W, int n, int k, int ld, const double* tau, double* Z, int rows, int ldz, int off, int lane) {
  for (int j = 0; j < k; ++j) {
    const double* v = W + j * ld; const double tj = tau[j];
    for (int i = lane; i < rows; i += 32) {
      double* zr = Z + i * ldz + off; double d = zr[j]; for (int c = j + 1; c < n; ++c) d += zr[c] * v[c];
      d *= tj; zr[j] -= d; for (int c = j + 1; c < n; ++c) zr[c] -= d * v[c];
    }
    __syncwarp();
  }
}

// Z (n x (n - k), row-major, leading dimension ldz) <- the last n - k columns of Q = H_0 H_1 ... H_{k-1} (the orthonormal basis of the null space of
// the factored matrix' row space complement).  One lane per column: the reflectors are applied right to left to the unit vectors e_{k+c}.
__device__ __forceinline__ void w_form_q_tail(const double* W, int n, int k, int ld, const double* tau, double* Z, int ldz, int lane) {
  const int nc = n - k;
  if (lane < nc) {
    double* col = Z + lane;
    for (int i = 0; i < n; ++i) col[i * ldz] = (i == k + lane) ? 1.0 : 0.0;
    for (int j = k - 1; j >= 0; --j) {
      const double* v = W + j * ld; double w0 = col[j * ldz], w1 = 0.0; int i = j + 1;
      for (; i + 1 < n; i += 2) { w0 = fma(v[i], col[i * ldz], w0); w1 = fma(v[i + 1], col[(i + 1) * ldz], w1); }
      if (i < n) w0 = fma(v[i], col[i * ldz], w0);
      const double w = (w0 + w1) * tau[j]; col[j * ldz] -= w; for (int i2 = j + 1; i2 < n; ++i2) col[i2 * ldz] = fma(-w, v[i2], col[i2 * ldz]);
    }
  }
  __syncwarp();
}

// Packed lower triangle: entry (a, b <= a) of a symmetric matrix sits at a(a+1)/2 + b.  Half the storage of a square scratch, half the entries to form,
// and the triangular row offsets are distinct mod 16, so a lane = row walk down a column is bank-conflict free.
__device__ __forceinline__ int tri(int a) { return (a * (a + 1)) >> 1; }
// In-place Cholesky (lower, packed) of the n x n symmetric matrix A, n <= 32.  Returns false when a pivot <= 0.
__device__ __forceinline__ bool w_cholesky(double* A, int n, int lane) {
  bool ok = true;
  for (int j = 0; j < n; ++j) {
    const double d = A[tri(j) + j];
    if (!(d > 0.0)) { ok = false; break; }
    const double s = sqrt(d);
    __syncwarp();
    double* row = A + tri(lane);
    if (lane == j) row[j] = s;
    if (lane > j && lane < n) row[j] /= s;
    __syncwarp();
    if (lane > j && lane < n) { const double lij = row[j]; int oc = tri(j + 1) + j; for (int c = j + 1; c <= lane; ++c) { row[c] -= lij * A[oc]; oc += c + 1; } }
    __syncwarp();
  }
  return ok;
}
// Solve L L^T x = b in place (b in shared memory, n <= 32); L lower packed from w_cholesky.
__device__ __forceinline__ void w_chol_solve(const double* L, int n, double* b, int lane) {
  const double* row = L + tri(lane);
  for (int j = 0; j < n; ++j) {   // forward
    if (lane == j) b[j] /= row[j];
    __syncwarp();
    if (lane > j && lane < n) b[lane] -= row[j] * b[j];
    __syncwarp();
  }
  for (int j = n - 1; j >= 0; --j) {   // backward with L^T
    if (lane == j) b[j] /= row[j];
    __syncwarp();
    if (lane < j) b[lane] -= L[tri(j) + lane] * b[j];
    __syncwarp();
  }
}

}  // namespace qmb
