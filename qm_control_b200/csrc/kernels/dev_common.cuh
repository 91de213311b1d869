// Device-side common definitions for the batched MPC+WBC kernels (sm_100a, fp64).
// One warp owns one robot; lanes cooperate over bodies / matrix rows; all per-robot state lives in
// shared memory or registers, HBM is touched only for the batch I/O buffers.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

// Scalar helpers are host + device: tests/nodeeval_host.cpp compiles the one-thread-per-node evaluator (node_eval.cuh) with g++ and checks it against the
// oracle on the CPU.  Everything that needs warp intrinsics stays inside #ifdef __CUDACC__.
#ifdef __CUDACC__
#define QMB_HD __host__ __device__ __forceinline__
#else
#define QMB_HD inline
#endif

namespace qmb {

constexpr int NQ = 24;   // generalized coordinates (WbcBase.cpp:36, task.info:150-189)
constexpr int NJ = 18;   // actuated joints
constexpr int NB = 19;   // bodies (base + one per joint)
constexpr int NX = 30;   // centroidal state
constexpr int NU = 30;   // input: 12 contact forces (LF,RF,LH,RH) + 18 joint velocities
constexpr int NDEC = 36; // WBC decision vector [vdot(24); F(12)]
constexpr unsigned FULL = 0xffffffffu;

// Model + settings constants, replicated per GPU (read-only, L1/L2 resident).
struct DevModel {
  // kinematic tree: joint j moves body j+1
  int parent[NJ];          // parent body index
  int axis[NJ];            // 0/1/2 = x/y/z in the joint frame
  int depth[NB];           // base 0
  int chain_start[NJ];     // first joint of the serial chain joint j belongs to
  double Rj[NJ][9];        // joint frame in parent body frame (row-major)
  double pj[NJ][3];
  double mass[NB];
  double com[NB][3];       // body frame
  double Ib[NB][9];        // about com, body frame
  int foot_body[4];        // contact order LF, RF, LH, RH
  int foot_leg[4];         // index of the leg's first joint (joint order LF, LH, RF, RH)
  int leg_foot[4];         // inverse map: leg (first joint / 3) → foot index
  double foot_p[4][3];
  int ee_body; double ee_R[9]; double ee_p[3];
  double total_mass;
  double I_nom[9], I_nom_inv[9], c_nom[3];   // SRBD centroidalInertiaNominal, its inverse, comToBasePositionNominal
  double effort[NJ];
  double arm_pos_lower[6], arm_pos_upper[6];
  // WBC gains (wbcWigeht.cfg:7-47) and friction (task.info:346-349)
  double kp_swing, kd_swing, base_height_kp, base_height_kd, base_linear_kp, base_linear_kd, base_angular_kp, base_angular_kd;
  double arm_joint_kp[6], arm_joint_kd[6], ee_linear_kp[3], ee_linear_kd[3], ee_angular_kp[3], ee_angular_kd[3];
  double wbc_friction;
  // MPC settings (task.info:75-92,138-147,192-343)
  double Q[NX * NX], R[NU * NU];
  double Qdiag[NX]; int q_is_diag;   // Q of task.info:192-233 is diagonal: fast path when the loaded matrix really is (checked at create), dense fallback otherwise
  double Rblk[8][9], Rarm[6];   // R is block diagonal (QMInterface.cpp:274-299): 3x3 blocks per foot force / per leg, diagonal for the arm; checked at create
  double mu_ee_pos, mu_ee_ori, mu_final_ee_pos, mu_final_ee_ori;
  double friction_mu, friction_barrier_mu, friction_barrier_delta, friction_reg, friction_hess_shift;
  double pos_limit_mu, pos_limit_delta, vel_limit_mu, vel_limit_delta;
  double arm_vel_lower[6], arm_vel_upper[6];
  double lift_off_velocity, touch_down_velocity, swing_height, swing_time_scale, position_error_gain;
  double dt, time_horizon, delta_tol, g_max, g_min, alpha_decay, alpha_min, gamma_c, armijo_factor;
  double rk_c, rk_w1, rk_w2;
  // solver variant (qmb200_mpc_set_solver): 0 = multiple-shooting SQP (sqp{}, what QMController runs), 1 = multiple-shooting IPM (ipm{}: this OCP has no
  // inequality rows, so the interior-point step IS the equality-constrained Newton step; only the tolerances differ), 2 = DDP (ddp{}: single-shooting rollouts,
  // discrete Riccati backward pass, rollout line search on the merit cost + penalty * sqrt(equality SSE))
  int solver; double ddp_penalty, ddp_min_step, ddp_max_step, ddp_armijo, ddp_contraction;
  int wbc_iter_cap0, wbc_iter_cap;   // WBC iteration caps: level-0 semismooth passes (30) and active-set iterations per level (80); qmb200_wbc_set_iteration_caps (diagnostics / tests)
  double cost_tol; int sqp_iterations;   // sqp.sqpIteration (task.info:28) and costTol [upstream ocs2_sqp default 1e-4]: SqpSolver::runImpl loop + checkConvergence
};

#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
// argmax over lanes: returns (value, index) of the maximum; ties → lowest index
__device__ __forceinline__ void warp_argmax(double& v, int& idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(FULL, v, o); int oi = __shfl_xor_sync(FULL, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
}
// argmax over the ACTIVE lanes of non-negative values (ties -> lowest lane; v = -1 when no lane is active): a non-negative double orders like its bit
// pattern, so two 32-bit REDUX maxima (high word, then low word among the lanes holding the high maximum) and a ballot replace five shuffle rounds
__device__ __forceinline__ void warp_argmax_nonneg(double& v, int& idx, bool active) {
  const unsigned long long b = active ? (unsigned long long)__double_as_longlong(v) : 0ull;
  const unsigned hi = (unsigned)(b >> 32), lo = (unsigned)b;
  const unsigned mh = __reduce_max_sync(FULL, hi);
  const unsigned ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
  const unsigned m = __ballot_sync(FULL, active && hi == mh && lo == ml);
  if (m == 0u) { v = -1.0; idx = 0; return; }
  idx = __ffs(m) - 1; v = __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
}
__device__ __forceinline__ void warp_argmin(double& v, int& idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(FULL, v, o); int oi = __shfl_xor_sync(FULL, idx, o);
    if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
}

#endif  // __CUDACC__

// ---- tiny 3-vector / 3x3 helpers on plain arrays ----
QMB_HD void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
QMB_HD void cross3_add(const double* a, const double* b, double* c) {
  c[0] += a[1] * b[2] - a[2] * b[1]; c[1] += a[2] * b[0] - a[0] * b[2]; c[2] += a[0] * b[1] - a[1] * b[0];
}
QMB_HD double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
QMB_HD void matvec3(const double* M, const double* v, double* o) {
  o[0] = M[0] * v[0] + M[1] * v[1] + M[2] * v[2]; o[1] = M[3] * v[0] + M[4] * v[1] + M[5] * v[2]; o[2] = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
}
QMB_HD void matTvec3(const double* M, const double* v, double* o) {
  o[0] = M[0] * v[0] + M[3] * v[1] + M[6] * v[2]; o[1] = M[1] * v[0] + M[4] * v[1] + M[7] * v[2]; o[2] = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
}
QMB_HD void matmul3(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
// C = A * B^T
QMB_HD void matmul3_nt(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
// R = Rz(z) Ry(y) Rx(x)   (ocs2 getRotationMatrixFromZyxEulerAngles)
QMB_HD void rot_zyx(double z, double y, double x, double* R) {
  double sz, cz, sy, cy, sx, cx; sincos(z, &sz, &cz); sincos(y, &sy, &cy); sincos(x, &sx, &cx);
  R[0] = cz * cy; R[1] = cz * sy * sx - sz * cx; R[2] = cz * sy * cx + sz * sx;
  R[3] = sz * cy; R[4] = sz * sy * sx + cz * cx; R[5] = sz * sy * cx - cz * sx;
  R[6] = -sy;     R[7] = cy * sx;                R[8] = cy * cx;
}
// T: euler-ZYX rates → world angular velocity (ocs2 getMappingFromEulerAnglesZyxDerivativeToGlobalAngularVelocity)
QMB_HD void euler_rate_map(double z, double y, double* T) {
  double sz, cz, sy, cy; sincos(z, &sz, &cz); sincos(y, &sy, &cy);
  T[0] = 0; T[1] = -sz; T[2] = cy * cz; T[3] = 0; T[4] = cz; T[5] = cy * sz; T[6] = 1; T[7] = 0; T[8] = -sy;
}
// Tdot * ed  (time derivative of T along euler rates ed=(zd,yd,xd), applied to ed)
QMB_HD void euler_rate_map_dot_times(double z, double y, const double* ed, double* o) {
  double sz, cz, sy, cy; sincos(z, &sz, &cz); sincos(y, &sy, &cy);
  const double zd = ed[0], yd = ed[1];
  // d/dt of columns: col1 = (-sz, cz, 0) → (-cz zd, -sz zd, 0); col2 = (cy cz, cy sz, -sy) → (-sy yd cz - cy sz zd, -sy yd sz + cy cz zd, -cy yd)
  o[0] = (-cz * zd) * ed[1] + (-sy * yd * cz - cy * sz * zd) * ed[2];
  o[1] = (-sz * zd) * ed[1] + (-sy * yd * sz + cy * cz * zd) * ed[2];
  o[2] = (-cy * yd) * ed[2];
}
// Variants taking precomputed trig values tr = {sin z, cos z, sin y, cos y, sin x, cos x} (one warp-wide sincos pass serves all users)
QMB_HD void rot_zyx_sc(const double* tr, double* R) {
  const double sz = tr[0], cz = tr[1], sy = tr[2], cy = tr[3], sx = tr[4], cx = tr[5];
  R[0] = cz * cy; R[1] = cz * sy * sx - sz * cx; R[2] = cz * sy * cx + sz * sx;
  R[3] = sz * cy; R[4] = sz * sy * sx + cz * cx; R[5] = sz * sy * cx - cz * sx;
  R[6] = -sy;     R[7] = cy * sx;                R[8] = cy * cx;
}
QMB_HD void euler_rate_map_sc(const double* tr, double* T) {
  const double sz = tr[0], cz = tr[1], sy = tr[2], cy = tr[3];
  T[0] = 0; T[1] = -sz; T[2] = cy * cz; T[3] = 0; T[4] = cz; T[5] = cy * sz; T[6] = 1; T[7] = 0; T[8] = -sy;
}
QMB_HD void euler_rate_map_dot_times_sc(const double* tr, const double* ed, double* o) {
  const double sz = tr[0], cz = tr[1], sy = tr[2], cy = tr[3]; const double zd = ed[0], yd = ed[1];
  o[0] = (-cz * zd) * ed[1] + (-sy * yd * cz - cy * sz * zd) * ed[2];
  o[1] = (-sz * zd) * ed[1] + (-sy * yd * sz + cy * cz * zd) * ed[2];
  o[2] = (-cy * yd) * ed[2];
}
QMB_HD void inv3(const double* m, double* o) {
  const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
  const double id = 1.0 / (m[0] * c00 + m[1] * c01 + m[2] * c02);
  o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}
// rotation vector of E = L * R^T  (ocs2 rotationErrorInWorld)
QMB_HD void rotation_error_world(const double* L, const double* Rr, double* e) {
  double E[9]; matmul3_nt(L, Rr, E);
  const double w[3] = {E[7] - E[5], E[2] - E[6], E[3] - E[1]};
  const double c = 0.5 * (E[0] + E[4] + E[8] - 1.0), s = 0.5 * sqrt(dot3(w, w));
  if (s < 1e-12) { e[0] = 0.5 * w[0]; e[1] = 0.5 * w[1]; e[2] = 0.5 * w[2]; return; }
  const double k = atan2(s, c) / (2.0 * s); e[0] = k * w[0]; e[1] = k * w[1]; e[2] = k * w[2];
}
// modeNumber2StanceLeg: bit3 LF, bit2 RF, bit1 LH, bit0 RH (contact order LF,RF,LH,RH)
QMB_HD bool contact_flag(int mode, int foot) { return (mode >> (3 - foot)) & 1; }

}  // namespace qmb
