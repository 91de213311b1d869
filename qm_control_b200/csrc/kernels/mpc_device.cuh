// Per-node device functions of the MPC path (one warp per node): centroidal flow map with ANALYTIC Jacobians,
// cost terms and equality-constraint linearisation.  They restate, for the SRBD quadruped-manipulator,
//   PinocchioCentroidalDynamicsAD behind QMDynamicsAD (qm_interface/src/dynamics/QMDynamicsAD.cpp:22-33),
//   LeggedRobotStateInputQuadraticCost (include/qm_interface/cost/LeggedRobotQuadraticTrackingCost.h:34-40),
//   EndEffectorConstraint soft cost (src/constraint/EndEffectorConstraint.cpp:36-113, QMInterface.cpp:147-172),
//   arm joint soft box (QMInterface.cpp:177-259), friction-cone soft constraint (QMInterface.cpp:344-358),
//   ZeroForce / ZeroVelocity / NormalVelocity equality constraints (QMInterface.cpp:116-131,
//   NormalVelocityConstraintCppAd.cpp:37-66, QMPreComputation.cpp:50-71).
// The reference differentiates these with CppAD tapes; here the derivatives are written out by hand
// (the sparsity of the SRBD model is what makes the per-node work small enough for one warp).
#pragma once
#include "dev_common.cuh"
#include "rbd.cuh"
#include "mpc_scalar.cuh"

namespace qmb {


struct PointWs {
  KinWs kin;
  double x[NX], u[NU], f[NX];
  double Ar[9 * NX];          // rows 3:12 of df/dx
  double Bh[3 * 12];          // rows 3:6, cols 0:12 of df/du
  double T[9], Tinv[9], W[9], c[3], rcom[3], omega[3], thd[3], dom[3][3];
  double Mpc[9], Mtw[9], vp[3][3], vt[3][3], hth[3][3];
  double pf[4][3], d[4][3], Jl[4][9], al[4][9];   // Jl[i][3*j + a]: component a of leg-Jacobian column j of foot i
};


// Evaluate the flow map (and its Jacobian rows if with_jac) at (ws->x, ws->u).
// max_depth: 6 = whole tree, 3 = base + legs (the flow map does not see the arm links; only the end-effector cost does).
// The base-frame algebra is spread over lanes wherever it is data parallel (entries of 3x3 products, the three euler columns): a
// single-lane restatement costs ~4x the instructions and the LQ / line-search kernels are issue-latency bound.
template <bool with_jac>
__device__ __forceinline__ void point_eval(const DevModel* __restrict__ mdl, PointWs* ws, int lane, int lfp, int max_depth = 6) {
  rbd_kinematics<false>(mdl, ws->x + 6, (const double*)nullptr, &ws->kin, lane, max_depth);
  // fp64 division is a ~30-instruction subroutine: divide once (im), multiply everywhere
  const double m = mdl->total_mass, im = 1.0 / m; const double* tr = ws->kin.trig; const double* R = ws->kin.R[0]; const double* ha = ws->x + 3;
  if (lane == 0) { euler_rate_map_sc(tr, ws->T); inv3(ws->T, ws->Tinv); }
  if (lane < 9) {   // W = m R I_nom^{-1} R' , entry (i, jj)
    const int i = lane / 3, jj = lane - 3 * i; const double* Ii = mdl->I_nom_inv; double acc = 0.0;
#pragma unroll
    for (int bq = 0; bq < 3; ++bq) { const double rib = R[3 * i] * Ii[bq] + R[3 * i + 1] * Ii[3 + bq] + R[3 * i + 2] * Ii[6 + bq]; acc = fma(rib, R[3 * jj + bq], acc); }
    ws->W[lane] = m * acc;
  }
  if (lane < 3) { const double cv = R[3 * lane] * mdl->c_nom[0] + R[3 * lane + 1] * mdl->c_nom[1] + R[3 * lane + 2] * mdl->c_nom[2]; ws->c[lane] = cv; ws->rcom[lane] = ws->x[6 + lane] - cv; }
  __syncwarp();
  double om = 0.0; if (lane < 3) om = ws->W[3 * lane] * ha[0] + ws->W[3 * lane + 1] * ha[1] + ws->W[3 * lane + 2] * ha[2];
  const double o0 = __shfl_sync(FULL, om, 0), o1 = __shfl_sync(FULL, om, 1), o2 = __shfl_sync(FULL, om, 2);
  double thl = 0.0; if (lane < 3) { thl = ws->Tinv[3 * lane] * o0 + ws->Tinv[3 * lane + 1] * o1 + ws->Tinv[3 * lane + 2] * o2; ws->omega[lane] = om; ws->thd[lane] = thl; }
  if (with_jac) {
    const double th1 = __shfl_sync(FULL, thl, 1), th2 = __shfl_sync(FULL, thl, 2);
    if (lane < 3) {   // lane = euler column k
      const int k = lane; const double sz = tr[0], cz = tr[1], sy = tr[2], cy = tr[3]; const double omv[3] = {o0, o1, o2};
      double dT[3];
      if (k == 0) { dT[0] = -cz * th1 - cy * sz * th2; dT[1] = -sz * th1 + cy * cz * th2; dT[2] = 0.0; }
      else if (k == 1) { dT[0] = -sy * cz * th2; dT[1] = -sy * sz * th2; dT[2] = -cy * th2; }
      else { dT[0] = 0.0; dT[1] = 0.0; dT[2] = 0.0; }
      const double Tk[3] = {ws->T[k], ws->T[3 + k], ws->T[6 + k]}; double t1[3], t2[3], t3[3];
      cross3(Tk, omv, t1); cross3(Tk, ha, t2); matvec3(ws->W, t2, t3);
      double domk[3]; for (int a = 0; a < 3; ++a) { domk[a] = t1[a] - t3[a]; ws->dom[k][a] = domk[a]; }          // d omega / d theta_k
      double tc[3], vpk[3]; cross3(Tk, ws->c, tc); cross3(domk, ws->c, vpk); cross3_add(omv, tc, vpk);           // d(omega x c)/d theta_k
      for (int a = 0; a < 3; ++a) ws->vp[k][a] = vpk[a];
      const double tmp[3] = {domk[0] - dT[0], domk[1] - dT[1], domk[2] - dT[2]}; matvec3(ws->Tinv, tmp, ws->vt[k]);
    }
    if (lane < 9) {   // Mpc = -S(c) W ; Mtw = Tinv W , entry (i, jj)
      const int i = lane / 3, jj = lane - 3 * i; const double* c = ws->c; const double* W = ws->W;
      const double s0 = (i == 0) ? 0.0 : (i == 1 ? c[2] : -c[1]), s1 = (i == 0) ? -c[2] : (i == 1 ? 0.0 : c[0]), s2 = (i == 0) ? c[1] : (i == 1 ? -c[0] : 0.0);
      ws->Mpc[lane] = -(s0 * W[jj] + s1 * W[3 + jj] + s2 * W[6 + jj]);
      ws->Mtw[lane] = ws->Tinv[3 * i] * W[jj] + ws->Tinv[3 * i + 1] * W[3 + jj] + ws->Tinv[3 * i + 2] * W[6 + jj];
    }
  }
  __syncwarp();
  if (lane < 4) {
    const int i = lane; const int body = mdl->foot_body[i]; double pl[3] = {mdl->foot_p[i][0], mdl->foot_p[i][1], mdl->foot_p[i][2]}, pw[3];
    matvec3(ws->kin.R[body], pl, pw); for (int a = 0; a < 3; ++a) { pw[a] += ws->kin.p[body][a]; ws->pf[i][a] = pw[a]; ws->d[i][a] = pw[a] - ws->rcom[a]; }
    const int first = mdl->foot_leg[i];
    for (int j = 0; j < 3; ++j) { const double* S = ws->kin.S[6 + first + j]; const double* o = ws->kin.p[first + j + 1]; const double r[3] = {pw[0] - o[0], pw[1] - o[1], pw[2] - o[2]}; double col[3]; cross3(S, r, col);
      for (int a = 0; a < 3; ++a) { ws->Jl[i][3 * j + a] = col[a]; ws->al[i][3 * j + a] = S[a]; } }
    if (with_jac) { for (int a = 0; a < 3; ++a) { const double ea[3] = {a == 0 ? 1.0 : 0.0, a == 1 ? 1.0 : 0.0, a == 2 ? 1.0 : 0.0}; double col[3]; cross3(ws->d[i], ea, col); for (int r = 0; r < 3; ++r) ws->Bh[r * 12 + 3 * i + a] = col[r] * im; } }
  }
  __syncwarp();
  if (with_jac && lane < 3) {   // sum_i (T_k x d_i) x F_i / m
    const int k = lane; const double Tk[3] = {ws->T[k], ws->T[3 + k], ws->T[6 + k]}; double acc[3] = {0, 0, 0};
    for (int i = 0; i < 4; ++i) { double t[3]; cross3(Tk, ws->d[i], t); cross3_add(t, ws->u + 3 * i, acc); }
    for (int a = 0; a < 3; ++a) ws->hth[k][a] = acc[a] * im;
  }
  if (lane < NX) {
    double val;
    if (lane < 3) { val = (ws->u[lane] + ws->u[3 + lane] + ws->u[6 + lane] + ws->u[9 + lane]) * im + (lane == 2 ? -9.81 : 0.0); }
    else if (lane < 6) { const int a = lane - 3; double acc = 0.0; for (int i = 0; i < 4; ++i) { const double* d = ws->d[i]; const double* F = ws->u + 3 * i; acc += (a == 0) ? d[1] * F[2] - d[2] * F[1] : (a == 1 ? d[2] * F[0] - d[0] * F[2] : d[0] * F[1] - d[1] * F[0]); } val = acc * im; }
    else if (lane < 9) { const int a = lane - 6; const double* o = ws->omega; const double* c = ws->c; const double oc = (a == 0) ? o[1] * c[2] - o[2] * c[1] : (a == 1 ? o[2] * c[0] - o[0] * c[2] : o[0] * c[1] - o[1] * c[0]); val = ws->x[a] + oc; }
    else if (lane < 12) val = ws->thd[lane - 9];
    else val = ws->u[lane];
    ws->f[lane] = val;
  }
  __syncwarp();
  if (with_jac) {
    // Ar (9 x 30: the non-trivial rows 3:12 of df/dx = d[hdot_ang; pdot; thetadot]/dx) is two thirds zeros: fill, then lane = column writes its own non-zeros
    // (no per-element index arithmetic; the fill and the column writes are separated by a warp barrier)
    for (int e = lane; e < 9 * NX; e += 32) ws->Ar[e] = 0.0;
    __syncwarp();
    if (lane < 24) {
      const int col = lane;
      if (col < 3) ws->Ar[(3 + col) * NX + col] = 1.0;                                                         // d pdot / d h_lin = I
      else if (col < 6) { for (int a = 0; a < 3; ++a) { ws->Ar[(3 + a) * NX + col] = ws->Mpc[3 * a + col - 3]; ws->Ar[(6 + a) * NX + col] = ws->Mtw[3 * a + col - 3]; } }   // d / d h_ang
      else if (col >= 9 && col < 12) { for (int a = 0; a < 3; ++a) { ws->Ar[a * NX + col] = ws->hth[col - 9][a]; ws->Ar[(3 + a) * NX + col] = ws->vp[col - 9][a]; ws->Ar[(6 + a) * NX + col] = ws->vt[col - 9][a]; } }   // d / d theta
      else if (col >= 12) { const int j = col - 12; const int i = foot_of_leg_joint(lfp, j); const double* J = ws->Jl[i] + 3 * (j % 3); const double* F = ws->u + 3 * i;   // d hdot_ang / d q_leg = (J_j x F) / m
        ws->Ar[col] = (J[1] * F[2] - J[2] * F[1]) * im; ws->Ar[NX + col] = (J[2] * F[0] - J[0] * F[2]) * im; ws->Ar[2 * NX + col] = (J[0] * F[1] - J[1] * F[0]) * im; }
    }
    __syncwarp();
  }
}

// Target trajectory references at time t: xnom[30] (TargetTrajectories::getDesiredState().head(30)), EE pose reference
// (EndEffectorConstraint::interpolateEndEffectorPose, EndEffectorConstraint.cpp:82-113; Eigen slerp semantics).  All lanes compute the same
// scalars; lane < 30 returns its own xnom component.
struct TargetRef { double xnom; double pref[3]; double qref[4]; };
// state reference of this lane only (the end-effector pose reference is consumed by the flow kernel)
__device__ __forceinline__ double target_xnom(const double* tt, const double* ts /*[K][37]*/, int nk, double t, int lane) {
  int idx; double a; time_segment(tt, nk, t, idx, a);
  const double* l = ts + (size_t)idx * 37; const double* rr = ts + (size_t)((nk > 1) ? idx + 1 : idx) * 37;
  if (nk <= 1) a = 1.0;
  return (lane < NX) ? a * l[lane] + (1.0 - a) * rr[lane] : 0.0;
}
__device__ __forceinline__ TargetRef target_reference(const double* tt, const double* ts /*[K][37]*/, int nk, double t, int lane) {
  TargetRef r; int idx; double a; time_segment(tt, nk, t, idx, a);
  const double* l = ts + (size_t)idx * 37; const double* rr = ts + (size_t)((nk > 1) ? idx + 1 : idx) * 37;
  if (nk <= 1) a = 1.0;
  r.xnom = (lane < NX) ? a * l[lane] + (1.0 - a) * rr[lane] : 0.0;
  for (int i = 0; i < 3; ++i) r.pref[i] = a * l[30 + i] + (1.0 - a) * rr[30 + i];
  if (nk > 1) {
    const double* ql = l + 33; const double* qr = rr + 33; const double tq = 1.0 - a; double d = 0.0; for (int i = 0; i < 4; ++i) d += ql[i] * qr[i];
    const double ad = fabs(d); double s0, s1;
    if (ad >= 1.0 - 2.220446049250313e-16) { s0 = 1.0 - tq; s1 = tq; } else { const double th = acos(ad), st = sin(th); const double ist = 1.0 / st; s0 = sin((1.0 - tq) * th) * ist; s1 = sin(tq * th) * ist; }
    if (d < 0.0) s1 = -s1; for (int i = 0; i < 4; ++i) r.qref[i] = s0 * ql[i] + s1 * qr[i];
  } else { for (int i = 0; i < 4; ++i) r.qref[i] = l[33 + i]; }
  return r;
}

// ---- cost --------------------------------------------------------------------------------------------
struct CostWs { double Je[6 * 12], e[6], quat[4], pee[3]; };

// End-effector error e = [p_ee - p_ref; quaternionDistance(q_ee, q_ref)] and (optionally) its Jacobian columns.  ws must hold the kinematics at x.
template <bool with_jac>
__device__ __forceinline__ void ee_error(const DevModel* __restrict__ mdl, const PointWs* ws, CostWs* cw, const TargetRef& ref, int lane) {
  const int body = mdl->ee_body;
  if (lane == 0) {
    double R[9]; matmul3(ws->kin.R[body], mdl->ee_R, R); double pl[3] = {mdl->ee_p[0], mdl->ee_p[1], mdl->ee_p[2]}, pw[3]; matvec3(ws->kin.R[body], pl, pw);
    for (int a = 0; a < 3; ++a) { pw[a] += ws->kin.p[body][a]; cw->pee[a] = pw[a]; cw->e[a] = pw[a] - ref.pref[a]; }
    // rotation → quaternion (w,x,y,z); sign free (quadratic penalty), same q used for e and its Jacobian
    double q[4]; const double tr = R[0] + R[4] + R[8];
    if (tr > 0.0) { const double s = sqrt(tr + 1.0) * 2.0, is = 1.0 / s; q[0] = 0.25 * s; q[1] = (R[7] - R[5]) * is; q[2] = (R[2] - R[6]) * is; q[3] = (R[3] - R[1]) * is; }
    else if (R[0] > R[4] && R[0] > R[8]) { const double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2.0, is = 1.0 / s; q[0] = (R[7] - R[5]) * is; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) * is; q[3] = (R[2] + R[6]) * is; }
    else if (R[4] > R[8]) { const double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2.0, is = 1.0 / s; q[0] = (R[2] - R[6]) * is; q[1] = (R[1] + R[3]) * is; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) * is; }
    else { const double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2.0, is = 1.0 / s; q[0] = (R[3] - R[1]) * is; q[1] = (R[2] + R[6]) * is; q[2] = (R[5] + R[7]) * is; q[3] = 0.25 * s; }
    for (int i = 0; i < 4; ++i) cw->quat[i] = q[i];
    const double* rv = ref.qref; const double rw = ref.qref[3]; const double* qv = q + 1; double cr[3]; cross3(qv, rv, cr);
    for (int a = 0; a < 3; ++a) cw->e[3 + a] = q[0] * rv[a] - rw * qv[a] + cr[a];   // ocs2 quaternionDistance(q, qRef) [upstream]
  }
  __syncwarp();
  if (with_jac && lane < 12) {
    // column lane: angular direction n and linear velocity of the EE point for a unit rate of the coordinate
    double n[3] = {0, 0, 0}, lin[3] = {0, 0, 0};
    if (lane < 3) { lin[lane] = 1.0; }
    else if (lane < 6) { const int k = lane - 3; n[0] = ws->T[k]; n[1] = ws->T[3 + k]; n[2] = ws->T[6 + k]; const double r[3] = {cw->pee[0] - ws->x[6], cw->pee[1] - ws->x[7], cw->pee[2] - ws->x[8]}; cross3(n, r, lin); }
    else { const int j = 12 + lane - 6; const double* S = ws->kin.S[6 + j]; const double* o = ws->kin.p[j + 1]; n[0] = S[0]; n[1] = S[1]; n[2] = S[2]; const double r[3] = {cw->pee[0] - o[0], cw->pee[1] - o[1], cw->pee[2] - o[2]}; cross3(n, r, lin); }
    const double* q = cw->quat; const double* qv = q + 1; const double* rv = ref.qref; const double rw = ref.qref[3];
    // qdot_w = -1/2 n.qv ; qdot_v = 1/2 (qw n + n x qv) ; de = qdot_w rv - rw qdot_v + qdot_v x rv
    const double dw = -0.5 * dot3(n, qv); double dv[3]; cross3(n, qv, dv); for (int a = 0; a < 3; ++a) dv[a] = 0.5 * (q[0] * n[a] + dv[a]);
    double cr[3]; cross3(dv, rv, cr);
    for (int a = 0; a < 3; ++a) { cw->Je[a * 12 + lane] = lin[a]; cw->Je[(3 + a) * 12 + lane] = dw * rv[a] - rw * dv[a] + cr[a]; }
  }
  __syncwarp();
}

// Intermediate (or terminal) cost value; when with_quad also the quadratic model in cw (NOT scaled by dt).  `flags` = contact flags bitmask (bit i = foot i).
// Returns the cost value (lane-uniform).
// EE_PRE: the end-effector error (cw->e) and its Jacobian (cw->Je) were produced by the thread-per-node flow kernel (node_eval.cuh); WS then only needs x / u.
template <bool with_quad, bool EE_PRE = false, class WS = PointWs, class CW = CostWs>
__device__ __forceinline__ double stage_cost(const DevModel* __restrict__ mdl, const WS* ws, CW* cw, QuadWs* qw, const TargetRef& ref, int flagmask, bool terminal, int lane) {
  double value = 0.0;
  if (with_quad) { for (int e = lane; e < 144; e += 32) qw->E[e] = 0.0; for (int e = lane; e < 36; e += 32) qw->fric[e] = 0.0; if (lane < NX) { qw->qdiag[lane] = 0.0; qw->rdiag[lane] = 0.0; qw->qf[lane] = 0.0; qw->rf[lane] = 0.0; } __syncwarp(); }
  int nst = 0; for (int i = 0; i < 4; ++i) nst += (flagmask >> i) & 1;
  if (!terminal) {
    // tracking cost: 1/2 dx'Q dx + 1/2 du'R du, u_nom = weightCompensatingInput(contact flags)
    double dx = 0.0, du = 0.0;
    if (lane < NX) { dx = ws->x[lane] - ref.xnom; double un = 0.0; if (lane < 12 && (lane % 3) == 2 && ((flagmask >> (lane / 3)) & 1)) un = mdl->total_mass * 9.81 / nst; du = ws->u[lane] - un; }
    double qd = 0.0, rd = 0.0;
    if (mdl->q_is_diag) { if (lane < NX) qd = mdl->Qdiag[lane] * dx; }
    else { const double* Qr = mdl->Q + (lane < NX ? lane : 0) * NX;
#pragma unroll 6
      for (int j = 0; j < NX; ++j) qd = fma(Qr[j], __shfl_sync(FULL, dx, j), qd); }
    { // R is block diagonal (checked at create): 3x3 blocks over the 8 force / leg-joint triples, diagonal over the arm
      const int blk = lane < 24 ? lane / 3 : 0, row = lane - 3 * blk; const double* Rb = mdl->Rblk[blk] + 3 * (lane < 24 ? row : 0);
      const double d0 = __shfl_sync(FULL, du, 3 * blk), d1 = __shfl_sync(FULL, du, 3 * blk + 1), d2 = __shfl_sync(FULL, du, 3 * blk + 2);
      if (lane < 24) rd = fma(Rb[0], d0, fma(Rb[1], d1, Rb[2] * d2)); else if (lane < NU) rd = mdl->Rarm[lane - 24] * du; }
    value += 0.5 * warp_sum(lane < NX ? dx * qd + du * rd : 0.0);
    if (with_quad && lane < NX) { qw->qf[lane] = qd; qw->rf[lane] = rd; }
    __syncwarp();
  }
  // end-effector soft constraint (quadratic penalty, Gauss-Newton)
  if constexpr (!EE_PRE) ee_error<with_quad>(mdl, ws, cw, ref, lane);
  {
    const double mup = terminal ? mdl->mu_final_ee_pos : mdl->mu_ee_pos, muo = terminal ? mdl->mu_final_ee_ori : mdl->mu_ee_ori;
    double v = 0.0; for (int r = 0; r < 6; ++r) v += 0.5 * (r < 3 ? mup : muo) * cw->e[r] * cw->e[r]; value += v;
    if (with_quad) {
      for (int e = lane; e < 144; e += 32) { const int i = e / 12, j = e % 12; double s = 0.0; for (int r = 0; r < 6; ++r) s += (r < 3 ? mup : muo) * cw->Je[r * 12 + i] * cw->Je[r * 12 + j]; qw->E[e] = s; }
      if (lane < 12) { double s = 0.0; for (int r = 0; r < 6; ++r) s += (r < 3 ? mup : muo) * cw->e[r] * cw->Je[r * 12 + lane]; qw->qf[ee_col(lane)] += s; }
      __syncwarp();
    }
  }
  if (!terminal) {
    // arm joint position (state 24:30) and velocity (input 24:30) soft box, relaxed log barrier
    double bv = 0.0;
    if (lane < 12) {
      const int i = lane % 6; const bool pos = lane < 6; const double val = pos ? ws->x[24 + i] : ws->u[24 + i];
      const double lo = pos ? mdl->arm_pos_lower[i] : mdl->arm_vel_lower[i], hi = pos ? mdl->arm_pos_upper[i] : mdl->arm_vel_upper[i];
      const double mu = pos ? mdl->pos_limit_mu : mdl->vel_limit_mu, de = pos ? mdl->pos_limit_delta : mdl->vel_limit_delta;
      double a0, a1, a2, b0, b1, b2; relaxed_barrier(mu, de, val - lo, a0, a1, a2); relaxed_barrier(mu, de, hi - val, b0, b1, b2);
      bv = a0 + b0;
      if (with_quad) { if (pos) { qw->qf[24 + i] += a1 - b1; qw->qdiag[24 + i] += a2 + b2; } else { qw->rf[24 + i] += a1 - b1; qw->rdiag[24 + i] += a2 + b2; } }
    }
    // friction cone soft constraints of the stance feet; hessianDiagonalShift acts on every state and input diagonal [upstream FrictionConeConstraint]
    double shift = 0.0;
    if (lane >= 12 && lane < 16) {
      const int i = lane - 12;
      if ((flagmask >> i) & 1) {
        const double Fx = ws->u[3 * i], Fy = ws->u[3 * i + 1], Fz = ws->u[3 * i + 2]; const double n2 = Fx * Fx + Fy * Fy + mdl->friction_reg, n = sqrt(n2), in = 1.0 / n, in32 = in * in * in;
        const double h = mdl->friction_mu * Fz - n; double p0, p1, p2; relaxed_barrier(mdl->friction_barrier_mu, mdl->friction_barrier_delta, h, p0, p1, p2); bv = p0;
        if (with_quad) {
          const double g[3] = {-Fx * in, -Fy * in, mdl->friction_mu}; const double H2[9] = {-(Fy * Fy + mdl->friction_reg) * in32, Fx * Fy * in32, 0, Fx * Fy * in32, -(Fx * Fx + mdl->friction_reg) * in32, 0, 0, 0, 0};
          for (int a = 0; a < 3; ++a) { qw->rf[3 * i + a] += p1 * g[a]; for (int b = 0; b < 3; ++b) qw->fric[i * 9 + 3 * a + b] = p2 * g[a] * g[b] + p1 * H2[3 * a + b]; }
          shift = -p1 * mdl->friction_hess_shift;
        }
      }
    }
    value += warp_sum(bv);
    if (with_quad) { shift = warp_sum(shift); __syncwarp(); if (lane < NX) { qw->qdiag[lane] += shift; qw->rdiag[lane] += shift; } }
  }
  __syncwarp();
  return value;
}

// ---- equality constraints ---------------------------------------------------------------------------
// The foot-velocity rows depend on 12 state columns only: h (0:6), euler angles (9:12) and the 3 joints of the own leg → compact storage.
struct ConWs {
  double C[4][3][12];   // dg/dx rows of foot i on its support columns (stance: 3 rows; swing: row 2 only)
  double e[4][3];       // constraint values (stance: foot velocity; swing: e[i][2] = v_z - zdot_ref)
};
// foot velocity v = h_lin + omega x d + sum_j Jl_j qd_j and (optionally) its state Jacobian; lanes 0..3 (one per foot)
template <bool with_jac>
__device__ __forceinline__ void foot_velocity(const DevModel* __restrict__ mdl, const PointWs* ws, ConWs* cn, int lane) {
  if (lane < 4) {
    const int i = lane; const int first = mdl->foot_leg[i]; const double* d = ws->d[i]; const double* om = ws->omega;
    double qd[3] = {ws->u[12 + first], ws->u[12 + first + 1], ws->u[12 + first + 2]};
    double w[3] = {0, 0, 0}; for (int j = 0; j < 3; ++j) for (int a = 0; a < 3; ++a) w[a] += ws->Jl[i][3 * j + a] * qd[j];
    double v[3]; cross3(om, d, v); for (int a = 0; a < 3; ++a) { v[a] += ws->x[a] + w[a]; cn->e[i][a] = v[a]; }
    if (with_jac) {
      for (int a = 0; a < 3; ++a) for (int c = 0; c < 12; ++c) cn->C[i][a][c] = (c == a) ? 1.0 : 0.0;
      const double Sd[9] = {0, -d[2], d[1], d[2], 0, -d[0], -d[1], d[0], 0}; double SW[9]; matmul3(Sd, ws->W, SW);
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) cn->C[i][a][3 + b] = -SW[3 * a + b];
      for (int k = 0; k < 3; ++k) { const double Tk[3] = {ws->T[k], ws->T[3 + k], ws->T[6 + k]}; double t[3], col[3]; cross3(ws->dom[k], d, col); cross3(Tk, d, t); cross3_add(om, t, col); cross3_add(Tk, w, col); for (int a = 0; a < 3; ++a) cn->C[i][a][6 + k] = col[a]; }
      for (int j = 0; j < 3; ++j) {
        const double* Jj = ws->Jl[i] + 3 * j; const double* aj = ws->al[i] + 3 * j; double above[3] = {0, 0, 0}, below[3] = {0, 0, 0};
        for (int l = j + 1; l < 3; ++l) for (int a = 0; a < 3; ++a) above[a] += ws->Jl[i][3 * l + a] * qd[l];
        for (int l = 0; l <= j; ++l) for (int a = 0; a < 3; ++a) below[a] += ws->al[i][3 * l + a] * qd[l];
        double col[3]; cross3(om, Jj, col); cross3_add(aj, above, col); cross3_add(below, Jj, col);
        for (int a = 0; a < 3; ++a) cn->C[i][a][9 + j] = col[a];
      }
    }
  }
  __syncwarp();
}

}  // namespace qmb
