// ORACLE — TEST INFRASTRUCTURE ONLY.
// PARITY UNPINNED: the reference ships no tests or golden vectors for this path and its stack (OCS2 / Pinocchio / qpOASES / HPIPM) cannot be
// built here, so this restatement is not checked against reference outputs; DESIGN.md section 5 lists the pins used instead
// (known answers from the reference's own config, an independent numpy/scipy twin, finite-difference identities, tests/golden).
// ORACLE — TEST INFRASTRUCTURE ONLY.  See wbc.h.  Each function cites the reference lines it restates.
#include "wbc.h"

#include <memory>

#include "centroidal.h"

namespace orc {

WbcGains load_wbc_gains(const std::string& gains_info_file, const std::string& task_file) {
  WbcGains g;
  if (!gains_info_file.empty()) {
    auto root = info_parse_file(gains_info_file); const InfoNode* n = root->child("wbcGains");
    if (n) {
      auto get = [&](const char* k, double d) { return n->child(k) ? std::stod(n->child(k)->value) : d; };
      g.kp_swing = get("kp_swing", g.kp_swing); g.kd_swing = get("kd_swing", g.kd_swing);
      g.base_height_kp = get("baseHeightKp", g.base_height_kp); g.base_height_kd = get("baseHeightKd", g.base_height_kd);
      g.base_linear_kp = get("kp_base_linear", g.base_linear_kp); g.base_linear_kd = get("kd_base_linear", g.base_linear_kd);
      g.base_angular_kp = get("kp_base_angular", g.base_angular_kp); g.base_angular_kd = get("kd_base_angular", g.base_angular_kd);
      for (int i = 0; i < 6; ++i) { g.arm_joint_kp[i] = get(("kp_arm_joint_" + std::to_string(i + 1)).c_str(), g.arm_joint_kp[i]); g.arm_joint_kd[i] = get(("kd_arm_joint_" + std::to_string(i + 1)).c_str(), g.arm_joint_kd[i]); }
      const char* ax[3] = {"x", "y", "z"};
      for (int i = 0; i < 3; ++i) { g.ee_linear_kp[i] = get((std::string("kp_ee_linear_") + ax[i]).c_str(), g.ee_linear_kp[i]); g.ee_linear_kd[i] = get((std::string("kd_ee_linear_") + ax[i]).c_str(), g.ee_linear_kd[i]);
        g.ee_angular_kp[i] = get((std::string("kp_ee_angular_") + ax[i]).c_str(), g.ee_angular_kp[i]); g.ee_angular_kd[i] = get((std::string("kd_ee_angular_") + ax[i]).c_str(), g.ee_angular_kd[i]); }
    }
  }
  if (!task_file.empty()) { auto root = info_parse_file(task_file); g.friction_coeff = root->num_or("frictionConeTask.frictionCoefficient", g.friction_coeff); }
  return g;
}

// Task.h:29-39
Task operator+(const Task& l, const Task& r) { return {vstack(l.a, r.a), vcat(l.b, r.b), vstack(l.d, r.d), vcat(l.f, r.f)}; }
Task operator*(const Task& t, double s) { Task o = t; for (auto& x : o.a.a) x *= s; for (auto& x : o.b) x *= s; for (auto& x : o.d.a) x *= s; for (auto& x : o.f) x *= s; return o; }

// ocs2_legged_robot MotionPhaseDefinition.h modeNumber2StanceLeg [upstream]: bit3 LF, bit2 RF, bit1 LH, bit0 RH
void mode_to_contact_flags(int mode, bool flags[4]) { flags[0] = mode & 8; flags[1] = mode & 4; flags[2] = mode & 2; flags[3] = mode & 1; }

namespace {

constexpr int NV = 24, NDEC = 36;

// ---- HoQp (HoQp.cpp:12-159) ----
struct HoQp {
  Task task, stackedTasksPrev, stackedTasks; std::shared_ptr<HoQp> higher;
  bool hasEq = false, hasIneq = false; int numSlack = 0, numDec = 0, numPrevSlack = 0;
  Mat stackedZPrev, stackedZ; Vec stackedSlackPrev, xPrev; Mat h, d; Vec c, f; Vec stackedSlack, slackSol, decSol; int iterations = 0, status = 0;

  HoQp(Task t, std::shared_ptr<HoQp> hp) : task(std::move(t)), higher(std::move(hp)) { initVars(); formulate(); solve(); buildZ(); stackSlack(); }
  Vec xPolished;   // see polish()
  Vec getSolutions() const { return xPolished.empty() ? xPrev + stackedZPrev * decSol : xPolished; }   // HoQp.h:31-34

  void initVars() {  // HoQp.cpp:21-51
    numSlack = task.d.r; hasEq = task.a.r > 0; hasIneq = numSlack > 0;
    if (higher) { stackedZPrev = higher->stackedZ; stackedTasksPrev = higher->stackedTasks; stackedSlackPrev = higher->stackedSlack; xPrev = higher->getSolutions();
      numPrevSlack = higher->stackedTasks.d.r; numDec = stackedZPrev.c;
    } else { numDec = std::max(task.a.c, task.d.c); stackedTasksPrev = Task{Mat(0, numDec), Vec(), Mat(0, numDec), Vec()}; stackedZPrev = Mat::identity(numDec); stackedSlackPrev = Vec(); xPrev = Vec(numDec, 0.0); numPrevSlack = 0; }
    stackedTasks = task + stackedTasksPrev;
  }
  void formulate() {  // HoQp.cpp:53-124
    const int n = numDec, s = numSlack;
    h = Mat(n + s, n + s);
    Vec ctop(n, 0.0);
    if (hasEq) { Mat aZ = task.a * stackedZPrev; Mat zz = aZ.T() * aZ; for (int i = 0; i < n; ++i) zz(i, i) += 1e-12; h.set_block(0, 0, zz);
      Vec r = task.a * xPrev - task.b; ctop = tmul(aZ, r); }
    for (int i = 0; i < s; ++i) h(n + i, n + i) = 1.0;
    c = vcat(ctop, Vec(s, 0.0));
    d = Mat(2 * s + numPrevSlack, n + s); f = Vec(2 * s + numPrevSlack, 0.0);
    for (int i = 0; i < s; ++i) d(i, n + i) = -1.0;
    if (numPrevSlack > 0) { d.set_block(s, 0, stackedTasksPrev.d * stackedZPrev); Vec fp = stackedTasksPrev.f - stackedTasksPrev.d * xPrev + stackedSlackPrev; for (int i = 0; i < numPrevSlack; ++i) f[s + i] = fp[i]; }
    if (hasIneq) { d.set_block(s + numPrevSlack, 0, task.d * stackedZPrev); for (int i = 0; i < s; ++i) d(s + numPrevSlack + i, n + i) = -1.0; Vec fm = task.f - task.d * xPrev; for (int i = 0; i < s; ++i) f[s + numPrevSlack + i] = fm[i]; }
  }
  void solve() {  // HoQp.cpp:135-150 (qpOASES cold start → oracle active set from a feasible point)
    const int n = numDec, s = numSlack; Vec z0(n + s, 0.0);
    // feasible start: z = 0; slack v = max(0, D x_prev - f) for own rows. Previous-level rows are feasible at z = 0 by construction.
    for (int i = 0; i < s; ++i) z0[n + i] = std::max(0.0, -f[s + numPrevSlack + i]);
    QpResult r = solve_qp_active_set(h, c, d, f, z0); iterations = r.iterations; status = r.status;
    decSol = seg(r.z, 0, n); slackSol = seg(r.z, n, s);
    polish(r.active);
  }
  // Numerical refinement, not a different optimum.  The literal level problem works in the coordinates of a fullPivLu kernel (HoQp.cpp:126-133) on the
  // normal-equation Hessian Z'A'AZ + 1e-12 I (HoQp.cpp:60-66): its condition number is the square of the task's, and where a level leaves directions almost
  // free (HierarchicalMpcWbc gives the arm no task; level 2 then trades 1e4 rad/s^2 of arm acceleration against the contact forces through a 3e3-conditioned
  // block) the active-set iterate carries 1e-6..1e-4 of noise.  With the active set W the QP has identified, the level optimum is the equality-constrained least
  // squares  min |A_p x - b_p|  s.t.  A_higher x = A_higher x_prev,  D_W x = f_W + v*_W  in the 36-dim decision space; it is re-solved here with Householder
  // factorisations only, as a minimum-norm CORRECTION of the QP's iterate (free directions keep the QP's feasible choice).  Levels that own slack variables (level 0) are left as the QP returns them.
  void polish(const std::vector<int>& active) {
    if (!higher || numSlack != 0 || !hasEq) return;
    std::vector<int> rows; for (int i : active) if (i >= numSlack && i < numSlack + numPrevSlack) rows.push_back(i - numSlack);
    Mat E(stackedTasksPrev.a.r + (int)rows.size(), numDecX()); Vec e(E.r, 0.0);
    Vec ax = stackedTasksPrev.a.r ? stackedTasksPrev.a * xPrev : Vec();
    for (int i = 0; i < stackedTasksPrev.a.r; ++i) { for (int j = 0; j < E.c; ++j) E(i, j) = stackedTasksPrev.a(i, j); e[i] = ax[i]; }
    for (size_t k = 0; k < rows.size(); ++k) { const int i = rows[k], o = stackedTasksPrev.a.r + (int)k; for (int j = 0; j < E.c; ++j) E(o, j) = stackedTasksPrev.d(i, j); e[o] = stackedTasksPrev.f[i] + stackedSlackPrev[i]; }
    // correction form: x = x_qp + delta with the minimum-norm delta, so that directions the level leaves free keep the QP's (feasible) choice
    const Vec xq = xPrev + stackedZPrev * decSol; const Vec delta = constrained_lstsq(task.a, task.b - task.a * xq, E, e - E * xq);
    xPolished = xq + delta;
  }
  int numDecX() const { return (int)xPrev.size(); }
  void buildZ() {  // HoQp.cpp:126-133
    if (hasEq) { FullPivLU lu(task.a * stackedZPrev); stackedZ = stackedZPrev * lu.kernel(); } else stackedZ = stackedZPrev;
  }
  void stackSlack() { stackedSlack = higher ? vcat(higher->stackedSlack, slackSol) : slackSol; }  // HoQp.cpp:152-158
};

Mat rows(const Mat& A, int r0, int nr) { return A.block(r0, 0, nr, A.c); }
// ocs2 rotationErrorInWorld(lhs, rhs) = rotationMatrixToRotationVector(lhs * rhs^T) [upstream RotationTransforms.h]
V3<double> rotation_error_in_world(const M3<double>& lhs, const M3<double>& rhs) {
  M3<double> E = lhs * transpose(rhs); V3<double> w(E(2, 1) - E(1, 2), E(0, 2) - E(2, 0), E(1, 0) - E(0, 1));
  double c = 0.5 * (E(0, 0) + E(1, 1) + E(2, 2) - 1.0); double s = 0.5 * std::sqrt(dot(w, w));
  double ang = std::atan2(s, c); if (s < 1e-12) return 0.5 * w; return (ang / (2.0 * s)) * w; }

}  // namespace

Vec wbc_update(const Model& model, const WbcGains& g, const double* xd, const double* ud, const double* rbd, int mode, double period, double time,
               double* input_last, int variant, WbcDebug* dbg) {
  // WbcBase::update (WbcBase.cpp:118-132)
  bool flag[4]; mode_to_contact_flags(mode, flag); int nc = 0; for (bool b : flag) nc += b;
  // updateMeasured (WbcBase.cpp:134-191)
  double qM[NV], vM[NV];
  for (int i = 0; i < 3; ++i) { qM[i] = rbd[3 + i]; qM[3 + i] = rbd[i]; vM[i] = rbd[NV + 3 + i]; }
  { // getEulerAnglesZyxDerivativesFromGlobalAngularVelocity = T^{-1} w
    M3<double> Tm = euler_rate_map<double>(qM[3], qM[4]); V3<double> wv(rbd[NV], rbd[NV + 1], rbd[NV + 2]); V3<double> ed = inverse3(Tm) * wv; for (int i = 0; i < 3; ++i) vM[3 + i] = ed[i]; }
  for (int j = 0; j < NJ; ++j) { qM[6 + j] = rbd[6 + j]; vM[6 + j] = rbd[NV + 6 + j]; }
  RbdData me; compute_rbd(model, qM, vM, me, 1);
  // updateDesired (WbcBase.cpp:193-226)
  double qD[NV], vD[NV]; for (int i = 0; i < NV; ++i) qD[i] = xd[6 + i];
  pinocchio_joint_velocity<double>(model, xd, ud, vD);
  Vec jointAccel(NJ); for (int j = 0; j < NJ; ++j) jointAccel[j] = (ud[12 + j] - input_last[12 + j]) / period;
  for (int i = 0; i < NU; ++i) input_last[i] = ud[i];
  SrbdBase<double> sb = srbd_base<double>(model, qD);     // Ab, AbInv bound BEFORE dccrba → SRBD
  RbdData de; compute_rbd(model, qD, vD, de, 2);          // dccrba → full-model Ag, dAg, com (WbcBase.cpp:219)
  Vec cmr(6, 0.0);                                        // m * getNormalizedCentroidalMomentumRate with the full-model COM
  { V3<double> lin(0, 0, -9.81 * model.mass), ang;
    for (int i = 0; i < 4; ++i) { V3<double> F(ud[3 * i], ud[3 * i + 1], ud[3 * i + 2]); lin = lin + F; ang = ang + cross(de.foot_pos[i] - de.com, F); }
    for (int i = 0; i < 3; ++i) { cmr[i] = lin[i]; cmr[3 + i] = ang[i]; } }
  for (int i = 0; i < 6; ++i) { cmr[i] -= de.dAg_v[i]; for (int j = 0; j < NJ; ++j) cmr[i] -= de.Ag(i, 6 + j) * jointAccel[j]; }
  Vec baseAcc(6);
  { V3<double> l(cmr[0], cmr[1], cmr[2]), a(cmr[3], cmr[4], cmr[5]); V3<double> wd = sb.A22inv * a; V3<double> vl = (1.0 / model.mass) * (l - sb.A12 * wd);
    for (int i = 0; i < 3; ++i) { baseAcc[i] = vl[i]; baseAcc[3 + i] = wd[i]; } }
  Vec vMv(vM, vM + NV);
  if (dbg) { dbg->q_meas.assign(qM, qM + NV); dbg->v_meas = vMv; dbg->q_des.assign(qD, qD + NV); dbg->v_des.assign(vD, vD + NV); dbg->base_acc_des = baseAcc; }

  const Mat& M = me.M; const Mat& J = me.Jfoot; const Mat& dJ = me.dJfoot;
  // formulateFloatingBaseEomTask (WbcBase.cpp:338-356)
  Task eom; eom.a = Mat(6, NDEC); eom.b = Vec(6);
  for (int i = 0; i < 6; ++i) { for (int j = 0; j < NV; ++j) eom.a(i, j) = M(i, j); for (int k = 0; k < 12; ++k) eom.a(i, NV + k) = -J(k, i); eom.b[i] = -me.nle[i]; }
  // formulateTorqueLimitsTask (WbcBase.cpp:360-383); limits from URDF effort: leg = joints 0..2, arm = last 6 (WbcBase.cpp:567-572)
  Task tl; tl.d = Mat(2 * NJ, NDEC); tl.f = Vec(2 * NJ);
  { double lim[NJ]; for (int l = 0; l < 4; ++l) for (int k = 0; k < 3; ++k) lim[3 * l + k] = model.joint[k].effort; for (int k = 0; k < 6; ++k) lim[12 + k] = model.joint[12 + k].effort;
    for (int i = 0; i < NJ; ++i) { for (int j = 0; j < NV; ++j) { tl.d(i, j) = M(6 + i, j); tl.d(NJ + i, j) = -M(6 + i, j); } for (int k = 0; k < 12; ++k) { tl.d(i, NV + k) = -J(k, 6 + i); tl.d(NJ + i, NV + k) = J(k, 6 + i); }
      tl.f[i] = lim[i] - me.nle[6 + i]; tl.f[NJ + i] = lim[i] + me.nle[6 + i]; } }
  // formulateNoContactMotionTask (WbcBase.cpp:386-401)
  Task ncm; ncm.a = Mat(3 * nc, NDEC); ncm.b = Vec(3 * nc);
  { int j = 0; for (int i = 0; i < 4; ++i) if (flag[i]) { for (int r = 0; r < 3; ++r) { double s = 0; for (int k = 0; k < NV; ++k) { ncm.a(3 * j + r, k) = J(3 * i + r, k); s += dJ(3 * i + r, k) * vM[k]; } ncm.b[3 * j + r] = -s; } ++j; } }
  // formulateFrictionConeTask (WbcBase.cpp:407-437)
  Task fc; fc.a = Mat(3 * (4 - nc), NDEC); fc.b = Vec(3 * (4 - nc), 0.0); fc.d = Mat(5 * nc + 3 * (4 - nc), NDEC); fc.f = Vec(fc.d.r, 0.0);
  { int j = 0; for (int i = 0; i < 4; ++i) if (!flag[i]) { for (int r = 0; r < 3; ++r) fc.a(3 * j + r, NV + 3 * i + r) = 1.0; ++j; }
    const double mu = g.friction_coeff; const double pyr[5][3] = {{0, 0, -1}, {1, 0, -mu}, {-1, 0, -mu}, {0, 1, -mu}, {0, -1, -mu}};
    j = 0; for (int i = 0; i < 4; ++i) if (flag[i]) { for (int r = 0; r < 5; ++r) for (int cc = 0; cc < 3; ++cc) fc.d(5 * j + r, NV + 3 * i + cc) = pyr[r][cc]; ++j; } }
  Task task0 = eom + tl + ncm + fc;

  // formulateBaseHeightMotionTask (WbcBase.cpp:296-308)
  Task bh; bh.a = Mat(1, NDEC); bh.a(0, 2) = 1.0; bh.b = Vec{baseAcc[2] + g.base_height_kp * (qD[2] - qM[2]) + g.base_height_kd * (vD[2] - vM[2])};
  // formulateBaseAngularMotionTask (WbcBase.cpp:258-293)
  Task ba; ba.a = Mat(3, NDEC); ba.b = Vec(3);
  { for (int r = 0; r < 3; ++r) for (int k = 0; k < NV; ++k) ba.a(r, k) = me.Jbase(3 + r, k);
    M3<double> TmM = euler_rate_map<double>(qM[3], qM[4]);
    V3<double> wM = TmM * V3<double>(vM[3], vM[4], vM[5]); V3<double> wDes = TmM * V3<double>(vD[3], vD[4], vD[5]);
    M3<double> Rm = rot_zyx<double>(qM[3], qM[4], qM[5]); M3<double> Rr = rot_zyx<double>(qD[3], qD[4], qD[5]);
    V3<double> err = rotation_error_in_world(Rr, Rm);
    // getGlobalAngularAccelerationFromEulerAnglesZyxDerivatives(eulerMeasured, vDesired euler rates, baseAccDesired euler acc) = T edd + Tdot(ed) ed
    Jet2 ez(qM[3], vD[3], baseAcc[3]), ey(qM[4], vD[4], baseAcc[4]);
    M3<Jet2> Tj = euler_rate_map<Jet2>(ez, ey); V3<Jet2> edj(Jet2(vD[3], baseAcc[3], 0), Jet2(vD[4], baseAcc[4], 0), Jet2(vD[5], baseAcc[5], 0)); V3<Jet2> wj = Tj * edj;
    V3<double> accDes(wj.x.d1, wj.y.d1, wj.z.d1);
    for (int r = 0; r < 3; ++r) { double s = 0; for (int k = 0; k < NV; ++k) s += me.dJbase(3 + r, k) * vM[k]; ba.b[r] = accDes[r] + g.base_angular_kp * err[r] + g.base_angular_kd * (wDes[r] - wM[r]) - s; } }
  // formulateBaseLinearMotionTask (WbcBase.cpp:228-240)
  Task bl; bl.a = Mat(2, NDEC); bl.a(0, 0) = bl.a(1, 1) = 1.0; bl.b = Vec(2);
  for (int i = 0; i < 2; ++i) bl.b[i] = baseAcc[i] + g.base_linear_kp * (qD[i] - qM[i]) + g.base_linear_kd * (vD[i] - vM[i]);
  // formulateSwingLegTask (WbcBase.cpp:311-334)
  Task sw; sw.a = Mat(3 * (4 - nc), NDEC); sw.b = Vec(3 * (4 - nc));
  { int j = 0; for (int i = 0; i < 4; ++i) if (!flag[i]) { for (int r = 0; r < 3; ++r) { double acc = g.kp_swing * (de.foot_pos[i][r] - me.foot_pos[i][r]) + g.kd_swing * (de.foot_vel[i][r] - me.foot_vel[i][r]);
        double s = 0; for (int k = 0; k < NV; ++k) { sw.a(3 * j + r, k) = J(3 * i + r, k); s += dJ(3 * i + r, k) * vM[k]; } sw.b[3 * j + r] = acc - s; } ++j; } }
  // formulateArmJointNomalTrackingTask (WbcBase.cpp:439-465)
  Task aj; aj.a = Mat(6, NDEC); aj.b = Vec(6);
  for (int i = 0; i < 6; ++i) { aj.a(i, NV - 6 + i) = 1.0; aj.b[i] = g.arm_joint_kp[i] * (qD[NV - 6 + i] - qM[NV - 6 + i]) + g.arm_joint_kd[i] * (vD[NV - 6 + i] - vM[NV - 6 + i]); }
  // formulateEeLinearMotionTrackingTask (WbcBase.cpp:467-492)
  Task el; el.a = Mat(3, NDEC); el.b = Vec(3);
  for (int r = 0; r < 3; ++r) { double s = 0; for (int k = 0; k < NV; ++k) { el.a(r, k) = me.Jee(r, k); s += me.dJee(r, k) * vM[k]; }
    el.b[r] = g.ee_linear_kp[r] * (de.ee_pos[r] - me.ee_pos[r]) + g.ee_linear_kd[r] * (de.ee_vel[r] - me.ee_vel[r]) - s; }
  // formulateEeAngularMotionTrackingTask (WbcBase.cpp:494-531): columns 3:6 of J and dJ angular rows zeroed, desired angular velocity unused
  Task ea; ea.a = Mat(3, NDEC); ea.b = Vec(3);
  { V3<double> err = rotation_error_in_world(de.ee_rot, me.ee_rot);
    for (int r = 0; r < 3; ++r) { double s = 0; for (int k = 0; k < NV; ++k) { const bool z = (k >= 3 && k < 6); ea.a(r, k) = z ? 0.0 : me.Jee(3 + r, k); s += (z ? 0.0 : me.dJee(3 + r, k)) * vM[k]; }
      ea.b[r] = g.ee_angular_kp[r] * err[r] + g.ee_angular_kd[r] * (-me.ee_angvel[r]) - s; } }
  // formulateContactForceTask (WbcBase.cpp:534-546)
  Task cf; cf.a = Mat(12, NDEC); cf.b = Vec(12); for (int i = 0; i < 12; ++i) { cf.a(i, NV + i) = 1.0; cf.b[i] = ud[i]; }

  Task t1, t2; bool use_init = false;
  if (variant == WBC_HIERARCHICAL) {  // HierarchicalWbc.cpp:23-43
    t1 = bh + ba + el + ea + sw * 100.0; t2 = cf + bl; use_init = time < 10;
  } else {                            // HierarchicalMpcWbc.cpp:23-33
    t1 = bh + ba + bl + sw * 100.0; t2 = cf;
  }
  auto l0 = std::make_shared<HoQp>(task0, nullptr);
  auto l1 = std::make_shared<HoQp>(use_init ? aj : t1, l0);
  HoQp l2(t2, l1);
  Vec x = l2.getSolutions();
  if (dbg) { dbg->hoqp_iterations[0] = l0->iterations; dbg->hoqp_iterations[1] = l1->iterations; dbg->hoqp_iterations[2] = l2.iterations; dbg->qp_status = l0->status | l1->status | l2.status;
    dbg->level_solutions = {l0->getSolutions(), l1->getSolutions(), x}; }
  // updateCmd (WbcBase.cpp:548-563)
  Vec cmd(NDEC + NJ);
  for (int i = 0; i < NDEC; ++i) cmd[i] = x[i];
  for (int i = 0; i < NJ; ++i) { double s = me.nle[6 + i]; for (int j = 0; j < NV; ++j) s += M(6 + i, j) * x[j]; for (int k = 0; k < 12; ++k) s -= J(k, 6 + i) * x[NV + k]; cmd[NDEC + i] = s; }
  return cmd;
}

}  // namespace orc
