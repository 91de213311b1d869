// ORACLE — TEST INFRASTRUCTURE ONLY.
// PARITY UNPINNED: the reference ships no tests or golden vectors for this path and its stack (OCS2 / Pinocchio / qpOASES / HPIPM) cannot be
// built here, so this restatement is not checked against reference outputs; DESIGN.md section 5 lists the pins used instead
// (known answers from the reference's own config, an independent numpy/scipy twin, finite-difference identities, tests/golden).
// ORACLE — TEST INFRASTRUCTURE ONLY.  See mpc.h.
#include "mpc.h"

#include <cmath>
#include <limits>

#include "centroidal.h"
#include "wbc.h"

namespace orc {

namespace {
constexpr double kWeakEps = 1e-6;  // ocs2 numeric_traits::weakEpsilon — interval start/end shift at event nodes [recalled]
using D60 = Dual<60, double>;

// ocs2::lookup::findIndexInTimeArray = std::lower_bound distance
int find_index(const std::vector<double>& times, double t) { return (int)(std::lower_bound(times.begin(), times.end(), t) - times.begin()); }
// ocs2::LinearInterpolation::timeSegment → (index, alpha), value = alpha*v[index] + (1-alpha)*v[index+1]
std::pair<int, double> time_segment(double t, const std::vector<double>& times) {
  if (times.size() <= 1) return {0, 1.0};
  int part = find_index(times, t); int index = (part != 0 || t != times.front()) ? part - 1 : 0;
  const int last = (int)times.size() - 1;
  if (index >= 0) {
    if (index < last) { const double len = times[index + 1] - times[index], till = times[index + 1] - t;
      if (len > 2.0 * std::numeric_limits<double>::epsilon()) return {index, till / len};
      return {index, till > 0.5 * len ? 1.0 : 0.0}; }
    return {std::max(last - 1, 0), 0.0};
  }
  return {0, 1.0};
}
Vec interpolate(double t, const std::vector<double>& times, const std::vector<Vec>& vals) {
  if (vals.size() == 1) return vals[0];
  auto ia = time_segment(t, times); const Vec& a = vals[ia.first]; const Vec& b = vals[std::min<size_t>(ia.first + 1, vals.size() - 1)];
  Vec r(a.size()); for (size_t i = 0; i < a.size(); ++i) r[i] = ia.second * a[i] + (1.0 - ia.second) * b[i]; return r;
}

double barrier(double mu, double delta, double h, int order) {  // ocs2 RelaxedBarrierPenalty [recalled]
  if (h > delta) return order == 0 ? -mu * std::log(h) : (order == 1 ? -mu / h : mu / (h * h));
  const double t = (h - 2.0 * delta) / delta;
  return order == 0 ? mu * (-std::log(delta) + 0.5 * t * t - 0.5) : (order == 1 ? mu * (h - 2.0 * delta) / (delta * delta) : mu / (delta * delta));
}

// ocs2 weightCompensatingInput
void weight_compensating_input(const Model& m, const bool flags[4], double* u) {
  for (int i = 0; i < NU; ++i) u[i] = 0.0; int n = 0; for (int i = 0; i < 4; ++i) n += flags[i];
  if (n > 0) for (int i = 0; i < 4; ++i) if (flags[i]) u[3 * i + 2] = m.mass * 9.81 / n;
}

// rotation matrix → quaternion (w,x,y,z), Shepperd; the global sign is irrelevant for the quadratic penalty
template <class T> void mat_to_quat(const M3<T>& R, T q[4]) {
  const double tr = value_of(R(0, 0)) + value_of(R(1, 1)) + value_of(R(2, 2));
  if (tr > 0) { T s = sqrt(R(0, 0) + R(1, 1) + R(2, 2) + 1.0) * 2.0; q[0] = s * 0.25; q[1] = (R(2, 1) - R(1, 2)) / s; q[2] = (R(0, 2) - R(2, 0)) / s; q[3] = (R(1, 0) - R(0, 1)) / s; }
  else if (value_of(R(0, 0)) > value_of(R(1, 1)) && value_of(R(0, 0)) > value_of(R(2, 2))) { T s = sqrt(R(0, 0) - R(1, 1) - R(2, 2) + 1.0) * 2.0; q[0] = (R(2, 1) - R(1, 2)) / s; q[1] = s * 0.25; q[2] = (R(0, 1) + R(1, 0)) / s; q[3] = (R(0, 2) + R(2, 0)) / s; }
  else if (value_of(R(1, 1)) > value_of(R(2, 2))) { T s = sqrt(R(1, 1) - R(0, 0) - R(2, 2) + 1.0) * 2.0; q[0] = (R(0, 2) - R(2, 0)) / s; q[1] = (R(0, 1) + R(1, 0)) / s; q[2] = s * 0.25; q[3] = (R(1, 2) + R(2, 1)) / s; }
  else { T s = sqrt(R(2, 2) - R(0, 0) - R(1, 1) + 1.0) * 2.0; q[0] = (R(1, 0) - R(0, 1)) / s; q[1] = (R(0, 2) + R(2, 0)) / s; q[2] = (R(1, 2) + R(2, 1)) / s; q[3] = s * 0.25; }
}
// EndEffectorConstraint::getValue (EndEffectorConstraint.cpp:40-53): [p_ee - p_ref; quaternionDistance(q_ee, q_ref)]
template <class T> void ee_error(const Model& m, const T* x, const double* pref, const double* qref_xyzw, T e[6]) {
  Kin<T> k; forward_kinematics<T>(m, x + 6, k); V3<T> p = frame_pos(m, k, m.ee_frame); M3<T> R = frame_rot(m, k, m.ee_frame);
  for (int i = 0; i < 3; ++i) e[i] = p[i] - pref[i];
  T q[4]; mat_to_quat<T>(R, q); const double rw = qref_xyzw[3]; V3<T> qv(q[1], q[2], q[3]); V3<T> rv; rv.x = T(qref_xyzw[0]); rv.y = T(qref_xyzw[1]); rv.z = T(qref_xyzw[2]);
  V3<T> cr = cross(qv, rv);
  for (int i = 0; i < 3; ++i) e[3 + i] = q[0] * rv[i] - rw * qv[i] + cr[i];   // ocs2 quaternionDistance(q, qRef) [recalled]
}
// EndEffectorConstraint::interpolateEndEffectorPose (EndEffectorConstraint.cpp:82-113), Eigen slerp semantics
void ee_reference(const TargetTrajectories& tt, double t, double pref[3], double qref[4]) {
  if (tt.states.size() > 1) {
    auto ia = time_segment(t, tt.times); const double a = ia.second; const Vec& l = tt.states[ia.first]; const Vec& r = tt.states[ia.first + 1];
    for (int i = 0; i < 3; ++i) pref[i] = a * l[30 + i] + (1.0 - a) * r[30 + i];
    const double* ql = &l[33]; const double* qr = &r[33]; const double tq = 1.0 - a;
    double d = 0; for (int i = 0; i < 4; ++i) d += ql[i] * qr[i]; const double ad = std::fabs(d); double s0, s1;
    if (ad >= 1.0 - std::numeric_limits<double>::epsilon()) { s0 = 1.0 - tq; s1 = tq; } else { double th = std::acos(ad), st = std::sin(th); s0 = std::sin((1.0 - tq) * th) / st; s1 = std::sin(tq * th) / st; }
    if (d < 0) s1 = -s1; for (int i = 0; i < 4; ++i) qref[i] = s0 * ql[i] + s1 * qr[i];
  } else { const Vec& s = tt.states[0]; for (int i = 0; i < 3; ++i) pref[i] = s[30 + i]; for (int i = 0; i < 4; ++i) qref[i] = s[33 + i]; }
}

struct Quad { double f = 0; Vec q, r; Mat Q, R, P; Quad() : q(NX, 0.0), r(NU, 0.0), Q(NX, NX), R(NU, NU), P(NU, NX) {} };

// intermediate cost L(x,u,t) incl. soft constraints; quadratic model if quad != nullptr (approximateCost)
double stage_cost(const Model& m, const MpcSettings& s, const TargetTrajectories& tt, double t, const bool flags[4], const double* x, const double* u, Quad* quad) {
  double f = 0;
  // LeggedRobotStateInputQuadraticCost (LeggedRobotQuadraticTrackingCost.h:34-40)
  Vec xn = interpolate(t, tt.times, tt.states); double un[NU]; weight_compensating_input(m, flags, un);
  Vec dx(NX), du(NU); for (int i = 0; i < NX; ++i) dx[i] = x[i] - xn[i]; for (int i = 0; i < NU; ++i) du[i] = u[i] - un[i];
  Vec Qdx = s.Q * dx, Rdu = s.R * du; f += 0.5 * dot(dx, Qdx) + 0.5 * dot(du, Rdu);
  if (quad) { quad->Q = s.Q; quad->R = s.R; quad->q = Qdx; quad->r = Rdu; }
  // end-effector soft constraint (QMInterface.cpp:147-172): quadratic penalties
  { double pref[3], qref[4]; ee_reference(tt, t, pref, qref);
    if (quad) { Dual<30, double> xd[NX], e[6]; for (int i = 0; i < NX; ++i) xd[i] = Dual<30, double>::variable(x[i], i); ee_error(m, xd, pref, qref, e);
      for (int c = 0; c < 6; ++c) { const double mu = c < 3 ? s.mu_ee_pos : s.mu_ee_ori; f += 0.5 * mu * e[c].v * e[c].v;
        for (int i = 0; i < NX; ++i) { quad->q[i] += mu * e[c].v * e[c].d[i]; for (int j = 0; j < NX; ++j) quad->Q(i, j) += mu * e[c].d[i] * e[c].d[j]; } }
    } else { double e[6]; ee_error<double>(m, x, pref, qref, e); for (int c = 0; c < 6; ++c) f += 0.5 * (c < 3 ? s.mu_ee_pos : s.mu_ee_ori) * e[c] * e[c]; } }
  // arm joint position / velocity soft box (QMInterface.cpp:177-259); the constant offset of initializeOffset is dropped (cancels in every comparison)
  for (int i = 0; i < 6; ++i) {
    { const double hl = x[24 + i] - s.arm_pos_lower[i], hu = s.arm_pos_upper[i] - x[24 + i]; f += barrier(s.pos_limit_mu, s.pos_limit_delta, hl, 0) + barrier(s.pos_limit_mu, s.pos_limit_delta, hu, 0);
      if (quad) { quad->q[24 + i] += barrier(s.pos_limit_mu, s.pos_limit_delta, hl, 1) - barrier(s.pos_limit_mu, s.pos_limit_delta, hu, 1); quad->Q(24 + i, 24 + i) += barrier(s.pos_limit_mu, s.pos_limit_delta, hl, 2) + barrier(s.pos_limit_mu, s.pos_limit_delta, hu, 2); } }
    { const double hl = u[24 + i] - s.arm_vel_lower[i], hu = s.arm_vel_upper[i] - u[24 + i]; f += barrier(s.vel_limit_mu, s.vel_limit_delta, hl, 0) + barrier(s.vel_limit_mu, s.vel_limit_delta, hu, 0);
      if (quad) { quad->r[24 + i] += barrier(s.vel_limit_mu, s.vel_limit_delta, hl, 1) - barrier(s.vel_limit_mu, s.vel_limit_delta, hu, 1); quad->R(24 + i, 24 + i) += barrier(s.vel_limit_mu, s.vel_limit_delta, hl, 2) + barrier(s.vel_limit_mu, s.vel_limit_delta, hu, 2); } }
  }
  // friction cone soft constraints (QMInterface.cpp:344-358; FrictionConeConstraint [upstream, recalled]) — active in contact only
  for (int i = 0; i < 4; ++i) if (flags[i]) {
    const double Fx = u[3 * i], Fy = u[3 * i + 1], Fz = u[3 * i + 2]; const double n2 = Fx * Fx + Fy * Fy + s.friction_reg, n = std::sqrt(n2), n32 = n * n2;
    const double h = s.friction_mu * Fz - n; f += barrier(s.friction_barrier_mu, s.friction_barrier_delta, h, 0);
    if (quad) { const double p1 = barrier(s.friction_barrier_mu, s.friction_barrier_delta, h, 1), p2 = barrier(s.friction_barrier_mu, s.friction_barrier_delta, h, 2);
      const double g[3] = {-Fx / n, -Fy / n, s.friction_mu}; double H2[3][3] = {{-(Fy * Fy + s.friction_reg) / n32, Fx * Fy / n32, 0}, {Fx * Fy / n32, -(Fx * Fx + s.friction_reg) / n32, 0}, {0, 0, 0}};
      for (int a = 0; a < 3; ++a) { quad->r[3 * i + a] += p1 * g[a]; for (int b = 0; b < 3; ++b) quad->R(3 * i + a, 3 * i + b) += p2 * g[a] * g[b] + p1 * H2[a][b]; }
      for (int a = 0; a < NU; ++a) quad->R(a, a) += p1 * (-s.friction_hess_shift);   // hessianDiagonalShift on the whole input diagonal
      for (int a = 0; a < NX; ++a) quad->Q(a, a) += p1 * (-s.friction_hess_shift);   // and on the state diagonal
    }
  }
  if (quad) quad->f = f;
  return f;
}
// final cost: finalEndEffector soft constraint only (QMInterface.cpp:104)
double final_cost(const Model& m, const MpcSettings& s, const TargetTrajectories& tt, double t, const double* x, Quad* quad) {
  double f = 0; double pref[3], qref[4]; ee_reference(tt, t, pref, qref);
  if (quad) { *quad = Quad(); Dual<30, double> xd[NX], e[6]; for (int i = 0; i < NX; ++i) xd[i] = Dual<30, double>::variable(x[i], i); ee_error(m, xd, pref, qref, e);
    for (int c = 0; c < 6; ++c) { const double mu = c < 3 ? s.mu_final_ee_pos : s.mu_final_ee_ori; f += 0.5 * mu * e[c].v * e[c].v;
      for (int i = 0; i < NX; ++i) { quad->q[i] += mu * e[c].v * e[c].d[i]; for (int j = 0; j < NX; ++j) quad->Q(i, j) += mu * e[c].d[i] * e[c].d[j]; } }
    quad->f = f;
  } else { double e[6]; ee_error<double>(m, x, pref, qref, e); for (int c = 0; c < 6; ++c) f += 0.5 * (c < 3 ? s.mu_final_ee_pos : s.mu_final_ee_ori) * e[c] * e[c]; }
  return f;
}

// state-input equality constraints g(x,u,t) (QMInterface.cpp:116-131): per foot zeroVelocity (stance) or zeroForce + normalVelocity (swing)
template <class T> int equality_constraints(const Model& m, const MpcSettings& s, const ModeSchedule& sched, double t, const bool flags[4], const T* x, const T* u, T* g) {
  int n = 0;
  for (int i = 0; i < 4; ++i) {
    if (flags[i]) { V3<T> v = frame_velocity<T>(m, x, u, m.foot_frame[i]); for (int a = 0; a < 3; ++a) g[n++] = v[a]; }
    else { for (int a = 0; a < 3; ++a) g[n++] = u[3 * i + a];
      V3<T> v = frame_velocity<T>(m, x, u, m.foot_frame[i]); double zp, zv; swing_reference(s, sched, i, t, &zp, &zv);
      T val = v.z - zv;                                                     // config.b = -zvel_ref, Av = [0 0 1] (QMPreComputation.cpp:56-62)
      if (s.position_error_gain != 0.0) { Kin<T> k; forward_kinematics<T>(m, x + 6, k); V3<T> p = frame_pos(m, k, m.foot_frame[i]); val = val + s.position_error_gain * (p.z - zp); }
      g[n++] = val; }
  }
  return n;
}

// discrete RK2 map
template <class T> void rk2_step(const Model& m, const MpcSettings& s, const T* x, const T* u, double dt, T* xn) {
  T k1[NX], k2[NX], x2[NX]; flow_map<T>(m, x, u, k1); for (int i = 0; i < NX; ++i) x2[i] = x[i] + (s.rk_c * dt) * k1[i]; flow_map<T>(m, x2, u, k2);
  for (int i = 0; i < NX; ++i) xn[i] = x[i] + (dt * s.rk_w1) * k1[i] + (dt * s.rk_w2) * k2[i];
}

double interval_start(const NodeInfo& n) { return n.event == 2 ? n.t + kWeakEps : n.t; }
double interval_end(const NodeInfo& n) { return n.event == 1 ? n.t - kWeakEps : n.t; }

struct Perf { double cost = 0, dyn = 0, eq = 0; };
// multiple_shooting computePerformance [recalled]
Perf compute_performance(const Model& m, const MpcSettings& s, const ModeSchedule& sched, const TargetTrajectories& tt, const std::vector<NodeInfo>& grid, const double* x0,
                         const std::vector<Vec>& x, const std::vector<Vec>& u) {
  Perf p; const int N = (int)grid.size() - 1;
  for (int i = 0; i < NX; ++i) { const double d = x0[i] - x[0][i]; p.dyn += d * d; }
  for (int k = 0; k < N; ++k) {
    if (grid[k].event == 1) { for (int i = 0; i < NX; ++i) { const double d = x[k][i] - x[k + 1][i]; p.dyn += d * d; } continue; }
    const double t = interval_start(grid[k]), dt = interval_end(grid[k + 1]) - t; bool fl[4]; mode_to_contact_flags(mode_at_time(sched, t), fl);
    double xn[NX]; rk2_step<double>(m, s, x[k].data(), u[k].data(), dt, xn); double ss = 0; for (int i = 0; i < NX; ++i) { const double d = xn[i] - x[k + 1][i]; ss += d * d; } p.dyn += dt * ss;
    p.cost += dt * stage_cost(m, s, tt, t, fl, x[k].data(), u[k].data(), nullptr);
    double g[16]; const int ng = equality_constraints<double>(m, s, sched, t, fl, x[k].data(), u[k].data(), g); double gs = 0; for (int i = 0; i < ng; ++i) gs += g[i] * g[i]; p.eq += dt * gs;
  }
  p.cost += final_cost(m, s, tt, interval_start(grid[N]), x[N].data(), nullptr);
  return p;
}

}  // namespace

MpcSettings load_mpc_settings(const Model& model, const std::string& task_file, const std::string& reference_file) {
  (void)reference_file;
  auto root = info_parse_file(task_file); MpcSettings s;
  s.dt = root->num_or("sqp.dt", s.dt); s.time_horizon = root->num_or("mpc.timeHorizon", s.time_horizon); s.delta_tol = root->num_or("sqp.deltaTol", s.delta_tol); s.sqp_iterations = std::max(1, (int)root->num_or("sqp.sqpIteration", 1.0)); s.cost_tol = root->num_or("sqp.costTol", s.cost_tol);
  s.g_max = root->num_or("sqp.g_max", s.g_max); s.g_min = root->num_or("sqp.g_min", s.g_min);
  s.Q = info_matrix(*root, "Q", NX, NX);
  Mat Rt = info_matrix(*root, "R", NU, NU);
  Mat init = info_matrix(*root, "initialState", NX, 1); for (int i = 0; i < NX; ++i) s.initial_state[i] = init(i, 0);
  // initializeInputCostWeight (QMInterface.cpp:274-299): leg-velocity block = J' R_task J with J = feet Jacobians wrt the 12 leg joints at initialState
  { double q[NQ], v[NQ] = {0}; for (int i = 0; i < NQ; ++i) q[i] = s.initial_state[6 + i]; RbdData d; compute_rbd(model, q, v, d, 1);
    Mat J = d.Jfoot.block(0, 6, 12, 12); Mat blk = J.T() * (Rt.block(12, 12, 12, 12) * J); s.R = Rt; s.R.set_block(12, 12, blk); }
  s.mu_ee_pos = root->num_or("endEffector.muPosition", 1.0); s.mu_ee_ori = root->num_or("endEffector.muOrientation", 1.0);
  s.mu_final_ee_pos = root->num_or("finalEndEffector.muPosition", 1.0); s.mu_final_ee_ori = root->num_or("finalEndEffector.muOrientation", 1.0);
  s.friction_mu = root->num_or("frictionConeSoftConstraint.frictionCoefficient", 1.0); s.friction_barrier_mu = root->num_or("frictionConeSoftConstraint.mu", 0.1); s.friction_barrier_delta = root->num_or("frictionConeSoftConstraint.delta", 5.0);
  s.pos_limit_mu = root->num_or("jointPositionLimits.mu", 1e-2); s.pos_limit_delta = root->num_or("jointPositionLimits.delta", 1e-3);
  s.vel_limit_mu = root->num_or("jointVelocityLimits.mu", 1e-2); s.vel_limit_delta = root->num_or("jointVelocityLimits.delta", 1e-3);
  Mat lo = info_matrix(*root, "jointVelocityLimits.lowerBound.arm", 6, 1), hi = info_matrix(*root, "jointVelocityLimits.upperBound.arm", 6, 1);
  for (int i = 0; i < 6; ++i) { s.arm_pos_lower[i] = model.joint[12 + i].lower; s.arm_pos_upper[i] = model.joint[12 + i].upper; s.arm_vel_lower[i] = lo(i, 0); s.arm_vel_upper[i] = hi(i, 0); }
  s.lift_off_velocity = root->num_or("swing_trajectory_config.liftOffVelocity", s.lift_off_velocity); s.touch_down_velocity = root->num_or("swing_trajectory_config.touchDownVelocity", s.touch_down_velocity);
  s.swing_height = root->num_or("swing_trajectory_config.swingHeight", s.swing_height); s.swing_time_scale = root->num_or("swing_trajectory_config.swingTimeScale", s.swing_time_scale);
  s.position_error_gain = root->num_or("model_settings.positionErrorGain", 0.0);
  return s;
}

void centroidal_state_from_rbd(const Model& m, const double* rbd, double* x) {
  double q[NQ], v[NQ];
  for (int i = 0; i < 3; ++i) { q[i] = rbd[3 + i]; q[3 + i] = rbd[i]; v[i] = rbd[NQ + 3 + i]; }
  M3<double> Tm = euler_rate_map<double>(q[3], q[4]); V3<double> ed = inverse3(Tm) * V3<double>(rbd[NQ], rbd[NQ + 1], rbd[NQ + 2]); for (int i = 0; i < 3; ++i) v[3 + i] = ed[i];
  for (int j = 0; j < NJ; ++j) { q[6 + j] = rbd[6 + j]; v[6 + j] = rbd[NQ + 6 + j]; }
  SrbdBase<double> s = srbd_base<double>(m, q); V3<double> vl(v[0], v[1], v[2]), ve(v[3], v[4], v[5]);
  V3<double> hl = m.mass * vl + s.A12 * ve, ha = s.A22 * ve;
  for (int i = 0; i < 3; ++i) { x[i] = hl[i] / m.mass; x[3 + i] = ha[i] / m.mass; }
  for (int i = 0; i < NQ; ++i) x[6 + i] = q[i];
}

int mode_at_time(const ModeSchedule& s, double t) { return s.mode_sequence[find_index(s.event_times, t)]; }

std::vector<NodeInfo> time_discretization_with_events(double t0, double tf, double dt, const std::vector<double>& ev) {
  const double dt_min = 10.0 * 1e-9;   // 10 * ocs2 numeric_traits::limitEpsilon (1e-9) [upstream, recalled]
  std::vector<NodeInfo> g; g.push_back({t0, 0}); size_t next_ev = (size_t)find_index(ev, t0);
  while (g.back().t < tf) {
    NodeInfo nn{g.back().t + dt, 0}; bool is_event = false;
    if (next_ev < ev.size() && nn.t >= ev[next_ev]) { nn.t = ev[next_ev]; is_event = true; nn.event = 1; ++next_ev; }
    if (nn.t >= tf) { is_event = false; nn.t = tf; nn.event = 0; }
    if (nn.t > g.back().t + dt_min) g.push_back(nn); else if (g.back().event != 2) g.back() = nn; else if (nn.t >= tf) break;
    if (is_event) { nn.event = 2; g.push_back(nn); }
  }
  return g;
}

// SwingTrajectoryPlanner (ocs2_legged_robot) [upstream, recalled]: two cubic segments per swing phase, terrain height 0
void swing_reference(const MpcSettings& s, const ModeSchedule& sched, int leg, double t, double* z_pos, double* z_vel) {
  const int np = (int)sched.mode_sequence.size(); const int p = find_index(sched.event_times, t);
  auto contact = [&](int ph) { bool f[4]; mode_to_contact_flags(sched.mode_sequence[ph], f); return f[leg]; };
  *z_pos = 0.0; *z_vel = 0.0; if (contact(p)) return;
  int start = -1; for (int ip = p - 1; ip >= 0; --ip) if (contact(ip)) { start = ip; break; }
  int fin = np - 1; for (int ip = p + 1; ip < np; ++ip) if (contact(ip)) { fin = ip - 1; break; }
  if (start < 0 || fin >= np - 1) throw std::runtime_error("swing_reference: swing phase not enclosed by stance phases in the mode schedule");
  const double t0 = sched.event_times[start], t1 = sched.event_times[fin]; const double scaling = std::min(1.0, (t1 - t0) / s.swing_time_scale);
  const double tm = 0.5 * (t0 + t1), zm = scaling * s.swing_height;
  auto cubic = [&](double ta, double pa, double va, double tb, double pb, double vb) {
    const double dtt = tb - ta, dp = pb - pa, dv = vb - va; const double c0 = pa, c1 = va * dtt, c2 = -(3.0 * va + dv) * dtt + 3.0 * dp, c3 = (2.0 * va + dv) * dtt - 2.0 * dp;
    const double tn = (t - ta) / dtt; *z_pos = c3 * tn * tn * tn + c2 * tn * tn + c1 * tn + c0; *z_vel = (3.0 * c3 * tn * tn + 2.0 * c2 * tn + c1) / dtt; };
  if (t < tm) cubic(t0, 0.0, scaling * s.lift_off_velocity, tm, zm, 0.0); else cubic(tm, zm, 0.0, t1, 0.0, scaling * s.touch_down_velocity);
}

void flow_map_jacobians(const Model& m, const double* x, const double* u, double* f, double* A, double* B) {
  D60 z[60]; for (int i = 0; i < 60; ++i) z[i] = D60::variable(i < 30 ? x[i] : u[i - 30], i); D60 fd[NX]; flow_map<D60>(m, z, z + 30, fd);
  for (int i = 0; i < NX; ++i) { f[i] = fd[i].v; for (int j = 0; j < 30; ++j) { A[30 * i + j] = fd[i].d[j]; B[30 * i + j] = fd[i].d[30 + j]; } }
}

MpcSolution mpc_solve(const Model& m, const MpcSettings& s, double t0, const double* x0, const ModeSchedule& sched, const TargetTrajectories& tt, const MpcSolution* prev, MpcDebug* dbg) {
  MpcSolution sol; sol.grid = time_discretization_with_events(t0, t0 + s.time_horizon, s.dt, sched.event_times);
  const std::vector<NodeInfo>& grid = sol.grid; const int N = (int)grid.size() - 1;
  // ---- initializeStateInputTrajectories [recalled] ----
  std::vector<Vec>& x = sol.x; std::vector<Vec>& u = sol.u; x.clear(); u.clear();
  std::vector<double> ptimes; std::vector<Vec> pu;
  double state_till = grid[0].t, input_till = grid[0].t;
  if (prev && prev->grid.size() >= 2) { for (auto& n : prev->grid) ptimes.push_back(n.t); state_till = ptimes.back(); input_till = ptimes[ptimes.size() - 2];
    pu = prev->u; for (size_t i = 0; i + 1 < prev->grid.size(); ++i) if (prev->grid[i].event == 1 && i > 0) pu[i] = pu[i - 1]; pu.push_back(pu.back()); }
  { const double ti = interval_start(grid[0]); if (ti < state_till) x.push_back(interpolate(ti, ptimes, prev->x)); else x.push_back(Vec(x0, x0 + NX)); }
  for (int k = 0; k < N; ++k) {
    if (grid[k].event == 1) { u.push_back(Vec(NU, 0.0)); x.push_back(x.back()); continue; }
    const double t = interval_start(grid[k]), tn = interval_end(grid[k + 1]);
    if (t > input_till || tn > state_till) {  // QMInitializer::compute (QMInitializer.cpp:33-41)
      bool fl[4]; mode_to_contact_flags(mode_at_time(sched, t), fl); Vec ui(NU); weight_compensating_input(m, fl, ui.data()); u.push_back(ui); x.push_back(x.back());
    } else { u.push_back(interpolate(t, ptimes, pu)); x.push_back(interpolate(tn, ptimes, prev->x)); }
  }
  // ---- DDP (GaussNewtonDDP::runImpl [upstream ocs2_ddp, recalled]; ddp{} of task.info:33-71): the nominal trajectory is a ROLLOUT of the nominal inputs from the
  //      measured state (single shooting: no dynamics defect).  Restated on the solver's own time grid with its RK2 step (the reference integrates with ODE45,
  //      rollout{} task.info:128-136, and sweeps a continuous-time Riccati equation for algorithm SLQ; this is the discrete-time form, ddp.algorithm ILQR) ----
  auto rollout_nominal = [&](std::vector<Vec>& xs, const std::vector<Vec>& us) {
    xs.assign(N + 1, Vec(NX)); for (int i = 0; i < NX; ++i) xs[0][i] = x0[i];
    for (int k = 0; k < N; ++k) {
      if (grid[k].event == 1) { xs[k + 1] = xs[k]; continue; }
      const double t = interval_start(grid[k]), dt = interval_end(grid[k + 1]) - t; (void)t; double xn[NX]; rk2_step<double>(m, s, xs[k].data(), us[k].data(), dt, xn); xs[k + 1].assign(xn, xn + NX); }
  };
  if (s.solver == 2) rollout_nominal(x, u);
  // ---- SqpSolver::runImpl [upstream ocs2_sqp, recalled]: for (iter < sqpIteration) { setupQuadraticSubproblem; getOCPSolution; takeStep; checkConvergence } ----
  for (int iteration = 0; iteration < s.sqp_iterations; ++iteration) {
  if (dbg) { dbg->A.clear(); dbg->B.clear(); dbg->b.clear(); dbg->Q.clear(); dbg->R.clear(); dbg->P.clear(); dbg->C.clear(); dbg->D.clear(); dbg->q.clear(); dbg->r.clear(); dbg->e.clear(); dbg->is_event.clear(); }
  // ---- setupQuadraticSubproblem ----
  struct Stage { bool event = false; int nu = 0; Mat A, B; Vec b; Mat Q, R, P; Vec q, r; Mat Px, Pu; Vec Pe; };
  std::vector<Stage> st(N); Quad qN; Perf base;
  for (int i = 0; i < NX; ++i) { const double d = x0[i] - x[0][i]; base.dyn += d * d; }
  for (int k = 0; k < N; ++k) {
    Stage& S = st[k];
    if (grid[k].event == 1) { S.event = true; S.nu = 0; S.A = Mat::identity(NX); S.b = x[k] - x[k + 1]; base.dyn += norm2(S.b);
      if (dbg) { dbg->A.push_back(S.A); dbg->B.push_back(Mat(NX, NU)); dbg->b.push_back(S.b); dbg->Q.push_back(Mat(NX, NX)); dbg->R.push_back(Mat(NU, NU)); dbg->P.push_back(Mat(NU, NX)); dbg->C.push_back(Mat(0, NX)); dbg->D.push_back(Mat(0, NU)); dbg->q.push_back(Vec(NX, 0.0)); dbg->r.push_back(Vec(NU, 0.0)); dbg->e.push_back(Vec()); dbg->is_event.push_back(1); }
      continue; }
    const double t = interval_start(grid[k]), dt = interval_end(grid[k + 1]) - t; bool fl[4]; mode_to_contact_flags(mode_at_time(sched, t), fl);
    D60 z[60]; for (int i = 0; i < 60; ++i) z[i] = D60::variable(i < 30 ? x[k][i] : u[k][i - 30], i);
    D60 xn[NX]; rk2_step<D60>(m, s, z, z + 30, dt, xn);
    Mat A(NX, NX), B(NX, NU); Vec b(NX); for (int i = 0; i < NX; ++i) { b[i] = xn[i].v - x[k + 1][i]; for (int j = 0; j < 30; ++j) { A(i, j) = xn[i].d[j]; B(i, j) = xn[i].d[30 + j]; } }
    base.dyn += dt * norm2(b);
    Quad c; stage_cost(m, s, tt, t, fl, x[k].data(), u[k].data(), &c); base.cost += dt * c.f;
    for (auto& v : c.q) v *= dt; for (auto& v : c.r) v *= dt; for (auto& v : c.Q.a) v *= dt; for (auto& v : c.R.a) v *= dt; for (auto& v : c.P.a) v *= dt;
    D60 g[16]; const int ng = equality_constraints<D60>(m, s, sched, t, fl, z, z + 30, g);
    Mat C(ng, NX), D(ng, NU); Vec e(ng); for (int i = 0; i < ng; ++i) { e[i] = g[i].v; for (int j = 0; j < 30; ++j) { C(i, j) = g[i].d[j]; D(i, j) = g[i].d[30 + j]; } }
    base.eq += dt * norm2(e);
    if (dbg) { dbg->A.push_back(A); dbg->B.push_back(B); dbg->b.push_back(b); dbg->Q.push_back(c.Q); dbg->R.push_back(c.R); dbg->P.push_back(c.P); dbg->C.push_back(C); dbg->D.push_back(D); dbg->q.push_back(c.q); dbg->r.push_back(c.r); dbg->e.push_back(e); dbg->is_event.push_back(0); }
    // luConstraintProjection + changeOfInputVariables [upstream ocs2_core/misc/LinearAlgebra, recalled]
    FullPivLU lu(D); S.Pu = lu.kernel(); S.Px = -1.0 * lu.solve(C); S.Pe = colvec(-1.0 * lu.solve(col(e))); S.nu = S.Pu.c;
    S.A = A + B * S.Px; S.b = b + B * S.Pe; S.B = B * S.Pu;
    Vec rs = c.r + c.R * S.Pe;                 // r shifted by u0 = Pe
    Vec qs = c.q + tmul(c.P, S.Pe);
    Mat RPx = c.R * S.Px;
    S.q = qs + tmul(S.Px, rs); S.Q = c.Q + S.Px.T() * c.P + c.P.T() * S.Px + S.Px.T() * RPx;
    S.P = S.Pu.T() * (c.P + RPx); S.r = tmul(S.Pu, rs); S.R = S.Pu.T() * (c.R * S.Pu);
  }
  base.cost += final_cost(m, s, tt, interval_start(grid[N]), x[N].data(), &qN);
  if (dbg) { dbg->QN = qN.Q; dbg->qN = qN.q; }
  // ---- QP: Riccati backward / forward (HPIPM with no inequality rows ≡ exact LQ solve) ----
  std::vector<Mat> K(N); std::vector<Vec> kff(N);
  Mat P = qN.Q; Vec p = qN.q;
  for (int k = N - 1; k >= 0; --k) {
    Stage& S = st[k]; Vec Pb = P * S.b; Vec pPb = p + Pb;
    if (S.event) { p = pPb; continue; }
    Mat PA = P * S.A, PB = P * S.B; Mat H = S.R + S.B.T() * PB; Mat G = S.P + S.B.T() * PA; Vec h = S.r + tmul(S.B, pPb);
    for (int i = 0; i < H.r; ++i) for (int j = 0; j < i; ++j) { double a = 0.5 * (H(i, j) + H(j, i)); H(i, j) = H(j, i) = a; }
    Mat L; if (!cholesky(H, L)) throw std::runtime_error("mpc_solve: projected input Hessian not positive definite");
    K[k] = -1.0 * chol_solve(L, G); kff[k] = colvec(-1.0 * chol_solve(L, col(h)));
    Mat Pn = S.Q + S.A.T() * PA + G.T() * K[k]; Vec pn = S.q + tmul(S.A, pPb) + tmul(G, kff[k]);
    for (int i = 0; i < NX; ++i) for (int j = 0; j < i; ++j) { double a = 0.5 * (Pn(i, j) + Pn(j, i)); Pn(i, j) = Pn(j, i) = a; }
    P = Pn; p = pn;
  }
  std::vector<Vec> dx(N + 1), du(N); dx[0] = Vec(NX); for (int i = 0; i < NX; ++i) dx[0][i] = x0[i] - x[0][i];
  double armijo = 0;
  for (int k = 0; k < N; ++k) {
    Stage& S = st[k];
    if (S.event) { dx[k + 1] = dx[k] + S.b; du[k] = Vec(NU, 0.0); continue; }
    Vec dut = K[k] * dx[k] + kff[k]; dx[k + 1] = S.A * dx[k] + S.B * dut + S.b; armijo += dot(S.q, dx[k]) + dot(S.r, dut);
    du[k] = S.Px * dx[k] + S.Pu * dut + S.Pe;
  }
  armijo += dot(qN.q, dx[N]);
  if (s.solver == 2) {
    // ---- DDP line search [upstream ocs2_ddp LineSearchStrategy, recalled]: step lengths maxStepLength * contraction^j >= minStepLength; the candidate is a ROLLOUT
    //      of the updated affine controller  u = u_nom + alpha * du_ff + K (x - x_nom)  (full-space law of the constrained LQ problem: K = Px + Pu K~,
    //      du_ff = Pu k~ + Pe); merit = cost + penalty * sqrt(equality-constraint SSE), accepted at the first sufficient decrease ----
    const double merit0 = base.cost + s.ddp_penalty * std::sqrt(base.eq);
    double alpha = s.ddp_max_step; bool accepted = false; Perf sp; int trials = 0; std::vector<Vec> xn(N + 1), un(N); double dxn = 0, dun = 0;
    for (auto& v : dx) dxn += norm2(v); for (auto& v : du) dun += norm2(v); dxn = std::sqrt(dxn); dun = std::sqrt(dun);
    while (alpha >= s.ddp_min_step) {
      ++trials; xn[0].assign(x0, x0 + NX);
      for (int k = 0; k < N; ++k) {
        Stage& S = st[k];
        if (S.event) { un[k] = Vec(NU, 0.0); xn[k + 1] = xn[k]; continue; }
        Vec ddx = xn[k] - x[k]; Vec dut = K[k] * ddx + alpha * kff[k]; Vec d = S.Px * ddx + S.Pu * dut + alpha * S.Pe; un[k] = u[k] + d;
        const double t = interval_start(grid[k]), dt = interval_end(grid[k + 1]) - t; (void)t; double xx[NX]; rk2_step<double>(m, s, xn[k].data(), un[k].data(), dt, xx); xn[k + 1].assign(xx, xx + NX);
      }
      sp = compute_performance(m, s, sched, tt, grid, x0, xn, un);
      const double merit = sp.cost + s.ddp_penalty * std::sqrt(sp.eq);
      if (merit < merit0 - s.ddp_armijo * alpha * std::fabs(merit0)) { accepted = true; break; }
      alpha *= s.ddp_contraction;
    }
    if (accepted) { x = xn; u = un; } else alpha = 0.0;
    if (dbg) { dbg->alpha = alpha; dbg->base_cost = base.cost; dbg->base_dyn_sse = base.dyn; dbg->base_eq_sse = base.eq; dbg->step_cost = sp.cost; dbg->step_dyn_sse = sp.dyn; dbg->step_eq_sse = sp.eq; dbg->armijo = armijo; dbg->trials = trials; dbg->dx = dx; dbg->du = du; dbg->iterations = iteration + 1; dbg->convergence = 0; }
    if (iteration + 1 >= s.sqp_iterations || !accepted) break;
    continue;
  }
  // ---- takeStep: filter line search (ocs2 FilterLinesearch) [recalled] ----
  auto violation = [](const Perf& p) { return std::sqrt(p.dyn + p.eq); };
  double dxn = 0, dun = 0; for (auto& v : dx) dxn += norm2(v); for (auto& v : du) dun += norm2(v); dxn = std::sqrt(dxn); dun = std::sqrt(dun);
  double alpha = 1.0; bool accepted = false; Perf stepPerf; int trials = 0;
  std::vector<Vec> xn(N + 1), un(N);
  do {
    ++trials;
    for (int k = 0; k <= N; ++k) xn[k] = x[k] + alpha * dx[k]; for (int k = 0; k < N; ++k) un[k] = u[k] + alpha * du[k];
    stepPerf = compute_performance(m, s, sched, tt, grid, x0, xn, un);
    const double bv = violation(base), sv = violation(stepPerf), am = alpha * armijo;
    if (sv > s.g_max) accepted = sv < (1.0 - s.gamma_c) * bv;
    else if (sv < s.g_min && bv < s.g_min && am < 0.0) accepted = stepPerf.cost < base.cost + s.armijo_factor * am;
    else accepted = stepPerf.cost < (base.cost - s.gamma_c * bv) || sv < (1.0 - s.gamma_c) * bv;
    if (accepted) break;
    alpha *= s.alpha_decay;
    if (alpha * dxn < s.delta_tol && alpha * dun < s.delta_tol) break;
  } while (alpha >= s.alpha_min);
  if (accepted) { x = xn; u = un; } else alpha = 0.0;
  if (dbg) { dbg->alpha = alpha; dbg->base_cost = base.cost; dbg->base_dyn_sse = base.dyn; dbg->base_eq_sse = base.eq; dbg->step_cost = stepPerf.cost; dbg->step_dyn_sse = stepPerf.dyn; dbg->step_eq_sse = stepPerf.eq; dbg->armijo = armijo; dbg->trials = trials; dbg->dx = dx; dbg->du = du; }
  // checkConvergence [upstream ocs2_sqp SqpSolver.cpp, recalled]
  int conv = -1;
  if (iteration + 1 >= s.sqp_iterations) conv = 0;
  else if (alpha < s.alpha_min) conv = 1;                                                                                   // a rejected step has stepSize 0
  else if (std::fabs((accepted ? stepPerf.cost : base.cost) - base.cost) < s.cost_tol && violation(accepted ? stepPerf : base) < s.g_min) conv = 2;
  else if (alpha * dxn < s.delta_tol && alpha * dun < s.delta_tol) conv = 3;
  if (dbg) { dbg->iterations = iteration + 1; dbg->convergence = conv < 0 ? 0 : conv; }
  if (conv >= 0) break;
  }
  return sol;
}

double stage_probe(const Model& m, const MpcSettings& s, const ModeSchedule& sched, const TargetTrajectories& tt, double t, const double* x, const double* u, double* q, double* r, double* g, int* ng) {
  bool fl[4]; mode_to_contact_flags(mode_at_time(sched, t), fl); Quad c; const double f = stage_cost(m, s, tt, t, fl, x, u, (q || r) ? &c : nullptr);
  if (q) for (int i = 0; i < NX; ++i) q[i] = c.q[i]; if (r) for (int i = 0; i < NU; ++i) r[i] = c.r[i];
  if (g && ng) *ng = equality_constraints<double>(m, s, sched, t, fl, x, u, g);
  return f;
}

void evaluate_policy(const MpcSolution& sol, const ModeSchedule& sched, double t, double* xd, double* ud, int* mode) {
  std::vector<double> times; for (auto& n : sol.grid) times.push_back(n.t);
  std::vector<Vec> uu = sol.u; for (size_t i = 0; i + 1 < sol.grid.size(); ++i) if (sol.grid[i].event == 1 && i > 0) uu[i] = uu[i - 1]; uu.push_back(uu.back());
  Vec xs = interpolate(t, times, sol.x), us = interpolate(t, times, uu);
  for (int i = 0; i < NX; ++i) xd[i] = xs[i]; for (int i = 0; i < NU; ++i) ud[i] = us[i]; *mode = mode_at_time(sched, t);
}

}  // namespace orc
