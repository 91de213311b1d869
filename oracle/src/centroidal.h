// ORACLE — TEST INFRASTRUCTURE ONLY.
// PARITY UNPINNED: the reference ships no tests or golden vectors for this path and its stack (OCS2 / Pinocchio / qpOASES / HPIPM) cannot be
// built here, so this restatement is not checked against reference outputs; DESIGN.md section 5 lists the pins used instead
// (known answers from the reference's own config, an independent numpy/scipy twin, finite-difference identities, tests/golden).
// Single-rigid-body centroidal model helpers [upstream ocs2_centroidal_model, recalled — SURVEY.md App. A.2]:
// updateCentroidalDynamics (SRBD branch), computeFloatingBaseCentroidalMomentumMatrixInverse,
// CentroidalModelPinocchioMapping::getPinocchioJointVelocity, getNormalizedCentroidalMomentumRate and the
// flow map behind qm_interface/src/dynamics/QMDynamicsAD.cpp:22-33.  Templated so that dual numbers
// reproduce the CppAD Jacobians.
#pragma once
#include "model.h"

namespace orc {

template <class T> inline M3<T> inverse3t(const M3<T>& A) {
  M3<T> B; const auto& m = A.m;
  T det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) + m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
  T id = T(1.0) / det;
  B.m[0][0] = (m[1][1] * m[2][2] - m[1][2] * m[2][1]) * id; B.m[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * id; B.m[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * id;
  B.m[1][0] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]) * id; B.m[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * id; B.m[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id;
  B.m[2][0] = (m[1][0] * m[2][1] - m[1][1] * m[2][0]) * id; B.m[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * id; B.m[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id;
  return B;
}

// SRBD blocks of A_b: A_b = [[m I, A12],[0, A22]],  A12 = m S(R c_nom) T,  A22 = R I_nom R^T T; also com = p - R c_nom.
template <class T> struct SrbdBase { M3<T> A12, A22, A22inv; V3<T> com; };
template <class T> inline SrbdBase<T> srbd_base(const Model& m, const T* q) {
  SrbdBase<T> s; M3<T> R = rot_zyx<T>(q[3], q[4], q[5]); M3<T> Tm = euler_rate_map<T>(q[3], q[4]);
  V3<T> c = R * cast3<T>(m.com_to_base_nominal);
  M3<T> S = skew(c); M3<T> ST = S * Tm; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) s.A12(i, j) = T(m.mass) * ST(i, j);
  s.A22 = (R * cast3<T>(m.I_nominal)) * (transpose(R) * Tm); s.A22inv = inverse3t(s.A22);
  s.com = V3<T>(q[0], q[1], q[2]) - c; return s;
}

// v_pinocchio = [A_b^{-1} (m h_normalized); u[12:30]]  (SRBD: no joint term)
template <class T> inline void pinocchio_joint_velocity(const Model& m, const T* x, const T* u, T* v) {
  SrbdBase<T> s = srbd_base<T>(m, x + 6);
  V3<T> hl(T(m.mass) * x[0], T(m.mass) * x[1], T(m.mass) * x[2]), ha(T(m.mass) * x[3], T(m.mass) * x[4], T(m.mass) * x[5]);
  V3<T> wd = s.A22inv * ha;                                  // euler rates
  V3<T> vl = T(1.0 / m.mass) * (hl - s.A12 * wd);            // base linear velocity
  for (int i = 0; i < 3; ++i) { v[i] = vl[i]; v[3 + i] = wd[i]; }
  for (int j = 0; j < NJ; ++j) v[6 + j] = u[12 + j];
}

// flow map xdot = [hdot_normalized(6); v_pinocchio(24)]
template <class T> inline void flow_map(const Model& m, const T* x, const T* u, T* f) {
  Kin<T> k; forward_kinematics<T>(m, x + 6, k);
  SrbdBase<T> s = srbd_base<T>(m, x + 6);
  V3<T> lin(T(0.0), T(0.0), T(-9.81 * m.mass)), ang;
  for (int i = 0; i < 4; ++i) { V3<T> F(u[3 * i], u[3 * i + 1], u[3 * i + 2]); V3<T> r = frame_pos(m, k, m.foot_frame[i]) - s.com; lin = lin + F; ang = ang + cross(r, F); }
  for (int i = 0; i < 3; ++i) { f[i] = lin[i] / m.mass; f[3 + i] = ang[i] / m.mass; }
  pinocchio_joint_velocity<T>(m, x, u, f + 6);
}

// foot / end-effector linear velocity in world (LOCAL_WORLD_ALIGNED) as function of (x,u):
// J_frame(q) * v_pinocchio(x,u), evaluated by pushing the Jet-free product through dual arithmetic:
// p(q + eps v) derivative == J v.  Implemented with explicit Jacobian-vector product via finite composition:
template <class T> inline V3<T> frame_velocity(const Model& m, const T* x, const T* u, int frame) {
  // v = d/dt p(q(t)) with qdot = v_pinocchio: use a 1-direction dual over T
  T v[NQ]; pinocchio_joint_velocity<T>(m, x, u, v);
  typedef Dual<1, T> DT; DT q[NQ]; for (int i = 0; i < NQ; ++i) { q[i].v = x[6 + i]; q[i].d[0] = v[i]; }
  Kin<DT> k; forward_kinematics<DT>(m, q, k); V3<DT> p = frame_pos(m, k, frame);
  return V3<T>(p.x.d[0], p.y.d[0], p.z.d[0]);
}

}  // namespace orc
