// ORACLE — TEST INFRASTRUCTURE ONLY.
// PARITY UNPINNED: the reference ships no tests or golden vectors for this path and its stack (OCS2 / Pinocchio / qpOASES / HPIPM) cannot be
// built here, so this restatement is not checked against reference outputs; DESIGN.md section 5 lists the pins used instead
// (known answers from the reference's own config, an independent numpy/scipy twin, finite-difference identities, tests/golden).
// Kinematic/dynamic model of the quadruped-manipulator as the reference builds it:
//   qm_interface/src/QMInterface.cpp:408-439 (setupModel → centroidal_model::createPinocchioInterface,
//   createCentroidalModelInfo) [upstream OCS2 FactoryFunctions.cpp, recalled]:
//   root joint = composite(Translation, SphericalZYX); q = [p_base, zyx euler, joints], v = qdot;
//   joints traversed depth-first with children ordered by joint name (urdfdom name-keyed maps)
//   ⇒ LF,LH,RF,RH,arm (matches qm_controllers/config/task.info:168-188); fixed joints lumped.
#pragma once
#include <string>
#include <vector>

#include "dual.h"
#include "info.h"
#include "linalg.h"

namespace orc {

constexpr int NQ = 24, NJ = 18, NB = 19, NX = 30, NU = 30;

struct JointDef { std::string name; int parent_body; M3<double> R; V3<double> p; int axis; double lower, upper, effort, velocity; };
struct BodyDef { double mass = 0; V3<double> com; M3<double> I; };  // I about com, body frame
struct FrameDef { std::string name; int body; M3<double> R; V3<double> p; };

struct Model {
  JointDef joint[NJ];       // joint j moves body j+1
  BodyDef body[NB];         // body 0 = base
  std::vector<FrameDef> frames;
  int foot_frame[4];        // contact order LF_FOOT, RF_FOOT, LH_FOOT, RH_FOOT (ModelSettings.h:38)
  int ee_frame;             // model_settings.eeFrame (task.info:20)
  int base_frame;           // frame "base" (WbcBase.cpp:182)
  // CentroidalModelInfo (SRBD)
  double mass = 0;
  double q_nominal[NQ];
  M3<double> I_nominal;     // centroidalInertiaNominal
  V3<double> com_to_base_nominal;
  int frame_id(const std::string& n) const { for (size_t i = 0; i < frames.size(); ++i) if (frames[i].name == n) return (int)i; return -1; }
};

Model load_model(const std::string& urdf_file, const std::vector<double>& default_joint_state, const std::string& ee_frame);

template <class T> inline M3<T> rot_axis(int axis, const T& q) {
  M3<T> R = M3<T>::identity(); T c = cos(q), s = sin(q);
  if (axis == 0) { R(1, 1) = c; R(1, 2) = -s; R(2, 1) = s; R(2, 2) = c; }
  else if (axis == 1) { R(0, 0) = c; R(0, 2) = s; R(2, 0) = -s; R(2, 2) = c; }
  else { R(0, 0) = c; R(0, 1) = -s; R(1, 0) = s; R(1, 1) = c; }
  return R;
}
// ocs2 getRotationMatrixFromZyxEulerAngles: R = Rz(e0) Ry(e1) Rx(e2)
template <class T> inline M3<T> rot_zyx(const T& z, const T& y, const T& x) { return rot_axis<T>(2, z) * rot_axis<T>(1, y) * rot_axis<T>(0, x); }
// ocs2 getMappingFromEulerAnglesZyxDerivativeToGlobalAngularVelocity
template <class T> inline M3<T> euler_rate_map(const T& z, const T& y) {
  M3<T> M; T sz = sin(z), cz = cos(z), sy = sin(y), cy = cos(y);
  M(0, 0) = T(0.0); M(0, 1) = -sz; M(0, 2) = cy * cz;
  M(1, 0) = T(0.0); M(1, 1) = cz;  M(1, 2) = cy * sz;
  M(2, 0) = T(1.0); M(2, 1) = T(0.0); M(2, 2) = -sy;
  return M;
}

template <class T> struct Kin { M3<T> R[NB]; V3<T> p[NB]; };

template <class T> void forward_kinematics(const Model& m, const T* q, Kin<T>& k) {
  k.p[0] = V3<T>(q[0], q[1], q[2]);
  k.R[0] = rot_zyx<T>(q[3], q[4], q[5]);
  for (int j = 0; j < NJ; ++j) {
    const JointDef& jd = m.joint[j]; const int pb = jd.parent_body;
    k.p[j + 1] = k.p[pb] + k.R[pb] * cast3<T>(jd.p);
    k.R[j + 1] = k.R[pb] * cast3<T>(jd.R) * rot_axis<T>(jd.axis, q[6 + j]);
  }
}
template <class T> V3<T> frame_pos(const Model& m, const Kin<T>& k, int f) { const FrameDef& fd = m.frames[f]; return k.p[fd.body] + k.R[fd.body] * cast3<T>(fd.p); }
template <class T> M3<T> frame_rot(const Model& m, const Kin<T>& k, int f) { const FrameDef& fd = m.frames[f]; return k.R[fd.body] * cast3<T>(fd.R); }
template <class T> V3<T> body_com(const Model& m, const Kin<T>& k, int b) { return k.p[b] + k.R[b] * cast3<T>(m.body[b].com); }

// Whole-body quantities at (q, v) in the coordinates above — what Pinocchio returns to
// WbcBase::updateMeasured (qm_wbc/src/WbcBase.cpp:150-190).
struct RbdData {
  Mat M;        // 24x24 (crba, symmetrised WbcBase.cpp:155)
  Vec nle;      // 24 (nonLinearEffects)
  Mat Jfoot;    // 12x24 LOCAL_WORLD_ALIGNED linear rows, contact order
  Mat dJfoot;   // 12x24
  Mat Jbase, dJbase;  // 6x24 frame "base"
  Mat Jee, dJee;      // 6x24 arm end-effector
  V3<double> foot_pos[4], foot_vel[4], ee_pos, ee_vel, ee_angvel;
  M3<double> ee_rot;
  // centroidal (full model): Ag 6x24 about the COM, dAg*v, com
  Mat Ag; Vec dAg_v; V3<double> com;
};
// what: bit0 = M/nle/J/dJ (measured side), bit1 = centroidal Ag/dAg (desired side)
void compute_rbd(const Model& m, const double* q, const double* v, RbdData& out, int what = 3);

// Frame Jacobian 6x24 (rows: linear 0:3, angular 3:6) and its time derivative along v.
void frame_jacobians(const Model& m, const double* q, const double* v, int frame, Mat& J, Mat& dJ);

double total_mass(const Model& m);

}  // namespace orc
