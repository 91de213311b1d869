// ORACLE — TEST INFRASTRUCTURE ONLY.
// Parity of this file: the functions restated here are reference-OWNED sources (cited line by line), pinned by the independent scipy
// transliteration in tests/test_ctrl_cpu.py; the upstream helpers they call (angles, ocs2 rotations) are recalled, i.e. PARITY UNPINNED.
// CPU restatement of the controller steps either side of the MPC+WBC path, one robot per call, written against the
// reference sources line by line (they are reference-owned code, not upstream):
//   observation_update      QMController::updateStateEstimation tail           qm_controllers/src/QMController.cpp:236-243
//   target_trajectories     cmdVel / EeCmdVel / EEgoalPose → TargetTrajectories  qm_controllers/src/QmTargetTrajectoriesPublisher_node.cpp:24-208
//   control_law             SafetyChecker + updateControlLaw (both controllers) SafetyChecker.h:22-35, QMController.cpp:177-190,427-445
//   HwSim::write            QMHWSim::writeSim                                   qm_gazebo/src/QMHWSim.cpp:98-116
// Upstream pieces restated from memory: angles::shortest_angular_distance (ROS angles), ocs2 getRotationMatrixFromZyxEulerAngles,
// ocs2 quaternionDistance, Eigen quaternion → rotation matrix.
#include <cmath>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "centroidal.h"
#include "info.h"
#include "mpc.h"

namespace orc {

struct Quat { double w, x, y, z; };
static Quat qmul(const Quat& a, const Quat& b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
// rotate v by the unit quaternion q (q v q*), i.e. Eigen's toRotationMatrix() * v
static V3<double> qrot(const Quat& q, const V3<double>& v) { Quat p{0.0, v.x, v.y, v.z}, c{q.w, -q.x, -q.y, -q.z}; Quat r = qmul(qmul(q, p), c); return {r.x, r.y, r.z}; }

// angles::normalize_angle / shortest_angular_distance (ROS angles package): result in (-pi, pi]
static double shortest_angular_distance(double from, double to) {
  double a = std::remainder(to - from, 2.0 * M_PI);          // [-pi, pi]
  if (a <= -M_PI) a += 2.0 * M_PI;
  return a;
}

struct TargetConstants { double com_height, disp_vel, rot_vel, time_to_target; std::vector<double> default_joint_state; };

void observation_update(const Model& m, const double* rbd55, double period, double* t_obs, double* x_obs30) {
  const double yaw_last = x_obs30[9];                                                   // QMController.cpp:238
  double x[NX]; centroidal_state_from_rbd(m, rbd55, x);                                 // :239-240 (head(2*24): no ee state)
  x[9] = yaw_last + shortest_angular_distance(yaw_last, x[9]);                          // :241
  *t_obs += period;                                                                     // :237
  std::memcpy(x_obs30, x, sizeof(x));
}

// targetPoseToTargetTrajectories (QmTargetTrajectoriesPublisher_node.cpp:44-68)
static void pose_to_trajectories(const TargetConstants& c, const double* ee_target7, const double* base_target6, double t_obs, const double* x_obs, const double* ee_current7, double t_reach,
                                 double* times2, double* states2x37) {
  times2[0] = t_obs; times2[1] = t_reach;
  double base_cur[6]; for (int i = 0; i < 6; ++i) base_cur[i] = x_obs[6 + i];
  base_cur[2] = c.com_height; base_cur[4] = 0.0; base_cur[5] = 0.0;
  for (int k = 0; k < 2; ++k) {
    double* s = states2x37 + 37 * k; for (int i = 0; i < 6; ++i) s[i] = 0.0;
    for (int i = 0; i < 6; ++i) s[6 + i] = k == 0 ? base_cur[i] : base_target6[i];
    for (int j = 0; j < NJ; ++j) s[12 + j] = c.default_joint_state[j];
    for (int i = 0; i < 7; ++i) s[30 + i] = k == 0 ? ee_current7[i] : ee_target7[i];
  }
}

void target_trajectories(const TargetConstants& c, int kind, const double* cmd, double t_obs, const double* x_obs, const double* ee_state7, double* last_ee7, double* times2, double* states2x37) {
  const double* base = x_obs + 6;
  if (kind == 0) {           // cmdVelToTargetTrajectories (:73-113)
    M3<double> R = rot_zyx<double>(base[3], base[4], base[5]); V3<double> v = R * V3<double>(cmd[0], cmd[1], cmd[2]);
    const double T = c.time_to_target;
    double bt[6] = {base[0] + v.x * T, base[1] + v.y * T, c.com_height, base[3] + cmd[3] * T, 0.0, 0.0};
    double d2 = 0.0; for (int i = 0; i < 3; ++i) d2 += (last_ee7[i] - ee_state7[i]) * (last_ee7[i] - ee_state7[i]);
    if (std::sqrt(d2) > 0.1) for (int i = 0; i < 3; ++i) last_ee7[i] = ee_state7[i];
    double et[7]; std::memcpy(et, last_ee7, sizeof(et));
    pose_to_trajectories(c, et, bt, t_obs, x_obs, /*eeStateLast*/ et, t_obs + T, times2, states2x37);
    for (int k = 0; k < 2; ++k) { states2x37[37 * k] = v.x; states2x37[37 * k + 1] = v.y; states2x37[37 * k + 2] = v.z; }
  } else if (kind == 1) {    // EeCmdVelToTargetTrajectories (:118-165)
    const Quat qi{-0.5, 0.5, -0.5, 0.5}, q{ee_state7[6], ee_state7[3], ee_state7[4], ee_state7[5]}, qi_conj{qi.w, -qi.x, -qi.y, -qi.z};
    V3<double> v = qrot(q, qrot(qi_conj, V3<double>(cmd[0], cmd[1], cmd[2])));          // quat.R * quat_init.R^T * cmdVel
    const double T = c.time_to_target;
    double et[7] = {ee_state7[0] + v.x * T, ee_state7[1] + v.y * T, last_ee7[2], last_ee7[3], last_ee7[4], last_ee7[5], last_ee7[6]};
    double bt[6] = {et[0] - 0.52, et[1] - 0.09, c.com_height, base[3], 0.0, 0.0};
    pose_to_trajectories(c, et, bt, t_obs, x_obs, ee_state7, t_obs + T, times2, states2x37);
  } else {                   // EEgoalPoseToTargetTrajectories (:172-208); processFeedback updates lastEeTarget_ (QmTargetTrajectoriesPublisher.cpp:107-108)
    double et[7]; std::memcpy(et, cmd, sizeof(et));
    double bt[6] = {cmd[0] - 0.52, cmd[1] - 0.09, c.com_height, base[3], 0.0, 0.0};
    // quaternionDistance(q_current, q_target) [upstream ocs2_robotic_tools]: q.w qRef.vec - qRef.w q.vec + q.vec x qRef.vec
    V3<double> vc(ee_state7[3], ee_state7[4], ee_state7[5]), vt(cmd[3], cmd[4], cmd[5]); V3<double> dq = ee_state7[6] * vt - cmd[6] * vc + cross(vc, vt);
    const double displacement = std::sqrt((cmd[0] - ee_state7[0]) * (cmd[0] - ee_state7[0]) + (cmd[1] - ee_state7[1]) * (cmd[1] - ee_state7[1]) + (cmd[2] - ee_state7[2]) * (cmd[2] - ee_state7[2]));
    const double rotation = std::sqrt(dq.x * dq.x + dq.y * dq.y + dq.z * dq.z);
    const double t_reach = t_obs + std::max(rotation / c.rot_vel, displacement / c.disp_vel);     // estimateTimeToTarget (:24-41)
    pose_to_trajectories(c, et, bt, t_obs, x_obs, ee_state7, t_reach, times2, states2x37);
    std::memcpy(last_ee7, cmd, 7 * sizeof(double));
  }
}

// returns SafetyChecker::check (true = safe)
bool control_law(int variant, double arm_kp, double arm_kd, const double* x_des, const double* u_des, const double* wbc54, double time, const double* x_obs, double* joint_cmd18x5, double* arm_pos6, double* last_time) {
  const double* pos_des = x_des + 12; const double* vel_des = u_des + 12; const double* torque = wbc54 + 36;
  auto set = [&](int j, double p, double v, double kp, double kd, double ff) { double* c = joint_cmd18x5 + 5 * j; c[0] = p; c[1] = v; c[2] = kp; c[3] = kd; c[4] = ff; };
  if (variant == 0) {
    if (time > 10) for (int j = 0; j < 12; ++j) set(j, pos_des[j], vel_des[j], 0, 3, torque[j]);                         // QMController.cpp:179-185
    for (int j = 12; j < 18; ++j) set(j, pos_des[j], 0.0, arm_kp, arm_kd, torque[j]);                                    // :187-189
  } else {
    for (int j = 0; j < 12; ++j) set(j, pos_des[j], vel_des[j], 0, 3, torque[j]);                                        // :428-430
    if (time - *last_time > 1.0 / 100.0) { for (int j = 0; j < 6; ++j) arm_pos6[j] = x_obs[24 + j] + vel_des[12 + j] * 1.0 / 100.0; *last_time = time; }   // :432-444
  }
  const double roll = x_obs[6 + 5];                                                                                      // getBasePose(state)(5), SafetyChecker.h:28-29
  return !(roll > M_PI_2 || roll < -M_PI_2);
}

struct HwCommand { double stamp; double c[18][5]; };
struct HwSim {
  double delay = 0.0; std::deque<HwCommand> buffer;
  void write(double time, double period, const double* joint_cmd18x5, const double* pos, const double* vel, double* effort) {
    if (time == period) buffer.clear();                                                                                  // QMHWSim.cpp:101-103
    while (!buffer.empty() && buffer.back().stamp + delay < time) buffer.pop_back();                                     // :105-107
    HwCommand n; n.stamp = time; std::memcpy(n.c, joint_cmd18x5, sizeof(n.c)); buffer.push_front(n);                     // :108-109
    const HwCommand& cmd = buffer.back();                                                                                // :111
    for (int j = 0; j < 18; ++j) effort[j] = cmd.c[j][2] * (cmd.c[j][0] - pos[j]) + cmd.c[j][3] * (cmd.c[j][1] - vel[j]) + cmd.c[j][4];   // :112-113
  }
};

}  // namespace orc

using namespace orc;
namespace { thread_local std::string g_ctrl_err; }

extern "C" {

// handle-free constants object (reads reference.info / task.info like the publisher node's main, :225-229)
void* orc_ctrl_create(const char* task, const char* reference) {
  try {
    auto troot = info_parse_file(task); auto rroot = info_parse_file(reference);
    auto* c = new TargetConstants(); c->com_height = rroot->num("comHeight"); c->disp_vel = rroot->num("targetDisplacementVelocity"); c->rot_vel = rroot->num("targetRotationVelocity");
    c->time_to_target = troot->num("mpc.timeHorizon"); Mat d = info_matrix(*rroot, "defaultJointState", NJ, 1); for (int j = 0; j < NJ; ++j) c->default_joint_state.push_back(d(j, 0));
    return c;
  } catch (const std::exception& e) { g_ctrl_err = e.what(); return nullptr; }
}
void orc_ctrl_destroy(void* c) { delete static_cast<TargetConstants*>(c); }
void orc_target_trajectories(void* c, int kind, const double* cmd7, double t_obs, const double* x_obs, const double* ee7, double* last_ee7, double* times2, double* states2x37) {
  target_trajectories(*static_cast<TargetConstants*>(c), kind, cmd7, t_obs, x_obs, ee7, last_ee7, times2, states2x37);
}
int orc_control_law(int variant, double arm_kp, double arm_kd, const double* x_des, const double* u_des, const double* wbc54, double time, const double* x_obs, double* joint_cmd, double* arm_pos6, double* last_time) {
  return control_law(variant, arm_kp, arm_kd, x_des, u_des, wbc54, time, x_obs, joint_cmd, arm_pos6, last_time) ? 1 : 0;
}
void* orc_hw_create(double delay) { auto* h = new HwSim(); h->delay = delay; return h; }
void orc_hw_destroy(void* h) { delete static_cast<HwSim*>(h); }
void orc_hw_write(void* h, double time, double period, const double* joint_cmd, const double* pos, const double* vel, double* effort) { static_cast<HwSim*>(h)->write(time, period, joint_cmd, pos, vel, effort); }

}  // extern "C"
