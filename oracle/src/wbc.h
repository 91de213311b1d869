// ORACLE — TEST INFRASTRUCTURE ONLY.
// PARITY UNPINNED: the reference ships no tests or golden vectors for this path and its stack (OCS2 / Pinocchio / qpOASES / HPIPM) cannot be
// built here, so this restatement is not checked against reference outputs; DESIGN.md section 5 lists the pins used instead
// (known answers from the reference's own config, an independent numpy/scipy twin, finite-difference identities, tests/golden).
// Restatement of qm_wbc: WbcBase (qm_wbc/src/WbcBase.cpp:118-563), Task (include/qm_wbc/Task.h:17-66),
// HoQp (src/HoQp.cpp:12-159), HierarchicalWbc (src/HierarchicalWbc.cpp:18-44) and
// HierarchicalMpcWbc (src/HierarchicalMpcWbc.cpp:18-34).
#pragma once
#include "model.h"
#include "qp.h"

namespace orc {

// defaults of qm_wbc/cfg/wbcWigeht.cfg:7-47 (the only source of these gains in the reference)
struct WbcGains {
  double kp_swing = 350, kd_swing = 37, base_height_kp = 400, base_height_kd = 140, base_linear_kp = 400, base_linear_kd = 100,
         base_angular_kp = 400, base_angular_kd = 140;
  double arm_joint_kp[6] = {4000, 4200, 4000, 4000, 4200, 6000}, arm_joint_kd[6] = {75, 75, 75, 75, 75, 75};
  double ee_linear_kp[3] = {3000, 3000, 3000}, ee_linear_kd[3] = {75, 75, 75}, ee_angular_kp[3] = {2000, 2000, 2000}, ee_angular_kd[3] = {75, 75, 75};
  double friction_coeff = 0.3;   // task.info frictionConeTask.frictionCoefficient (WbcBase.cpp:590)
};
WbcGains load_wbc_gains(const std::string& gains_info_file, const std::string& task_file);

struct Task { Mat a; Vec b; Mat d; Vec f; };
Task operator+(const Task& l, const Task& r);
Task operator*(const Task& t, double s);

struct WbcDebug {  // intermediate values for the tests
  Vec q_meas, v_meas, q_des, v_des, base_acc_des; int hoqp_iterations[3] = {0, 0, 0}; int qp_status = 0;
  std::vector<Vec> level_solutions;
};

enum WbcVariant { WBC_HIERARCHICAL = 0, WBC_HIERARCHICAL_MPC = 1 };

// One WbcBase::update + HierarchicalWbc::update call. input_last is WbcBase::inputLast_ (in/out).
// Returns [x*(36); tau(18)] (WbcBase::updateCmd, WbcBase.cpp:548-563).
Vec wbc_update(const Model& model, const WbcGains& gains, const double* state_desired, const double* input_desired,
               const double* rbd_state_measured, int mode, double period, double time, double* input_last, int variant,
               WbcDebug* dbg = nullptr);

// helpers shared with the MPC oracle
void mode_to_contact_flags(int mode, bool flags[4]);

}  // namespace orc
