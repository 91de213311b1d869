// ORACLE — TEST INFRASTRUCTURE ONLY.
// PARITY UNPINNED: the reference ships no tests or golden vectors for this path and its stack (OCS2 / Pinocchio / qpOASES / HPIPM) cannot be
// built here, so this restatement is not checked against reference outputs; DESIGN.md section 5 lists the pins used instead
// (known answers from the reference's own config, an independent numpy/scipy twin, finite-difference identities, tests/golden).
// Restatement of the MPC tick: the OCP that qm_interface defines (QMInterface.cpp:79-142) solved by ONE
// multiple-shooting SQP iteration as the controller configures it (QMController.cpp:287-288, task.info:75-92).
// All upstream OCS2 pieces are [recalled] — see SURVEY.md Appendix A and DESIGN.md for the conventions fixed here.
#pragma once
#include <vector>

#include "model.h"

namespace orc {

struct MpcSettings {
  // sqp{} / mpc{} (task.info:75-92,138-147)
  double dt = 0.015, time_horizon = 1.0, delta_tol = 1e-4, g_max = 1e-2, g_min = 1e-6;
  double alpha_decay = 0.5, alpha_min = 1e-4, gamma_c = 1e-6, armijo_factor = 1e-4;     // [upstream defaults]
  int sqp_iterations = 1; double cost_tol = 1e-4;   // sqp.sqpIteration (task.info:28), costTol [upstream ocs2_sqp default]: SqpSolver::runImpl loop + checkConvergence
  // RK2 form x+ = x + dt (w1 k1 + w2 k2), k2 = f(x + c dt k1): Heun (c=1,w=1/2,1/2) is OCS2's SensitivityIntegrator rk2 [recalled]
  double rk_c = 1.0, rk_w1 = 0.5, rk_w2 = 0.5;
  // solver variant: 0 = multiple-shooting SQP (sqp{}), 1 = multiple-shooting IPM (ipm{}: no inequality rows in this OCP => the same Newton step, other tolerances),
  // 2 = DDP (ddp{}): single-shooting rollouts, discrete Riccati backward pass, rollout line search on merit = cost + penalty * sqrt(equality SSE)
  int solver = 0; double ddp_penalty = 20.0, ddp_min_step = 1e-2, ddp_max_step = 1.0, ddp_armijo = 1e-4, ddp_contraction = 0.5;
  // cost (task.info:192-287, QMInterface.cpp:274-319)
  Mat Q, R;
  double mu_ee_pos = 2000, mu_ee_ori = 1000, mu_final_ee_pos = 2000, mu_final_ee_ori = 1000;   // task.info:235-245
  // friction cone soft constraint (task.info:290-297; FrictionConeConstraint::Config defaults [upstream])
  double friction_mu = 0.3, friction_barrier_mu = 0.1, friction_barrier_delta = 5.0, friction_reg = 25.0, friction_hess_shift = 1e-6;
  // arm joint soft box (task.info:299-343, URDF limits)
  double pos_limit_mu = 0.1, pos_limit_delta = 1e-3, vel_limit_mu = 0.1, vel_limit_delta = 1e-3;
  double arm_pos_lower[6], arm_pos_upper[6], arm_vel_lower[6], arm_vel_upper[6];
  // swing_trajectory_config (task.info:23-30), model_settings (task.info:8-21)
  double lift_off_velocity = 0.05, touch_down_velocity = -0.1, swing_height = 0.15, swing_time_scale = 0.15;
  double position_error_gain = 0.0;
  double initial_state[NX];
};
MpcSettings load_mpc_settings(const Model& model, const std::string& task_file, const std::string& reference_file);

void centroidal_state_from_rbd(const Model& m, const double* rbd48, double* x30);
struct ModeSchedule; struct TargetTrajectories;
// one evaluation of the intermediate cost and the equality constraints at (t, x, u) for finite-difference checks of their derivatives: returns the cost value,
// fills gradient q[30], r[30] (if non-null), constraint values g[<=16] and their count
double stage_probe(const Model& m, const MpcSettings& s, const ModeSchedule& sched, const TargetTrajectories& tt, double t, const double* x, const double* u, double* q, double* r, double* g, int* ng);

// ocs2::ModeSchedule: modeSequence.size() == eventTimes.size() + 1
struct ModeSchedule { std::vector<double> event_times; std::vector<int> mode_sequence; };
// ocs2::TargetTrajectories (state part only; 37 = 30 + EE pose [pos(3), quat xyzw(4)])
struct TargetTrajectories { std::vector<double> times; std::vector<Vec> states; };

struct NodeInfo { double t; int event; /*0 none, 1 pre-event, 2 post-event*/ };
struct MpcSolution { std::vector<NodeInfo> grid; std::vector<Vec> x, u; };

struct MpcDebug { int iterations = 0; int convergence = 0; /* 0 ITERATIONS, 1 STEPSIZE, 2 METRICS, 3 PRIMAL */ double alpha = 0; double base_cost = 0, base_dyn_sse = 0, base_eq_sse = 0, step_cost = 0, step_dyn_sse = 0, step_eq_sse = 0, armijo = 0; int trials = 0;
  std::vector<Mat> A, B; std::vector<Vec> b; std::vector<Vec> dx, du;
  // the QP of the (last) SQP iteration as setupQuadraticSubproblem built it, per interval (cost already scaled by dt): ½ dx'Q dx + du'P dx + ½ du'R du + q'dx + r'du,
  // C dx + D du + e = 0; event nodes carry empty matrices; QN, qN = final cost
  std::vector<Mat> Q, R, P, C, D; std::vector<Vec> q, r, e; std::vector<int> is_event; Mat QN; Vec qN; };

// One SqpSolver::run(t0, x0, t0 + horizon) with sqpIteration = 1.  `previous` may be empty (cold start → QMInitializer).
MpcSolution mpc_solve(const Model& model, const MpcSettings& s, double t0, const double* x0, const ModeSchedule& schedule, const TargetTrajectories& target,
                      const MpcSolution* previous, MpcDebug* dbg = nullptr);

// MPC_MRT_Interface::evaluatePolicy with a feed-forward policy: linear interpolation + modeAtTime.
void evaluate_policy(const MpcSolution& sol, const ModeSchedule& schedule, double t, double* x_des, double* u_des, int* mode);

// time grid (ocs2 timeDiscretizationWithEvents)
std::vector<NodeInfo> time_discretization_with_events(double t0, double tf, double dt, const std::vector<double>& event_times);
int mode_at_time(const ModeSchedule& s, double t);
// SwingTrajectoryPlanner::getZvelocityConstraint / getZpositionConstraint for foot `leg` (contact order) at time t
void swing_reference(const MpcSettings& s, const ModeSchedule& sched, int leg, double t, double* z_pos, double* z_vel);

// flow map + Jacobians (PinocchioCentroidalDynamicsAD) for the tests
void flow_map_jacobians(const Model& m, const double* x, const double* u, double* f30, double* A900, double* B900);

}  // namespace orc
