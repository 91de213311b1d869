// Host-side C++ use of the path exactly as a qm_control maintainer would write it (include/qmb200.hpp mirrors QMInterface, HierarchicalWbc,
// SqpMpc and the numerical body of QMController).  Also the fixture of tests/test_cpp_mirror_*.py:
//   plugin_demo <assets_dir>              error conventions only (runs without a GPU): prints "no-gpu" or "gpu"
//   plugin_demo <assets_dir> <input.txt>  input = x_des[30] u_des[30] rbd[55] mode period time t_start; prints the WBC 54-vector of
//                                          HierarchicalWbc::update, then starting → advanceMpc (stance schedule, constant target) → update
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>

#include "qmb200.hpp"

static void print_vec(const char* tag, const qm::vector_t& v) { std::printf("%s", tag); for (double x : v) std::printf(" %.17g", x); std::printf("\n"); }

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: plugin_demo <assets_dir> [input.txt]\n"); return 2; }
  const std::string dir = argv[1];
  // QMInterface.cpp:45,53,61: a missing file is an std::invalid_argument
  try { qm::QMInterface bad(dir + "/does_not_exist.info", dir + "/qm_robot.urdf", dir + "/qm_reference.info"); std::printf("FAIL: missing task file accepted\n"); return 1; }
  catch (const std::invalid_argument& e) { std::printf("invalid_argument: %s\n", e.what()); }
  qm::QMInterface interface(dir + "/qm_task.info", dir + "/qm_robot.urdf", dir + "/qm_reference.info");
  std::shared_ptr<qm::Solver> solver;
  try { solver = std::make_shared<qm::Solver>(interface, 1, 0, QMB200_WBC_HIERARCHICAL, 1.0, 0.015); }
  catch (const std::runtime_error& e) {
    const std::string msg = e.what();
    if (msg.find("no CPU fallback") != std::string::npos) { std::printf("no-gpu: %s\n", msg.c_str()); return 0; }   // the product path fails loudly without a CUDA device
    std::printf("FAIL: %s\n", msg.c_str()); return 1;
  }
  std::printf("gpu\n");
  if (argc < 3) return 0;
  std::ifstream in(argv[2]); qm::vector_t x_des(30), u_des(30), rbd(55); double mode_d, period, time, t_start;
  for (double& v : x_des) in >> v;
  for (double& v : u_des) in >> v;
  for (double& v : rbd) in >> v;
  in >> mode_d >> period >> time >> t_start;
  if (!in) { std::printf("FAIL: bad input file\n"); return 1; }
  // WBC seam (QMController.cpp:146)
  qm::HierarchicalWbc wbc(solver);
  print_vec("wbc", wbc.update(x_des, u_des, rbd, static_cast<size_t>(mode_d), period, time)); std::printf("wbc_status %d\n", wbc.lastStatus());
  // controller body: starting → advanceMpc → update
  qm::QMController ctrl(interface, 0, false);
  ctrl.starting(rbd, t_start);
  qm::ModeSchedule sched; sched.eventTimes = {t_start - 1.0, t_start + 5.0}; sched.modeSequence = {15, 15, 15};
  qm::TargetTrajectories tt; tt.timeTrajectory = {t_start, t_start + 1.0}; qm::vector_t target(37, 0.0);
  for (int i = 6; i < 30; ++i) target[i] = ctrl.observationState()[i];
  for (int i = 0; i < 7; ++i) target[30 + i] = rbd[48 + i];
  tt.stateTrajectory = {target, target};
  const qm::PrimalSolution sol = ctrl.advanceMpc(sched, tt);
  std::printf("mpc_nodes %zu status %d step %.17g\n", sol.timeTrajectory.size(), sol.status, sol.stepSize);
  print_vec("mpc_x1", sol.stateTrajectory[1]);
  qm::vector_t out; const bool safe = ctrl.update(rbd, period, out);
  print_vec("update", out); std::printf("update_status %d safe %d time %.17g\n", ctrl.lastStatus(), safe ? 1 : 0, ctrl.observationTime());
  qm::vector_t jc;
  for (int j = 0; j < 18; ++j) { auto c = ctrl.jointCommand(j); jc.insert(jc.end(), {c.posDes, c.velDes, c.kp, c.kd, c.ff}); }
  print_vec("joint_cmd", jc);
  return 0;
}
