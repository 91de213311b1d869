// ORACLE — TEST INFRASTRUCTURE ONLY.
// PARITY UNPINNED: the reference ships no tests or golden vectors for this path and its stack (OCS2 / Pinocchio / qpOASES / HPIPM) cannot be
// built here, so this restatement is not checked against reference outputs; DESIGN.md section 5 lists the pins used instead
// (known answers from the reference's own config, an independent numpy/scipy twin, finite-difference identities, tests/golden).
// C entry points (ctypes) of the CPU oracle.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library; the product never does.
#include <atomic>
#include <cstring>
#include <string>
#include <thread>

#include "centroidal.h"
#include "model.h"
#include "mpc.h"
#include "wbc.h"

namespace orc { void observation_update(const Model& m, const double* rbd55, double period, double* t_obs, double* x_obs30); }
using namespace orc;

namespace {
thread_local std::string g_err;
struct Handle { Model model; WbcGains gains; MpcSettings mpc; };
template <class F> int guarded(F&& f) { try { f(); return 0; } catch (const std::exception& e) { g_err = e.what(); return -1; } }
}  // namespace

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

void* orc_create(const char* urdf, const char* task, const char* reference, const char* gains) {
  Handle* h = nullptr;
  int rc = guarded([&] {
    auto troot = info_parse_file(task); auto rroot = info_parse_file(reference);
    Mat djs = info_matrix(*rroot, "defaultJointState", NJ, 1);
    std::vector<double> d(NJ); for (int i = 0; i < NJ; ++i) d[i] = djs(i, 0);
    h = new Handle();
    h->model = load_model(urdf, d, troot->str("model_settings.eeFrame"));
    h->gains = load_wbc_gains(gains ? gains : "", task);
    h->mpc = load_mpc_settings(h->model, task, reference);
  });
  if (rc != 0) { delete h; return nullptr; }
  return h;
}
void orc_destroy(void* hp) { delete static_cast<Handle*>(hp); }

int orc_model_info(void* hp, double* mass, double* inertia_nominal9, double* com_to_base3, double* q_nominal24, double* effort18, double* lower18, double* upper18) {
  Handle* h = static_cast<Handle*>(hp);
  *mass = h->model.mass;
  for (int i = 0; i < 3; ++i) { com_to_base3[i] = h->model.com_to_base_nominal[i]; for (int j = 0; j < 3; ++j) inertia_nominal9[3 * i + j] = h->model.I_nominal(i, j); }
  for (int i = 0; i < NQ; ++i) q_nominal24[i] = h->model.q_nominal[i];
  for (int j = 0; j < NJ; ++j) { effort18[j] = h->model.joint[j].effort; lower18[j] = h->model.joint[j].lower; upper18[j] = h->model.joint[j].upper; }
  return 0;
}
int orc_joint_name(void* hp, int j, char* out, int cap) { Handle* h = static_cast<Handle*>(hp); std::strncpy(out, h->model.joint[j].name.c_str(), cap); return 0; }

// rigid-body quantities at (q, v): M[24x24], nle[24], Jfoot[12x24], dJfoot[12x24], Jbase/dJbase/Jee/dJee[6x24], Ag[6x24], dAg_v[6], com[3],
// foot_pos[12], foot_vel[12], ee_pos[3], ee_rot[9]
int orc_rbd(void* hp, const double* q, const double* v, double* M, double* nle, double* Jfoot, double* dJfoot, double* Jbase, double* dJbase, double* Jee, double* dJee,
            double* Ag, double* dAg_v, double* com, double* foot_pos, double* foot_vel, double* ee_pos, double* ee_rot) {
  Handle* h = static_cast<Handle*>(hp);
  return guarded([&] {
    RbdData d; compute_rbd(h->model, q, v, d, 3);
    std::memcpy(M, d.M.a.data(), sizeof(double) * 576); std::memcpy(nle, d.nle.data(), sizeof(double) * 24);
    std::memcpy(Jfoot, d.Jfoot.a.data(), sizeof(double) * 288); std::memcpy(dJfoot, d.dJfoot.a.data(), sizeof(double) * 288);
    std::memcpy(Jbase, d.Jbase.a.data(), sizeof(double) * 144); std::memcpy(dJbase, d.dJbase.a.data(), sizeof(double) * 144);
    std::memcpy(Jee, d.Jee.a.data(), sizeof(double) * 144); std::memcpy(dJee, d.dJee.a.data(), sizeof(double) * 144);
    std::memcpy(Ag, d.Ag.a.data(), sizeof(double) * 144); std::memcpy(dAg_v, d.dAg_v.data(), sizeof(double) * 6);
    for (int i = 0; i < 3; ++i) { com[i] = d.com[i]; ee_pos[i] = d.ee_pos[i]; for (int j = 0; j < 3; ++j) ee_rot[3 * i + j] = d.ee_rot(i, j); }
    for (int f = 0; f < 4; ++f) for (int i = 0; i < 3; ++i) { foot_pos[3 * f + i] = d.foot_pos[f][i]; foot_vel[3 * f + i] = d.foot_vel[f][i]; }
  });
}

// CentroidalModelRbdConversions::computeCentroidalStateFromRbdModel (QMController.cpp:238-241) [upstream, recalled]
int orc_centroidal_state_from_rbd(void* hp, const double* rbd48, double* x30) {
  Handle* h = static_cast<Handle*>(hp);
  return guarded([&] { centroidal_state_from_rbd(h->model, rbd48, x30); });
}

// QMController::updateStateEstimation tail (QMController.cpp:236-243), see ctrl.cpp
int orc_observation_update(void* hp, const double* rbd55, double period, double* t_obs, double* x_obs30) {
  Handle* h = static_cast<Handle*>(hp);
  return guarded([&] { observation_update(h->model, rbd55, period, t_obs, x_obs30); });
}

int orc_wbc_update(void* hp, const double* x_des, const double* u_des, const double* rbd, int mode, double period, double time, double* input_last, int variant,
                   double* cmd54, int* iters3) {
  Handle* h = static_cast<Handle*>(hp);
  return guarded([&] {
    WbcDebug dbg; Vec out = wbc_update(h->model, h->gains, x_des, u_des, rbd, mode, period, time, input_last, variant, &dbg);
    std::memcpy(cmd54, out.data(), sizeof(double) * 54);
    if (iters3) { for (int i = 0; i < 3; ++i) iters3[i] = dbg.hoqp_iterations[i]; }
    if (dbg.qp_status) throw std::runtime_error("oracle QP iteration cap reached");
  });
}

// intermediate quantities of one WBC update (tests): q_meas[24], v_meas[24], q_des[24], v_des[24], base_acc_des[6], level solutions [3x36]
int orc_wbc_debug(void* hp, const double* x_des, const double* u_des, const double* rbd, int mode, double period, double time, const double* input_last_in, int variant,
                  double* q_meas, double* v_meas, double* q_des, double* v_des, double* base_acc, double* levels) {
  Handle* h = static_cast<Handle*>(hp);
  return guarded([&] {
    double il[NU]; std::memcpy(il, input_last_in, sizeof(il));
    WbcDebug dbg; wbc_update(h->model, h->gains, x_des, u_des, rbd, mode, period, time, il, variant, &dbg);
    std::memcpy(q_meas, dbg.q_meas.data(), 24 * 8); std::memcpy(v_meas, dbg.v_meas.data(), 24 * 8); std::memcpy(q_des, dbg.q_des.data(), 24 * 8); std::memcpy(v_des, dbg.v_des.data(), 24 * 8);
    std::memcpy(base_acc, dbg.base_acc_des.data(), 6 * 8);
    for (int l = 0; l < 3; ++l) std::memcpy(levels + 36 * l, dbg.level_solutions[l].data(), 36 * 8);
  });
}

// batch over robots with a std::thread pool (CPU baseline; one robot per task)
int orc_wbc_update_batch(void* hp, int B, const double* x_des, const double* u_des, const double* rbd, const int* mode, const double* period, const double* time,
                         double* input_last, int variant, double* cmd, int nthreads) {
  Handle* h = static_cast<Handle*>(hp);
  std::atomic<int> next(0), fail(0);
  auto work = [&] { for (;;) { int b = next.fetch_add(1); if (b >= B) break;
      try { Vec out = wbc_update(h->model, h->gains, x_des + 30 * b, u_des + 30 * b, rbd + 55 * b, mode[b], period[b], time[b], input_last + 30 * b, variant, nullptr); std::memcpy(cmd + 54 * b, out.data(), 54 * 8); }
      catch (...) { fail++; } } };
  std::vector<std::thread> th; for (int t = 0; t < std::max(1, nthreads); ++t) th.emplace_back(work); for (auto& t : th) t.join();
  return fail.load() ? -1 : 0;
}

}  // extern "C"

#include "capi_mpc.inc"
