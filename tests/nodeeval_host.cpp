// TEST INFRASTRUCTURE: host build (g++) of the thread-per-node evaluator of the product (qm_control_b200/csrc/kernels/node_eval.cuh) so that the CPU suite
// can check it against the oracle without a GPU (tests/test_node_eval_cpu.py).  The evaluator is compiled from the very header the CUDA kernels include.
#include <cstring>
#include <string>

#include "host/qm_config.h"
#include "kernels/node_eval.cuh"

using namespace qmb;

extern "C" {

void* nev_create(const char* task, const char* urdf, const char* reference, const char* gains) {
  try { return new HostModel(build_host_model(task, urdf, reference, gains)); } catch (const std::exception&) { return nullptr; }
}
void nev_destroy(void* h) { delete static_cast<HostModel*>(h); }

// flow map with dense Jacobians assembled from the blocks (the way K2b expands them)
void nev_flow(void* h, const double* x, const double* u, double* f, double* A, double* B) {
  const DevModel* mdl = &static_cast<HostModel*>(h)->dev; ne::BaseKin bk; ne::FlowBlk fb; ne::FlowAcc acc; double d[4][3], JxF[4][9];
  ne::base_eval<true>(mdl, x, bk); ne::flow_acc_init(acc);
  for (int i = 0; i < 4; ++i) ne::foot_eval<true>(mdl, x, u, bk, i, acc, d[i], nullptr, nullptr, nullptr, JxF[i]);
  ne::flow_finish<true>(mdl, x, bk, acc, fb.f, &fb);
  for (int i = 0; i < 12; ++i) f[i] = fb.f[i];
  for (int i = 12; i < 30; ++i) f[i] = u[i];
  std::memset(A, 0, 900 * sizeof(double)); std::memset(B, 0, 900 * sizeof(double));
  const double im = 1.0 / mdl->total_mass;
  for (int a = 0; a < 3; ++a) {
    A[(6 + a) * 30 + a] = 1.0;
    for (int c = 0; c < 3; ++c) { A[(6 + a) * 30 + 3 + c] = fb.Mpc[3 * a + c]; A[(9 + a) * 30 + 3 + c] = fb.Mtw[3 * a + c];
      A[(3 + a) * 30 + 9 + c] = fb.hth[c][a]; A[(6 + a) * 30 + 9 + c] = fb.vp[c][a]; A[(9 + a) * 30 + 9 + c] = fb.vt[c][a]; }
    for (int i = 0; i < 4; ++i) { for (int j = 0; j < 3; ++j) A[(3 + a) * 30 + 12 + mdl->foot_leg[i] + j] = JxF[i][3 * j + a]; B[a * 30 + 3 * i + a] = im; }
  }
  for (int i = 0; i < 4; ++i) for (int a = 0; a < 3; ++a) { double ea[3] = {0, 0, 0}; ea[a] = 1.0; double col[3]; cross3(d[i], ea, col); for (int r = 0; r < 3; ++r) B[(3 + r) * 30 + 3 * i + a] = col[r] * im; }
  for (int j = 12; j < 30; ++j) B[j * 30 + j] = 1.0;
}

// cost value, equality residuals (per foot: 3 velocity rows, swing: row 2 holds v_z - zdot_ref [+ gain * (z - z_ref)]), end-effector error and the Jacobians
// C [4][3][12] and Je [6][12] at one node
int nev_stage(void* h, int ne_, const double* ev, const int* modes, int nk, const double* tt, const double* ts, double t, const double* x, const double* u, int terminal,
              double* cost, double* eq_ss, double* foot_e /*12*/, double* ee_e /*6*/, double* C /*144*/, double* Je /*72*/) {
  const DevModel* mdl = &static_cast<HostModel*>(h)->dev; ne::BaseKin bk; ne::FlowAcc acc; ne::FootBlk fb[4]; double al[9], fe[4][3], pf[4][3];
  ne::base_eval<true>(mdl, x, bk); ne::flow_acc_init(acc);
  const int mode = mode_at_time(ev, modes, ne_, t); int fm = 0; for (int i = 0; i < 4; ++i) if (contact_flag(mode, i)) fm |= 1 << i; if (terminal) fm = 0;
  for (int i = 0; i < 4; ++i) { ne::foot_eval<true>(mdl, x, u, bk, i, acc, fb[i].d, fb[i].pf, fb[i].Jl, al, fb[i].JxF); ne::foot_velocity_1<true>(mdl, x, u, bk, i, fb[i].d, fb[i].Jl, al, fb[i].e, fb[i].C);
    for (int a = 0; a < 3; ++a) { fe[i][a] = fb[i].e[a]; pf[i][a] = fb[i].pf[a]; } std::memcpy(C + 36 * i, fb[i].C, sizeof(fb[i].C)); }
  const ne::TargetSeg sg = ne::target_segment(tt, ts, nk, t); double pref[3], qref[4]; ne::target_pose(sg, nk, pref, qref);
  ne::EeRec ee; ne::ee_eval<true>(mdl, x, bk, pref, qref, ee.e, ee.Je);
  *cost = ne::cost_value(mdl, x, u, sg, ee.e, fm, terminal != 0);
  bool ok = true; *eq_ss = terminal ? 0.0 : ne::equality_ss(mdl, u, fe, pf, fm, ev, modes, ne_, t, &ok);
  std::memcpy(foot_e, fe, sizeof(fe)); std::memcpy(ee_e, ee.e, sizeof(ee.e)); std::memcpy(Je, ee.Je, sizeof(ee.Je));
  return ok ? 0 : 1;
}

}  // extern "C"
