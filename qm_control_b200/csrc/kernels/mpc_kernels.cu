// Batched one-iteration multiple-shooting SQP (the MPC tick of qm_control):
//   SqpSolver::runImpl as QMController configures it (qm_controllers/src/QMController.cpp:287-288, task.info:75-92)
//   [upstream ocs2_sqp / ocs2_oc multiple_shooting, recalled — SURVEY.md App. A.5]:
//     K1 mpc_setup_kernel      timeDiscretizationWithEvents + initializeStateInputTrajectories (QMInitializer.cpp:33-41 when cold)
//     K2 mpc_lq_kernel         setupQuadraticSubproblem: RK2 sensitivities, cost quadratic model, equality constraints, projection
//     K3 mpc_riccati_kernel    OCP-QP (HPIPM without inequality rows ≡ Riccati backward/forward sweep) + armijo metric
//     K4 mpc_linesearch_kernel takeStep: filter line search, trajectory update
// Parallelisation: K2 is node-parallel (one warp per (robot, node)); K3 is one warp per robot (the recursion is
// sequential in time) with every 30x30 block in shared memory; K4 is one CTA per robot, warps striding over nodes.
#include "mpc_api.cuh"
#include "mpc_device.cuh"
#include "wlinalg.cuh"

namespace qmb {

constexpr int LQ_WARPS = 4, RIC_WARPS = 4, LS_WARPS = 4, SETUP_WARPS = 4;
enum { MST_ITER_CAP = 1, MST_OVERFLOW = 2, MST_NAN = 4, MST_NOT_PD = 8, MST_NO_STEP = 16 };

__device__ __forceinline__ double interval_start(double t, int ev) { return ev == 2 ? t + WEAK_EPS : t; }
__device__ __forceinline__ double interval_end(double t, int ev) { return ev == 1 ? t - WEAK_EPS : t; }
__device__ __forceinline__ int flag_mask(int mode) { int m = 0; for (int i = 0; i < 4; ++i) if (contact_flag(mode, i)) m |= 1 << i; return m; }

// =====================================================================================================
// K1: time grid + initial guess
__global__ void __launch_bounds__(32 * SETUP_WARPS) mpc_setup_kernel(const DevModel* __restrict__ mdl, int B, int nmax, MpcProblemDev p, MpcSolutionDev prev, MpcSolutionDev next, int32_t* __restrict__ status) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31; const int b = blockIdx.x * SETUP_WARPS + warp; if (b >= B) return;
  const double t0 = p.t0[b], tf = t0 + mdl->time_horizon, dt = mdl->dt; const int ne = p.n_events[b]; const double* ev = p.event_times + (size_t)b * EMAX; const int32_t* modes = p.modes + (size_t)b * (EMAX + 1);
  double* gt = next.t + (size_t)b * nmax; int32_t* ge = next.event + (size_t)b * nmax; int st = 0;
  // ---- timeDiscretizationWithEvents [upstream ocs2_oc/oc_data/TimeDiscretization.cpp] ----
  int n = 0;
  if (lane == 0) {
    const double dt_min = 10.0 * 1e-9; /* 10 * ocs2 numeric_traits::limitEpsilon [upstream] */ gt[0] = t0; ge[0] = 0; n = 1; int next_ev = lower_bound_idx(ev, ne, t0);
    while (gt[n - 1] < tf) {
      double nt = gt[n - 1] + dt; int nev = 0; bool is_event = false;
      if (next_ev < ne && nt >= ev[next_ev]) { nt = ev[next_ev]; is_event = true; nev = 1; ++next_ev; }
      if (nt >= tf) { is_event = false; nt = tf; nev = 0; }
      if (nt > gt[n - 1] + dt_min) { if (n >= nmax) { st |= MST_OVERFLOW; break; } gt[n] = nt; ge[n] = nev; ++n; } else if (ge[n - 1] != 2) { gt[n - 1] = nt; ge[n - 1] = nev; }
      if (is_event) { if (n >= nmax) { st |= MST_OVERFLOW; break; } gt[n] = nt; ge[n] = 2; ++n; }
    }
    next.n_nodes[b] = n;
  }
  n = __shfl_sync(FULL, n, 0); st = __shfl_sync(FULL, st, 0);
  __syncwarp();
  // ---- initializeStateInputTrajectories [upstream ocs2_oc/multiple_shooting/Initialization.cpp] ----
  const int np = prev.n_nodes ? prev.n_nodes[b] : 0; const bool has_prev = np >= 2;
  const double* pt = prev.t + (size_t)b * nmax; const double* px = prev.x + (size_t)b * nmax * NX; const double* pu = prev.u + (size_t)b * nmax * NU;
  const double state_till = has_prev ? pt[np - 1] : t0, input_till = has_prev ? pt[np - 2] : t0;
  double* gx = next.x + (size_t)b * nmax * NX; double* gu = next.u + (size_t)b * nmax * NU;
  auto interp = [&](const double* traj, double t) { int idx; double a; time_segment(pt, np, t, idx, a); return (lane < NX) ? a * traj[(size_t)idx * NX + lane] + (1.0 - a) * traj[(size_t)(idx + 1 < np ? idx + 1 : idx) * NX + lane] : 0.0; };
  double xk;
  { const double ti = interval_start(gt[0], ge[0]); xk = (has_prev && ti < state_till) ? interp(px, ti) : (lane < NX ? p.x0[(size_t)b * NX + lane] : 0.0); }
  if (lane < NX) gx[lane] = xk;
  for (int k = 0; k < n - 1; ++k) {
    double uk = 0.0;
    if (ge[k] != 1) {
      const double t = interval_start(gt[k], ge[k]), tn = interval_end(gt[k + 1], ge[k + 1]);
      if (!has_prev || t > input_till || tn > state_till) {   // QMInitializer::compute: weight-compensating input, state held
        const int mode = mode_at_time(ev, modes, ne, t); int nst = 0; for (int i = 0; i < 4; ++i) nst += contact_flag(mode, i);
        if (lane < 12 && (lane % 3) == 2 && contact_flag(mode, lane / 3)) uk = mdl->total_mass * 9.81 / nst;
      } else { uk = interp(pu, t); xk = interp(px, tn); }
    }
    if (lane < NX) { gu[(size_t)k * NU + lane] = uk; gx[(size_t)(k + 1) * NX + lane] = xk; }
  }
  if (lane < NX && n >= 1) gu[(size_t)(n - 1) * NU + lane] = 0.0;
  if (lane == 0) status[b] = st;
}

// =====================================================================================================
// K2: linear-quadratic approximation + projection of one node
struct LqSmem {
  PointWs pt; QuadWs quad; CostWs cost; ConWs con;
  double xs[NX], us[NU], xnext[NX], f1[NX], A1r[9 * NX], B1h[36];
  double Ard[9 * NX], Brd[9 * NX], bvec[NX];        // discrete: rows 3:12 of (A_d - I) and of B_d; defect
  double Pxd[MAXDEP * NX], Pud[MAXDEP * MU], Ped[MAXDEP], Pe_full[NU], rs[NU], RPx[NU * 31];
  int dep_idx[MAXDEP], free_idx[MU], col_of_input[NU], piv_dep[4], piv_free[4][2];
};

__global__ void __launch_bounds__(32 * LQ_WARPS) mpc_lq_kernel(const DevModel* __restrict__ mdl, int B, int nmax, MpcProblemDev p, MpcSolutionDev sol, double* __restrict__ stage, int32_t* __restrict__ stage_i, int32_t* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31; const long long gid = (long long)blockIdx.x * LQ_WARPS + warp;
  const int b = (int)(gid / nmax), k = (int)(gid % nmax); if (b >= B) return;
  const int n = sol.n_nodes[b]; if (k >= n) return;
  LqSmem& sm = reinterpret_cast<LqSmem*>(smem_raw)[warp];
  const double* gt = sol.t + (size_t)b * nmax; const int32_t* ge = sol.event + (size_t)b * nmax;
  double* sg = stage + ((size_t)b * nmax + k) * STAGE_DBL; int32_t* si = stage_i + ((size_t)b * nmax + k) * STAGE_INT;
  const int ne = p.n_events[b]; const double* ev = p.event_times + (size_t)b * EMAX; const int32_t* modes = p.modes + (size_t)b * (EMAX + 1);
  const int nk = p.n_target[b]; const double* tt = p.target_times + (size_t)b * KMAX; const double* ts = p.target_states + (size_t)b * KMAX * TARGET_DIM;
  const double* xk = sol.x + ((size_t)b * nmax + k) * NX; const double* uk = sol.u + ((size_t)b * nmax + k) * NU;
  const bool terminal = (k == n - 1);
  if (lane < NX) { sm.xs[lane] = xk[lane]; sm.us[lane] = terminal ? 0.0 : uk[lane]; sm.xnext[lane] = terminal ? 0.0 : xk[NX + lane]; }
  __syncwarp();
  if (!terminal && ge[k] == 1) {   // event node: identity jump map, no input, no cost (setupEventNode)
    double d = 0.0; if (lane < NX) { d = sm.xs[lane] - sm.xnext[lane]; sg[ST_b + lane] = d; }
    const double ss = warp_sum(d * d);
    if (lane == 0) { si[SI_TYPE] = 1; si[SI_M] = 0; si[SI_NDEP] = 0; sg[ST_PERF] = 0.0; sg[ST_PERF + 1] = ss; sg[ST_PERF + 2] = 0.0; }
    return;
  }
  const double t = interval_start(gt[k], ge[k]);
  if (lane < NX) { sm.pt.x[lane] = sm.xs[lane]; sm.pt.u[lane] = sm.us[lane]; }
  __syncwarp();
  if (terminal) {   // setupTerminalNode: finalEndEffector soft constraint only (QMInterface.cpp:104)
    point_eval<false>(mdl, &sm.pt, lane);
    TargetRef ref = target_reference(tt, ts, nk, t, lane);
    const double val = stage_cost<true>(mdl, &sm.pt, &sm.cost, &sm.quad, ref, 0, true, lane);
    for (int e = lane; e < NX * NX; e += 32) sg[ST_Q + e] = sm.quad.Qf[(e / NX) * 31 + (e % NX)];
    if (lane < NX) sg[ST_q + lane] = sm.quad.qf[lane];
    if (lane == 0) { si[SI_TYPE] = 2; si[SI_M] = 0; si[SI_NDEP] = 0; sg[ST_PERF] = val; sg[ST_PERF + 1] = 0.0; sg[ST_PERF + 2] = 0.0; }
    return;
  }
  const double dt = interval_end(gt[k + 1], ge[k + 1]) - t;
  const int mode = mode_at_time(ev, modes, ne, t); const int fm = flag_mask(mode);
  // ---- first flow evaluation at (x,u): dynamics Jacobians, cost, constraints ----
  point_eval<true>(mdl, &sm.pt, lane);
  TargetRef ref = target_reference(tt, ts, nk, t, lane);
  const double cost_val = stage_cost<true>(mdl, &sm.pt, &sm.cost, &sm.quad, ref, fm, false, lane);
  foot_velocity<true>(mdl, &sm.pt, &sm.con, lane);
  // swing references and the dependent-input bookkeeping (one lane per foot)
  int nd_before = 0; for (int i = 0; i < 4; ++i) if (i < lane) nd_before += ((fm >> i) & 1) ? 3 : 4;
  int ndep = 0; for (int i = 0; i < 4; ++i) ndep += ((fm >> i) & 1) ? 3 : 4;
  const int m = NU - ndep;
  for (int e = lane; e < MAXDEP * MU; e += 32) sm.Pud[e] = 0.0;
  for (int e = lane; e < MAXDEP * NX; e += 32) sm.Pxd[e] = 0.0;
  if (lane < NU) sm.Pe_full[lane] = 0.0;
  double eq_ss = 0.0; bool swing_ok = true; int pivot = -1;
  if (lane < 4) {
    const int i = lane; const int first = mdl->foot_leg[i];
    if ((fm >> i) & 1) { for (int j = 0; j < 3; ++j) sm.dep_idx[nd_before + j] = 12 + first + j; for (int a = 0; a < 3; ++a) eq_ss += sm.con.e[i][a] * sm.con.e[i][a]; }
    else {
      double zp, zv; swing_ok = swing_reference(mdl, ev, modes, ne, i, t, zp, zv);
      double ez = sm.con.e[i][2] - zv; if (mdl->position_error_gain != 0.0) ez += mdl->position_error_gain * (sm.pt.pf[i][2] - zp);
      sm.con.e[i][2] = ez;
      for (int a = 0; a < 3; ++a) { sm.dep_idx[nd_before + a] = 3 * i + a; eq_ss += sm.pt.u[3 * i + a] * sm.pt.u[3 * i + a]; }
      eq_ss += ez * ez;
      // pivot joint of the normal-velocity row: largest |d v_z / d qdot_j|
      double best = -1.0; for (int j = 0; j < 3; ++j) { const double a = fabs(sm.pt.Jl[i][3 * j + 2]); if (a > best) { best = a; pivot = j; } }
      sm.dep_idx[nd_before + 3] = 12 + first + pivot;
    }
  }
  eq_ss = warp_sum(eq_ss);
  if (!__all_sync(FULL, swing_ok)) { if (lane == 0) atomicOr(&status[b], MST_OVERFLOW); }
  __syncwarp();
  // free / dependent partition of the 30 inputs
  bool is_dep = false; if (lane < NU) for (int d = 0; d < ndep; ++d) is_dep |= (sm.dep_idx[d] == lane);
  const unsigned free_mask = __ballot_sync(FULL, lane < NU && !is_dep);
  if (lane < NU) { const int rank = __popc(free_mask & ((1u << lane) - 1u)); sm.col_of_input[lane] = is_dep ? -1 : rank; if (!is_dep) sm.free_idx[rank] = lane; }
  __syncwarp();
  // projection rows (structured elimination; the projected optimum does not depend on the null-space basis)
  if (lane < 4) {
    const int i = lane; const int first = mdl->foot_leg[i];
    if ((fm >> i) & 1) {   // zero velocity: Jl dqd = -(C dx + e)  →  dqd = -Jl^{-1} (C dx + e)
      double Jm[9], Ji[9]; for (int a = 0; a < 3; ++a) for (int j = 0; j < 3; ++j) Jm[3 * a + j] = sm.pt.Jl[i][3 * j + a]; inv3(Jm, Ji);
      for (int j = 0; j < 3; ++j) { const int d = nd_before + j; double pe = 0.0; for (int a = 0; a < 3; ++a) pe -= Ji[3 * j + a] * sm.con.e[i][a]; sm.Ped[d] = pe;
        for (int c = 0; c < NX; ++c) { double s = 0.0; for (int a = 0; a < 3; ++a) s -= Ji[3 * j + a] * sm.con.C[i][a][c]; sm.Pxd[d * NX + c] = s; } }
      sm.piv_dep[i] = -1;
    } else {               // zero force: dF = -F ; normal velocity: pivot joint eliminated
      for (int a = 0; a < 3; ++a) { sm.Ped[nd_before + a] = -sm.pt.u[3 * i + a]; }
      const int d = nd_before + 3; const double piv = sm.pt.Jl[i][3 * pivot + 2];
      sm.Ped[d] = -sm.con.e[i][2] / piv; for (int c = 0; c < NX; ++c) sm.Pxd[d * NX + c] = -sm.con.C[i][2][c] / piv;
      int nf = 0; for (int j = 0; j < 3; ++j) if (j != pivot) { const int col = sm.col_of_input[12 + first + j]; sm.Pud[d * MU + col] = -sm.pt.Jl[i][3 * j + 2] / piv; sm.piv_free[i][nf++] = col; }
      sm.piv_dep[i] = d;
    }
  }
  __syncwarp();
  if (lane < ndep) sm.Pe_full[sm.dep_idx[lane]] = sm.Ped[lane];
  // keep k1 data, then second flow evaluation at x + c dt k1
  if (lane < NX) sm.f1[lane] = sm.pt.f[lane];
  for (int e = lane; e < 9 * NX; e += 32) sm.A1r[e] = sm.pt.Ar[e];
  for (int e = lane; e < 36; e += 32) sm.B1h[e] = sm.pt.Bh[e];
  __syncwarp();
  if (lane < NX) sm.pt.x[lane] = sm.xs[lane] + mdl->rk_c * dt * sm.f1[lane];
  __syncwarp();
  point_eval<true>(mdl, &sm.pt, lane);
  const double w1 = mdl->rk_w1, w2 = mdl->rk_w2, cdt = mdl->rk_c * dt, mass = mdl->total_mass, dtw = dt * (w1 + w2);
  // defect b = x + dt (w1 k1 + w2 k2) - x_{k+1}
  double bb = 0.0; if (lane < NX) { bb = sm.xs[lane] + dt * (w1 * sm.f1[lane] + w2 * sm.pt.f[lane]) - sm.xnext[lane]; sm.bvec[lane] = bb; }
  const double dyn_ss = warp_sum(bb * bb);
  // A_d - I (rows 3:12) = dt (w1 A1 + w2 (A2 + c dt A2 A1)) ; B_d rows 3:12 = dt (w1 B1 + w2 (B2 + c dt A2 B1))
  for (int e = lane; e < 9 * NX; e += 32) {
    const int r = e / NX, c = e % NX; const double* a2 = sm.pt.Ar + r * NX;
    double aa = 0.0; for (int q = 0; q < 9; ++q) aa += a2[3 + q] * sm.A1r[q * NX + c];
    sm.Ard[e] = dt * (w1 * sm.A1r[e] + w2 * (a2[c] + cdt * aa));
    double b1 = 0.0, b2 = 0.0, ab = 0.0;
    if (c < 12) { if (r < 3) { b1 = sm.B1h[r * 12 + c]; b2 = sm.pt.Bh[r * 12 + c]; } ab = a2[c % 3] / mass; for (int q = 0; q < 3; ++q) ab += a2[3 + q] * sm.B1h[q * 12 + c]; }
    else ab = a2[c];
    sm.Brd[e] = dt * (w1 * b1 + w2 * (b2 + cdt * ab));
  }
  __syncwarp();
  // ---- projected dynamics: A~ = A_d + B_d Px, B~ = B_d Pu, b~ = b + B_d Pe ----
  // rows 0:3 and 12:30 of B_d are dtw/m selectors (forces) and dtw identity (joint velocities)
  if (lane < NX) {
    const int i = lane; double* Arow = sg + ST_A + (size_t)i * NX; double* Brow = sg + ST_B + (size_t)i * MU;
    double bt = sm.bvec[i];
    if (i < 3) {
      for (int c = 0; c < NX; ++c) Arow[c] = (c == i) ? 1.0 : 0.0;
      for (int a = 0; a < m; ++a) { const int fa = sm.free_idx[a]; Brow[a] = (fa < 12 && fa % 3 == i) ? dtw / mass : 0.0; }
      for (int f = 0; f < 4; ++f) bt += (dtw / mass) * sm.Pe_full[3 * f + i];
    } else if (i < 12) {
      const int r = i - 3; const double* br = sm.Brd + r * NX;
      for (int c = 0; c < NX; ++c) { double s = sm.Ard[r * NX + c] + ((c == i) ? 1.0 : 0.0); for (int d = 0; d < ndep; ++d) s += br[sm.dep_idx[d]] * sm.Pxd[d * NX + c]; Arow[c] = s; }
      for (int a = 0; a < m; ++a) { double s = br[sm.free_idx[a]]; for (int f = 0; f < 4; ++f) { const int d = sm.piv_dep[f]; if (d >= 0) s += br[sm.dep_idx[d]] * sm.Pud[d * MU + a]; } Brow[a] = s; }
      for (int d = 0; d < ndep; ++d) bt += br[sm.dep_idx[d]] * sm.Ped[d];
    } else {
      const int col = sm.col_of_input[i];   // input i (joint velocity) is free (col >= 0) or dependent
      int dd = -1; if (col < 0) for (int d = 0; d < ndep; ++d) if (sm.dep_idx[d] == i) dd = d;
      for (int c = 0; c < NX; ++c) Arow[c] = ((c == i) ? 1.0 : 0.0) + (dd >= 0 ? dtw * sm.Pxd[dd * NX + c] : 0.0);
      for (int a = 0; a < m; ++a) Brow[a] = (col == a) ? dtw : (dd >= 0 ? dtw * sm.Pud[dd * MU + a] : 0.0);
      if (dd >= 0) bt += dtw * sm.Ped[dd];
    }
    for (int a = m; a < MU; ++a) Brow[a] = 0.0;
    sg[ST_b + i] = bt;
  }
  // ---- projected cost (changeOfInputVariables [upstream]); quadratic model scaled by dt ----
  // rs = r + R Pe ; RPx = R Px
  if (lane < NU) { double s = sm.quad.rf[lane]; for (int d = 0; d < ndep; ++d) s += sm.quad.Rf[lane * 31 + sm.dep_idx[d]] * sm.Ped[d]; sm.rs[lane] = s; }
  for (int e = lane; e < NU * NX; e += 32) { const int i = e / NX, c = e % NX; double s = 0.0; for (int d = 0; d < ndep; ++d) s += sm.quad.Rf[i * 31 + sm.dep_idx[d]] * sm.Pxd[d * NX + c]; sm.RPx[i * 31 + c] = s; }
  __syncwarp();
  if (lane < NX) {   // q~ = q + Px' rs ; Q~ = Q + Px' R Px   (the cost has no state-input cross term before projection)
    double s = sm.quad.qf[lane]; for (int d = 0; d < ndep; ++d) s += sm.Pxd[d * NX + lane] * sm.rs[sm.dep_idx[d]]; sg[ST_q + lane] = dt * s;
    double* Qrow = sg + ST_Q + (size_t)lane * NX;
    for (int c = 0; c < NX; ++c) { double q = sm.quad.Qf[lane * 31 + c]; for (int d = 0; d < ndep; ++d) q += sm.Pxd[d * NX + lane] * sm.RPx[sm.dep_idx[d] * 31 + c]; Qrow[c] = dt * q; }
  }
  if (lane < MU) {   // r~ = Pu' rs ; S~ = Pu' (R Px) ; R~ = Pu' R Pu
    const int a = lane; double* Srow = sg + ST_S + (size_t)a * NX; double* Rrow = sg + ST_R + (size_t)a * MU;
    if (a < m) {
      const int fa = sm.free_idx[a]; double r = sm.rs[fa];
      for (int f = 0; f < 4; ++f) { const int d = sm.piv_dep[f]; if (d >= 0) r += sm.Pud[d * MU + a] * sm.rs[sm.dep_idx[d]]; }
      sg[ST_r + a] = dt * r;
      for (int c = 0; c < NX; ++c) { double s = sm.RPx[fa * 31 + c]; for (int f = 0; f < 4; ++f) { const int d = sm.piv_dep[f]; if (d >= 0) s += sm.Pud[d * MU + a] * sm.RPx[sm.dep_idx[d] * 31 + c]; } Srow[c] = dt * s; }
      for (int c = 0; c < m; ++c) {
        const int fc = sm.free_idx[c]; double s = sm.quad.Rf[fa * 31 + fc];
        for (int f = 0; f < 4; ++f) { const int d = sm.piv_dep[f]; if (d < 0) continue; const int di = sm.dep_idx[d];
          s += sm.Pud[d * MU + a] * sm.quad.Rf[di * 31 + fc] + sm.quad.Rf[fa * 31 + di] * sm.Pud[d * MU + c];
          for (int g = 0; g < 4; ++g) { const int d2 = sm.piv_dep[g]; if (d2 >= 0) s += sm.Pud[d * MU + a] * sm.quad.Rf[di * 31 + sm.dep_idx[d2]] * sm.Pud[d2 * MU + c]; } }
        Rrow[c] = dt * s;
      }
      for (int c = m; c < MU; ++c) Rrow[c] = 0.0;
    } else { sg[ST_r + a] = 0.0; for (int c = 0; c < NX; ++c) Srow[c] = 0.0; for (int c = 0; c < MU; ++c) Rrow[c] = (c == a) ? 1.0 : 0.0; }
  }
  // projection data for the forward pass
  for (int e = lane; e < MAXDEP * NX; e += 32) sg[ST_PXD + e] = sm.Pxd[e];
  for (int e = lane; e < MAXDEP * MU; e += 32) sg[ST_PUD + e] = sm.Pud[e];
  if (lane < MAXDEP) { sg[ST_PED + lane] = (lane < ndep) ? sm.Ped[lane] : 0.0; si[SI_DEP + lane] = (lane < ndep) ? sm.dep_idx[lane] : -1; }
  if (lane < MU) si[SI_FREE + lane] = (lane < m) ? sm.free_idx[lane] : -1;
  if (lane == 0) { si[SI_TYPE] = 0; si[SI_M] = m; si[SI_NDEP] = ndep; sg[ST_PERF] = dt * cost_val; sg[ST_PERF + 1] = dt * dyn_ss; sg[ST_PERF + 2] = dt * eq_ss; }
}

// =====================================================================================================
// K3: Riccati backward sweep + forward rollout of the projected LQ problem (one warp per robot)
constexpr int LDX = 31, LDU = MU + 1;
struct RicSmem {
  double P[NX * LDX], A[NX * LDX], W[NX * LDX];       // value Hessian, stage A~, P A~
  double Bm[NX * LDU], PB[NX * LDU], G[MU * LDX], H[MU * LDU];
  double p[NX], b[NX], q[NX], r[MU], pPb[NX], h[MU], dx[NX], dut[MU], tmp[NX];
};
// copy a rows x cols matrix from global (row stride gs) into shared (row stride ls)
__device__ __forceinline__ void load_mat(const double* __restrict__ g, int rows, int cols, int gs, double* s, int ls, int lane) {
  for (int e = lane; e < rows * cols; e += 32) { const int i = e / cols, j = e % cols; s[i * ls + j] = g[(size_t)i * gs + j]; }
}

__global__ void __launch_bounds__(32 * RIC_WARPS) mpc_riccati_kernel(const DevModel* __restrict__ mdl, int B, int nmax, MpcProblemDev p, MpcSolutionDev sol, const double* __restrict__ stage, const int32_t* __restrict__ stage_i,
                                                                   double* __restrict__ gains, double* __restrict__ dxo, double* __restrict__ duo, double* __restrict__ robot, int32_t* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31; const int b = blockIdx.x * RIC_WARPS + warp; if (b >= B) return;
  RicSmem& sm = reinterpret_cast<RicSmem*>(smem_raw)[warp];
  const int n = sol.n_nodes[b]; const int N = n - 1; int st = 0;
  const double* sgb = stage + (size_t)b * nmax * STAGE_DBL; const int32_t* sib = stage_i + (size_t)b * nmax * STAGE_INT; double* gb = gains + (size_t)b * nmax * GAIN_DBL;
  // terminal value function and baseline performance
  load_mat(sgb + (size_t)N * STAGE_DBL + ST_Q, NX, NX, NX, sm.P, LDX, lane);
  if (lane < NX) sm.p[lane] = sgb[(size_t)N * STAGE_DBL + ST_q + lane];
  double perf[3] = {0, 0, 0};
  for (int k = lane; k < n; k += 32) { const double* pf = sgb + (size_t)k * STAGE_DBL + ST_PERF; perf[0] += pf[0]; perf[1] += pf[1]; perf[2] += pf[2]; }
  for (int i = 0; i < 3; ++i) perf[i] = warp_sum(perf[i]);
  { double d = 0.0; if (lane < NX) { d = p.x0[(size_t)b * NX + lane] - sol.x[(size_t)b * nmax * NX + lane]; sm.dx[lane] = d; } perf[1] += warp_sum(d * d); }
  __syncwarp();
  for (int k = N - 1; k >= 0; --k) {
    const double* sg = sgb + (size_t)k * STAGE_DBL; const int32_t* si = sib + (size_t)k * STAGE_INT; const int type = si[SI_TYPE], m = si[SI_M];
    if (lane < NX) sm.b[lane] = sg[ST_b + lane];
    __syncwarp();
    if (lane < NX) { double s = sm.p[lane]; for (int j = 0; j < NX; ++j) s += sm.P[lane * LDX + j] * sm.b[j]; sm.pPb[lane] = s; }   // p + P b
    __syncwarp();
    if (type == 1) { if (lane < NX) sm.p[lane] = sm.pPb[lane]; __syncwarp(); continue; }   // event: A = I, no input
    load_mat(sg + ST_A, NX, NX, NX, sm.A, LDX, lane); load_mat(sg + ST_B, NX, m, MU, sm.Bm, LDU, lane);
    load_mat(sg + ST_S, m, NX, NX, sm.G, LDX, lane); load_mat(sg + ST_R, m, m, MU, sm.H, LDU, lane);
    if (lane < NX) sm.q[lane] = sg[ST_q + lane]; if (lane < m) sm.r[lane] = sg[ST_r + lane];
    __syncwarp();
    // W = P A ; PB = P B        (lane = row; P[i][:] walks shared memory conflict-free, A rows are broadcast)
    if (lane < NX) {
      double acc[NX];
#pragma unroll
      for (int j = 0; j < NX; ++j) acc[j] = 0.0;
      for (int kk = 0; kk < NX; ++kk) { const double pik = sm.P[lane * LDX + kk]; const double* ar = sm.A + kk * LDX;
#pragma unroll
        for (int j = 0; j < NX; ++j) acc[j] = fma(pik, ar[j], acc[j]); }
#pragma unroll
      for (int j = 0; j < NX; ++j) sm.W[lane * LDX + j] = acc[j];
      double accb[MU];
#pragma unroll
      for (int j = 0; j < MU; ++j) accb[j] = 0.0;
      for (int kk = 0; kk < NX; ++kk) { const double pik = sm.P[lane * LDX + kk]; const double* br = sm.Bm + kk * LDU;
#pragma unroll
        for (int j = 0; j < MU; ++j) accb[j] = fma(pik, br[j], accb[j]); }
#pragma unroll
      for (int j = 0; j < MU; ++j) sm.PB[lane * LDU + j] = accb[j];
    }
    __syncwarp();
    // G = S + B'W ; H = R + B'PB ; h = r + B'(p + P b)      (lane = projected input)
    if (lane < m) {
      double acc[NX];
#pragma unroll
      for (int j = 0; j < NX; ++j) acc[j] = sm.G[lane * LDX + j];
      double hh = sm.r[lane];
      for (int kk = 0; kk < NX; ++kk) { const double bka = sm.Bm[kk * LDU + lane]; const double* wr = sm.W + kk * LDX; hh = fma(bka, sm.pPb[kk], hh);
#pragma unroll
        for (int j = 0; j < NX; ++j) acc[j] = fma(bka, wr[j], acc[j]); }
#pragma unroll
      for (int j = 0; j < NX; ++j) sm.G[lane * LDX + j] = acc[j];
      sm.h[lane] = hh;
      double acch[MU];
#pragma unroll
      for (int j = 0; j < MU; ++j) acch[j] = sm.H[lane * LDU + j];
      for (int kk = 0; kk < NX; ++kk) { const double bka = sm.Bm[kk * LDU + lane]; const double* pr = sm.PB + kk * LDU;
#pragma unroll
        for (int j = 0; j < MU; ++j) acch[j] = fma(bka, pr[j], acch[j]); }
#pragma unroll
      for (int j = 0; j < MU; ++j) sm.H[lane * LDU + j] = acch[j];
    }
    __syncwarp();
    // P <- Q + A'W ; p <- q + A'(p + P b)
    if (lane < NX) {
      double acc[NX]; const double* qrow = sg + ST_Q + (size_t)lane * NX;
#pragma unroll
      for (int j = 0; j < NX; ++j) acc[j] = qrow[j];
      double pp = sm.q[lane];
      for (int kk = 0; kk < NX; ++kk) { const double aki = sm.A[kk * LDX + lane]; const double* wr = sm.W + kk * LDX; pp = fma(aki, sm.pPb[kk], pp);
#pragma unroll
        for (int j = 0; j < NX; ++j) acc[j] = fma(aki, wr[j], acc[j]); }
#pragma unroll
      for (int j = 0; j < NX; ++j) sm.P[lane * LDX + j] = acc[j];
      sm.p[lane] = pp;
    }
    __syncwarp();
    // symmetrise H, factor, Y = L^{-1} [G | h]
    if (lane < m) for (int j = 0; j < lane; ++j) { const double a = 0.5 * (sm.H[lane * LDU + j] + sm.H[j * LDU + lane]); sm.H[lane * LDU + j] = a; }
    __syncwarp();
    if (!w_cholesky(sm.H, m, LDU, lane)) { st |= MST_NOT_PD; break; }
    if (lane <= NX) {   // columns 0..29 of G and column 30 = h
      for (int a = 0; a < m; ++a) { double s = (lane < NX) ? sm.G[a * LDX + lane] : sm.h[a]; for (int c = 0; c < a; ++c) s -= sm.H[a * LDU + c] * ((lane < NX) ? sm.G[c * LDX + lane] : sm.h[c]); s /= sm.H[a * LDU + a]; if (lane < NX) sm.G[a * LDX + lane] = s; else sm.h[a] = s; }
    }
    __syncwarp();
    // P -= Y'Y ; p -= Y' yh ; symmetrise
    if (lane < NX) {
      double acc[NX];
#pragma unroll
      for (int j = 0; j < NX; ++j) acc[j] = sm.P[lane * LDX + j];
      double pp = sm.p[lane];
      for (int a = 0; a < m; ++a) { const double yai = sm.G[a * LDX + lane]; const double* yr = sm.G + a * LDX; pp = fma(-yai, sm.h[a], pp);
#pragma unroll
        for (int j = 0; j < NX; ++j) acc[j] = fma(-yai, yr[j], acc[j]); }
#pragma unroll
      for (int j = 0; j < NX; ++j) sm.P[lane * LDX + j] = acc[j];
      sm.p[lane] = pp;
    }
    __syncwarp();
    if (lane < NX) for (int j = 0; j < lane; ++j) { const double a = 0.5 * (sm.P[lane * LDX + j] + sm.P[j * LDX + lane]); sm.P[lane * LDX + j] = a; }
    __syncwarp();
    if (lane < NX) for (int j = lane + 1; j < NX; ++j) sm.P[lane * LDX + j] = sm.P[j * LDX + lane];
    // K = -L^{-T} Y ; k = -L^{-T} yh     → global
    if (lane <= NX) {
      for (int a = m - 1; a >= 0; --a) { double s = (lane < NX) ? sm.G[a * LDX + lane] : sm.h[a]; for (int c = a + 1; c < m; ++c) s -= sm.H[c * LDU + a] * ((lane < NX) ? sm.G[c * LDX + lane] : sm.h[c]); s /= sm.H[a * LDU + a]; if (lane < NX) sm.G[a * LDX + lane] = s; else sm.h[a] = s; }
    }
    __syncwarp();
    double* gk = gb + (size_t)k * GAIN_DBL;
    for (int e = lane; e < m * NX; e += 32) gk[e] = -sm.G[(e / NX) * LDX + (e % NX)];
    if (lane < m) gk[MU * NX + lane] = -sm.h[lane];
    __syncwarp();
  }
  // ---- forward rollout: dx_{k+1} = A~ dx + B~ du~ + b~ ; du = Px dx + Pu du~ + Pe ; armijo = sum q~'dx + r~'du~ ----
  double armijo = 0.0, dxn2 = 0.0, dun2 = 0.0;
  if (!(st & MST_NOT_PD)) {
    for (int k = 0; k < N; ++k) {
      const double* sg = sgb + (size_t)k * STAGE_DBL; const int32_t* si = sib + (size_t)k * STAGE_INT; const int type = si[SI_TYPE], m = si[SI_M], ndep = si[SI_NDEP];
      double* dxk = dxo + ((size_t)b * nmax + k) * NX; double* duk = duo + ((size_t)b * nmax + k) * NU;
      const double dxi = (lane < NX) ? sm.dx[lane] : 0.0; if (lane < NX) dxk[lane] = dxi; dxn2 += dxi * dxi;
      if (type == 1) { if (lane < NX) { duk[lane] = 0.0; sm.tmp[lane] = dxi + sg[ST_b + lane]; } __syncwarp(); if (lane < NX) sm.dx[lane] = sm.tmp[lane]; __syncwarp(); continue; }
      const double* gk = gb + (size_t)k * GAIN_DBL;
      load_mat(gk, m, NX, NX, sm.G, LDX, lane); load_mat(sg + ST_A, NX, NX, NX, sm.A, LDX, lane); load_mat(sg + ST_B, NX, m, MU, sm.Bm, LDU, lane);
      __syncwarp();
      double dut = 0.0; if (lane < m) { dut = gk[MU * NX + lane]; for (int j = 0; j < NX; ++j) dut += sm.G[lane * LDX + j] * sm.dx[j]; sm.dut[lane] = dut; }
      armijo += ((lane < NX) ? sg[ST_q + lane] * dxi : 0.0) + ((lane < m) ? sg[ST_r + lane] * dut : 0.0);
      __syncwarp();
      if (lane < NX) { double s = sg[ST_b + lane]; for (int j = 0; j < NX; ++j) s += sm.A[lane * LDX + j] * sm.dx[j]; for (int a = 0; a < m; ++a) s += sm.Bm[lane * LDU + a] * sm.dut[a]; sm.tmp[lane] = s; }
      // full-space input step
      if (lane < m) duk[si[SI_FREE + lane]] = dut;
      double dud = 0.0;
      if (lane < ndep) { dud = sg[ST_PED + lane]; const double* px = sg + ST_PXD + (size_t)lane * NX; const double* pu = sg + ST_PUD + (size_t)lane * MU; for (int j = 0; j < NX; ++j) dud += px[j] * sm.dx[j]; for (int a = 0; a < m; ++a) dud += pu[a] * sm.dut[a]; duk[si[SI_DEP + lane]] = dud; }
      dun2 += ((lane < m) ? dut * dut : 0.0) + dud * dud;
      __syncwarp();
      if (lane < NX) sm.dx[lane] = sm.tmp[lane];
      __syncwarp();
    }
    { const double dxi = (lane < NX) ? sm.dx[lane] : 0.0; if (lane < NX) { dxo[((size_t)b * nmax + N) * NX + lane] = dxi; duo[((size_t)b * nmax + N) * NU + lane] = 0.0; } dxn2 += dxi * dxi;
      armijo += (lane < NX) ? sgb[(size_t)N * STAGE_DBL + ST_q + lane] * dxi : 0.0; }
  }
  armijo = warp_sum(armijo); dxn2 = warp_sum(dxn2); dun2 = warp_sum(dun2);
  if (lane == 0) { double* rb = robot + (size_t)b * ROBOT_DBL; rb[0] = armijo; rb[1] = perf[0]; rb[2] = perf[1]; rb[3] = perf[2]; rb[4] = sqrt(dxn2); rb[5] = sqrt(dun2); if (st) atomicOr(&status[b], st); }
}

// =====================================================================================================
// K4: filter line search (one CTA per robot; warps stride over nodes) + trajectory update + input fix-up
struct LsSmem { PointWs pt; CostWs cost; ConWs con; double xa[NX], ua[NU], xna[NX], f1[NX]; };

__device__ __forceinline__ void fixup_inputs(MpcSolutionDev sol, int b, int nmax, int n, int tid, int nthreads) {
  // toPrimalSolution [upstream]: input at a pre-event node repeats the previous one; last input repeated
  const int32_t* ge = sol.event + (size_t)b * nmax; double* gu = sol.u + (size_t)b * nmax * NU;
  for (int k = 1; k < n; ++k) { const bool copy = (k == n - 1) || (ge[k] == 1); if (copy) { for (int i = tid; i < NU; i += nthreads) gu[(size_t)k * NU + i] = gu[(size_t)(k - 1) * NU + i]; } __syncthreads(); }
}

__global__ void __launch_bounds__(32 * LS_WARPS) mpc_linesearch_kernel(const DevModel* __restrict__ mdl, int B, int nmax, MpcProblemDev p, MpcSolutionDev sol, const double* __restrict__ dxo, const double* __restrict__ duo,
                                                                     const double* __restrict__ robot, int32_t* __restrict__ status, double* __restrict__ step_info) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[LS_WARPS][3]; __shared__ int decision;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31; const int b = blockIdx.x; if (b >= B) return;
  LsSmem& sm = reinterpret_cast<LsSmem*>(smem_raw)[warp];
  const int n = sol.n_nodes[b]; const int N = n - 1;
  const double* gt = sol.t + (size_t)b * nmax; const int32_t* ge = sol.event + (size_t)b * nmax;
  double* gx = sol.x + (size_t)b * nmax * NX; double* gu = sol.u + (size_t)b * nmax * NU; const double* gdx = dxo + (size_t)b * nmax * NX; const double* gdu = duo + (size_t)b * nmax * NU;
  const int ne = p.n_events[b]; const double* ev = p.event_times + (size_t)b * EMAX; const int32_t* modes = p.modes + (size_t)b * (EMAX + 1);
  const int nk = p.n_target[b]; const double* tt = p.target_times + (size_t)b * KMAX; const double* ts = p.target_states + (size_t)b * KMAX * TARGET_DIM;
  const double* rb = robot + (size_t)b * ROBOT_DBL; const double armijo = rb[0], base_cost = rb[1], base_viol = sqrt(rb[2] + rb[3]), dxn = rb[4], dun = rb[5];
  const bool failed = (status[b] & MST_NOT_PD) != 0;
  double alpha = 1.0; bool accepted = false; double sc = base_cost, sd = rb[2], se = rb[3];
  const double w1 = mdl->rk_w1, w2 = mdl->rk_w2;
  while (!failed) {
    double cost = 0.0, dyn = 0.0, eq = 0.0;
    for (int k = warp; k <= N; k += LS_WARPS) {
      if (lane < NX) { sm.xa[lane] = gx[(size_t)k * NX + lane] + alpha * gdx[(size_t)k * NX + lane]; if (k < N) { sm.ua[lane] = gu[(size_t)k * NU + lane] + alpha * gdu[(size_t)k * NU + lane]; sm.xna[lane] = gx[(size_t)(k + 1) * NX + lane] + alpha * gdx[(size_t)(k + 1) * NX + lane]; } else sm.ua[lane] = 0.0; }
      __syncwarp();
      if (k == 0) { double d = (lane < NX) ? p.x0[(size_t)b * NX + lane] - sm.xa[lane] : 0.0; dyn += warp_sum(d * d); }
      if (k < N && ge[k] == 1) { double d = (lane < NX) ? sm.xa[lane] - sm.xna[lane] : 0.0; dyn += warp_sum(d * d); __syncwarp(); continue; }
      const double t = interval_start(gt[k], ge[k]);
      if (lane < NX) { sm.pt.x[lane] = sm.xa[lane]; sm.pt.u[lane] = sm.ua[lane]; }
      __syncwarp();
      point_eval<false>(mdl, &sm.pt, lane);
      TargetRef ref = target_reference(tt, ts, nk, t, lane);
      if (k == N) { cost += stage_cost<false>(mdl, &sm.pt, &sm.cost, (QuadWs*)nullptr, ref, 0, true, lane); __syncwarp(); continue; }
      const double dt = interval_end(gt[k + 1], ge[k + 1]) - t; const int mode = mode_at_time(ev, modes, ne, t); const int fm = flag_mask(mode);
      cost += dt * stage_cost<false>(mdl, &sm.pt, &sm.cost, (QuadWs*)nullptr, ref, fm, false, lane);
      foot_velocity<false>(mdl, &sm.pt, &sm.con, lane);
      double es = 0.0;
      if (lane < 4) { const int i = lane; if ((fm >> i) & 1) { for (int a = 0; a < 3; ++a) es += sm.con.e[i][a] * sm.con.e[i][a]; }
        else { double zp, zv; swing_reference(mdl, ev, modes, ne, i, t, zp, zv); double ez = sm.con.e[i][2] - zv; if (mdl->position_error_gain != 0.0) ez += mdl->position_error_gain * (sm.pt.pf[i][2] - zp); es += ez * ez; for (int a = 0; a < 3; ++a) es += sm.ua[3 * i + a] * sm.ua[3 * i + a]; } }
      eq += dt * warp_sum(es);
      if (lane < NX) { sm.f1[lane] = sm.pt.f[lane]; }
      __syncwarp();
      if (lane < NX) sm.pt.x[lane] = sm.xa[lane] + mdl->rk_c * dt * sm.f1[lane];
      __syncwarp();
      point_eval<false>(mdl, &sm.pt, lane);
      double d = (lane < NX) ? sm.xa[lane] + dt * (w1 * sm.f1[lane] + w2 * sm.pt.f[lane]) - sm.xna[lane] : 0.0; dyn += dt * warp_sum(d * d);
      __syncwarp();
    }
    if (lane == 0) { red[warp][0] = cost; red[warp][1] = dyn; red[warp][2] = eq; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double c = 0, dd = 0, e = 0; for (int w = 0; w < LS_WARPS; ++w) { c += red[w][0]; dd += red[w][1]; e += red[w][2]; }
      // FilterLinesearch::acceptStep [upstream ocs2_oc/search_strategy/FilterLinesearch.cpp]
      const double sv = sqrt(dd + e), am = alpha * armijo; bool acc;
      if (sv > mdl->g_max) acc = sv < (1.0 - mdl->gamma_c) * base_viol;
      else if (sv < mdl->g_min && base_viol < mdl->g_min && am < 0.0) acc = c < base_cost + mdl->armijo_factor * am;
      else acc = (c < base_cost - mdl->gamma_c * base_viol) || (sv < (1.0 - mdl->gamma_c) * base_viol);
      red[0][0] = c; red[0][1] = dd; red[0][2] = e;
      int dec = 0; if (acc) dec = 1; else { const double an = alpha * mdl->alpha_decay; if ((an * dxn < mdl->delta_tol && an * dun < mdl->delta_tol) || an < mdl->alpha_min) dec = 2; }
      decision = dec;
    }
    __syncthreads();
    const int dec = decision; sc = red[0][0]; sd = red[0][1]; se = red[0][2];
    __syncthreads();
    if (dec == 1) { accepted = true; break; }
    if (dec == 2) break;
    alpha *= mdl->alpha_decay;
  }
  if (accepted) {
    for (int e = threadIdx.x; e < n * NX; e += blockDim.x) { gx[e] += alpha * gdx[e]; if (e < N * NU) gu[e] += alpha * gdu[e]; }
  } else { alpha = 0.0; sc = base_cost; sd = rb[2]; se = rb[3]; }
  __syncthreads();
  fixup_inputs(sol, b, nmax, n, threadIdx.x, blockDim.x);
  if (threadIdx.x == 0) { if (!accepted) atomicOr(&status[b], MST_NO_STEP); double* si = step_info + (size_t)b * 4; si[0] = alpha; si[1] = sc; si[2] = sd; si[3] = se; }
}

__global__ void mpc_fixup_kernel(int B, int nmax, MpcSolutionDev sol) { const int b = blockIdx.x; if (b >= B) return; const int n = sol.n_nodes[b]; if (n >= 2) fixup_inputs(sol, b, nmax, n, threadIdx.x, blockDim.x); }

// =====================================================================================================
// MPC_MRT_Interface::evaluatePolicy with a feed-forward policy (QMController.cpp:141): linear interpolation of the stored solution, modeAtTime
__global__ void mpc_policy_eval_kernel(int B, int nmax, MpcSolutionDev sol, const int32_t* __restrict__ n_events, const double* __restrict__ event_times, const int32_t* __restrict__ modes,
                                       const double* __restrict__ tq, double* __restrict__ x_des, double* __restrict__ u_des, int32_t* __restrict__ mode_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31; const int b = blockIdx.x * (blockDim.x >> 5) + warp; if (b >= B) return;
  const int n = sol.n_nodes[b]; const double* gt = sol.t + (size_t)b * nmax; const double t = tq[b];
  int idx; double a; time_segment(gt, n, t, idx, a); const int i2 = (idx + 1 < n) ? idx + 1 : idx;
  const double* gx = sol.x + (size_t)b * nmax * NX; const double* gu = sol.u + (size_t)b * nmax * NU;
  if (lane < NX) { x_des[(size_t)b * NX + lane] = a * gx[(size_t)idx * NX + lane] + (1.0 - a) * gx[(size_t)i2 * NX + lane]; u_des[(size_t)b * NU + lane] = a * gu[(size_t)idx * NU + lane] + (1.0 - a) * gu[(size_t)i2 * NU + lane]; }
  if (lane == 0) mode_out[b] = mode_at_time(event_times + (size_t)b * EMAX, modes + (size_t)b * (EMAX + 1), n_events[b], t);
}

// =====================================================================================================
bool mpc_alloc(MpcBuffers& m, int B, int nmax, std::string& err, std::vector<void*>& allocs) {
  m.B = B; m.nmax = nmax; m.cur = 0;
  auto A = [&](auto** p, size_t count) { void* q = nullptr; const size_t bytes = count * sizeof(**p); cudaError_t e = cudaMalloc(&q, bytes); if (e != cudaSuccess) { err = std::string("cudaMalloc (MPC buffers) failed: ") + cudaGetErrorString(e); return false; } cudaMemset(q, 0, bytes); allocs.push_back(q); *p = static_cast<std::remove_reference_t<decltype(**p)>*>(q); return true; };
  const size_t Bn = (size_t)B * nmax;
  bool ok = A(&m.t0, B) && A(&m.x0, (size_t)B * NX) && A(&m.n_events, B) && A(&m.event_times, (size_t)B * EMAX) && A(&m.modes, (size_t)B * (EMAX + 1)) && A(&m.n_target, B) && A(&m.target_times, (size_t)B * KMAX) && A(&m.target_states, (size_t)B * KMAX * TARGET_DIM);
  for (int s = 0; s < 2 && ok; ++s) ok = A(&m.sol[s].n_nodes, B) && A(&m.sol[s].t, Bn) && A(&m.sol[s].event, Bn) && A(&m.sol[s].x, Bn * NX) && A(&m.sol[s].u, Bn * NU);
  ok = ok && A(&m.stage, Bn * STAGE_DBL) && A(&m.stage_i, Bn * STAGE_INT) && A(&m.gains, Bn * GAIN_DBL) && A(&m.dx, Bn * NX) && A(&m.du, Bn * NU) && A(&m.robot, (size_t)B * ROBOT_DBL) && A(&m.status, B) && A(&m.step_info, (size_t)B * 4);
  return ok;
}

int mpc_solve_launch(const DevModel* mdl, const DevModel& hm, MpcBuffers& m, const MpcProblemDev& p, cudaStream_t stream, cudaEvent_t* ev) {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(mpc_lq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(LqSmem) * LQ_WARPS));
    cudaFuncSetAttribute(mpc_riccati_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(RicSmem) * RIC_WARPS));
    cudaFuncSetAttribute(mpc_linesearch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(LsSmem) * LS_WARPS));
    configured = true;
  }
  (void)hm;
  const int B = m.B, nmax = m.nmax; MpcSolutionDev prev = m.sol[m.cur], next = m.sol[1 - m.cur];
  if (ev) cudaEventRecord(ev[0], stream);
  mpc_setup_kernel<<<(B + SETUP_WARPS - 1) / SETUP_WARPS, 32 * SETUP_WARPS, 0, stream>>>(mdl, B, nmax, p, prev, next, m.status);
  if (ev) cudaEventRecord(ev[1], stream);
  const long long nodes = (long long)B * nmax;
  mpc_lq_kernel<<<(unsigned)((nodes + LQ_WARPS - 1) / LQ_WARPS), 32 * LQ_WARPS, sizeof(LqSmem) * LQ_WARPS, stream>>>(mdl, B, nmax, p, next, m.stage, m.stage_i, m.status);
  if (ev) cudaEventRecord(ev[2], stream);
  mpc_riccati_kernel<<<(B + RIC_WARPS - 1) / RIC_WARPS, 32 * RIC_WARPS, sizeof(RicSmem) * RIC_WARPS, stream>>>(mdl, B, nmax, p, next, m.stage, m.stage_i, m.gains, m.dx, m.du, m.robot, m.status);
  if (ev) cudaEventRecord(ev[3], stream);
  mpc_linesearch_kernel<<<B, 32 * LS_WARPS, sizeof(LsSmem) * LS_WARPS, stream>>>(mdl, B, nmax, p, next, m.dx, m.du, m.robot, m.status, m.step_info);
  if (ev) cudaEventRecord(ev[4], stream);
  m.cur = 1 - m.cur;
  return 4;
}

int mpc_policy_eval_launch(const MpcBuffers& m, const double* t, double* x_des, double* u_des, int32_t* mode, cudaStream_t stream) {
  mpc_policy_eval_kernel<<<(m.B + 3) / 4, 128, 0, stream>>>(m.B, m.nmax, m.sol[m.cur], m.n_events, m.event_times, m.modes, t, x_des, u_des, mode);
  return 1;
}
int mpc_fixup_launch(const MpcBuffers& m, cudaStream_t stream) { mpc_fixup_kernel<<<m.B, 32, 0, stream>>>(m.B, m.nmax, m.sol[m.cur]); return 1; }

// ---- fp64 FMA throughput probe: 8 independent chains per thread, enough CTAs to fill every SM ----
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) { x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b); x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b); }
  if (x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 == 12345.678) out[0] = x0;
}
double measure_fp64_peak(cudaStream_t stream) {
  double* d = nullptr; cudaMalloc(&d, 8); cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int dev = 0, sms = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int blocks = sms * 8, iters = 1 << 15; double best = 0.0;
  fp64_peak_kernel<<<blocks, 256, 0, stream>>>(d, 1024, 0.999999, 1e-9);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0, stream); fp64_peak_kernel<<<blocks, 256, 0, stream>>>(d, iters, 0.999999, 1e-9); cudaEventRecord(e1, stream); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); const double tf = 2.0 * 8.0 * iters * 256.0 * blocks / (ms * 1e-3) / 1e12; if (tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d); return best;
}

}  // namespace qmb
