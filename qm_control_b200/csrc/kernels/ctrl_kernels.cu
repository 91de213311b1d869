// Batched kernels for the steps either side of the MPC+WBC path:
//   ctrl_observation_kernel   QMController::updateStateEstimation tail (qm_controllers/src/QMController.cpp:236-243):
//                             currentObservation_.time += period; state = computeCentroidalStateFromRbdModel(rbd); yaw unwrap
//   ctrl_target_kernel        cmdVelToTargetTrajectories / EeCmdVelToTargetTrajectories / EEgoalPoseToTargetTrajectories
//                             (qm_controllers/src/QmTargetTrajectoriesPublisher_node.cpp:44-208) incl. the lastEeTarget_ bookkeeping
//                             (QmTargetTrajectoriesPublisher.h:55-57, QmTargetTrajectoriesPublisher.cpp:107-108)
//   ctrl_control_law_kernel   SafetyChecker::check (SafetyChecker.h:22-35) + QMController::updateControlLaw (QMController.cpp:177-190)
//                             / QMMpcController::updateControlLaw (QMController.cpp:427-445)
//   ctrl_hw_write_kernel      QMHWSim::writeSim (qm_gazebo/src/QMHWSim.cpp:98-116): command delay buffer + hybrid joint law
//                             tau = kp (posDes - q) + kd (velDes - qd) + ff (HybridJointInterface.h:55-61)
// All of them are maps over robots with O(100) flops and O(1 KB) of I/O per robot: HBM/latency bound.  One thread owns one robot
// (or one joint of a robot); rows are staged through shared memory so that every global access is a coalesced row-major sweep.
#include "ctrl_api.cuh"

namespace qmb {

namespace {
constexpr int OBS_ROBOTS = 64;    // robots per CTA of the row-staged kernels
constexpr double PI = 3.14159265358979323846;

// coalesced copy of `rows` consecutive rows of width W between global memory and a shared tile with leading dimension LD (odd → the
// thread-per-row accesses that follow are bank-conflict free)
template <int W, int LD> __device__ __forceinline__ void tile_load(double* tile, const double* __restrict__ g, int rows) {
  for (int i = threadIdx.x; i < rows * W; i += blockDim.x) tile[(i / W) * LD + (i % W)] = g[i];
}
template <int W, int LD> __device__ __forceinline__ void tile_store(double* __restrict__ g, const double* tile, int rows) {
  for (int i = threadIdx.x; i < rows * W; i += blockDim.x) g[i] = tile[(i / W) * LD + (i % W)];
}

// angles::shortest_angular_distance(from, to) = normalize_angle(to - from), normalize_angle(a) = fmod(fmod(a, 2pi) + 2pi, 2pi), shifted into (-pi, pi]
__device__ __forceinline__ double shortest_angular_distance(double from, double to) {
  double a = fmod(fmod(to - from, 2.0 * PI) + 2.0 * PI, 2.0 * PI);
  if (a > PI) a -= 2.0 * PI;
  return a;
}

// Eigen::Quaterniond(w, x, y, z).toRotationMatrix()
__device__ __forceinline__ void quat_to_rot(double w, double x, double y, double z, double* R) {
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z, twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
}
}  // namespace

// -----------------------------------------------------------------------------------------------------------------
// Observation: rbd[55] → centroidal state (SRBD mapping, the same arithmetic as qmb200_centroidal_state_from_rbd:
// CentroidalModelRbdConversions::computeCentroidalStateFromRbdModel [upstream, recalled]) with the controller's yaw unwrap.
__global__ void __launch_bounds__(OBS_ROBOTS) ctrl_observation_kernel(const DevModel* __restrict__ mdl, int B, const double* __restrict__ rbd, const double* __restrict__ period,
                                                                       double* __restrict__ t_obs, double* __restrict__ x_obs) {
  __shared__ double s_rbd[OBS_ROBOTS * 55];   // leading dimension 55 (odd)
  __shared__ double s_x[OBS_ROBOTS * 31];
  const int b0 = blockIdx.x * OBS_ROBOTS, rows = min(OBS_ROBOTS, B - b0), r = threadIdx.x;
  tile_load<55, 55>(s_rbd, rbd + (size_t)b0 * 55, rows);
  __syncthreads();
  if (r < rows) {
    const double* s = s_rbd + r * 55; double* o = s_x + r * 31;
    double R[9]; rot_zyx(s[0], s[1], s[2], R);
    const double w[3] = {s[NQ], s[NQ + 1], s[NQ + 2]};
    double c[3]; matvec3(R, mdl->c_nom, c);
    double cw[3]; cross3(c, w, cw);                                     // h_lin/m = v_lin + (R c_nom) x w
    double Rtw[3], IRtw[3], L[3]; matTvec3(R, w, Rtw); matvec3(mdl->I_nom, Rtw, IRtw); matvec3(R, IRtw, L);   // h_ang = R I R^T w
    const double inv_m = 1.0 / mdl->total_mass;
#pragma unroll
    for (int i = 0; i < 3; ++i) { o[i] = s[NQ + 3 + i] + cw[i]; o[3 + i] = L[i] * inv_m; o[6 + i] = s[3 + i]; o[9 + i] = s[i]; }
#pragma unroll
    for (int j = 0; j < NJ; ++j) o[12 + j] = s[6 + j];
    const double yaw_last = x_obs[(size_t)(b0 + r) * NX + 9];           // currentObservation_.state(9) of the previous update
    o[9] = yaw_last + shortest_angular_distance(yaw_last, o[9]);
    t_obs[b0 + r] += period[b0 + r];
  }
  __syncthreads();
  tile_store<NX, 31>(x_obs + (size_t)b0 * NX, s_x, rows);
}

// -----------------------------------------------------------------------------------------------------------------
// Target front-end.  kind 0: /cmd_vel (cmd = vx, vy, vz, yaw rate), 1: /ee_cmd_vel (cmd = vx, vy, vz), 2: goal pose (cmd = pos(3), quat xyzw(4)).
// Output: the 2-knot TargetTrajectories [time; 37-dim state = (0_6 | v, base pose, defaultJointState, EE pose)] in the solver's layout.
__global__ void __launch_bounds__(128) ctrl_target_kernel(TargetParams prm, int kind, int B, const double* __restrict__ cmd /*[B][7]*/, const double* __restrict__ t_obs,
                                                           const double* __restrict__ x_obs, const double* __restrict__ ee_state /*[B][7]*/, double* __restrict__ last_ee_target /*[B][7]*/,
                                                           int32_t* __restrict__ n_target, double* __restrict__ target_times /*[B][KMAX]*/, double* __restrict__ target_states /*[B][KMAX][37]*/) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x; if (b >= B) return;
  const double* c = cmd + (size_t)b * 7; const double* x = x_obs + (size_t)b * NX; const double* ee = ee_state + (size_t)b * 7; double* le = last_ee_target + (size_t)b * 7;
  const double t = t_obs[b];
  double base_cur[6]; for (int i = 0; i < 6; ++i) base_cur[i] = x[6 + i];
  double base_tgt[6], ee_cur[7], ee_tgt[7], vel[3] = {0.0, 0.0, 0.0}, t_reach;
  if (kind == 0) {            // cmdVelToTargetTrajectories (:73-113)
    double R[9]; rot_zyx(base_cur[3], base_cur[4], base_cur[5], R); matvec3(R, c, vel);
    base_tgt[0] = base_cur[0] + vel[0] * prm.time_to_target; base_tgt[1] = base_cur[1] + vel[1] * prm.time_to_target; base_tgt[2] = prm.com_height;
    base_tgt[3] = base_cur[3] + c[3] * prm.time_to_target; base_tgt[4] = 0.0; base_tgt[5] = 0.0;
    const double d0 = le[0] - ee[0], d1 = le[1] - ee[1], d2 = le[2] - ee[2];
    if (sqrt(d0 * d0 + d1 * d1 + d2 * d2) > 0.1) { le[0] = ee[0]; le[1] = ee[1]; le[2] = ee[2]; }
    for (int i = 0; i < 7; ++i) { ee_tgt[i] = le[i]; ee_cur[i] = le[i]; }   // eeStateLast.state = EeTargetPose (:104-105)
    t_reach = t + prm.time_to_target;
  } else if (kind == 1) {     // EeCmdVelToTargetTrajectories (:118-165)
    double Rq[9], Ri[9], M[9]; quat_to_rot(ee[6], ee[3], ee[4], ee[5], Rq); quat_to_rot(-0.5, 0.5, -0.5, 0.5, Ri); matmul3_nt(Rq, Ri, M);
    double v[3]; matvec3(M, c, v);
    for (int i = 0; i < 7; ++i) ee_cur[i] = ee[i];
    ee_tgt[0] = ee[0] + v[0] * prm.time_to_target; ee_tgt[1] = ee[1] + v[1] * prm.time_to_target; for (int i = 2; i < 7; ++i) ee_tgt[i] = le[i];
    for (int i = 0; i < 6; ++i) base_tgt[i] = base_cur[i];
    base_tgt[0] = ee_tgt[0] - 0.52; base_tgt[1] = ee_tgt[1] - 0.09; base_tgt[2] = prm.com_height; base_tgt[4] = 0.0; base_tgt[5] = 0.0;
    t_reach = t + prm.time_to_target;
  } else {                    // EEgoalPoseToTargetTrajectories (:172-208) + processFeedback's lastEeTarget_ update
    for (int i = 0; i < 7; ++i) { ee_cur[i] = ee[i]; ee_tgt[i] = c[i]; }
    for (int i = 0; i < 6; ++i) base_tgt[i] = base_cur[i];
    base_tgt[0] = c[0] - 0.52; base_tgt[1] = c[1] - 0.09; base_tgt[2] = prm.com_height; base_tgt[4] = 0.0; base_tgt[5] = 0.0;
    // quaternionDistance(q_current, q_target) = w_c v_t - w_t v_c + v_c x v_t [upstream ocs2_robotic_tools, recalled]
    const double wc = ee[6], wt = c[6]; const double vc[3] = {ee[3], ee[4], ee[5]}, vt[3] = {c[3], c[4], c[5]}; double cr[3]; cross3(vc, vt, cr);
    double dl = 0.0, dr = 0.0;
    for (int i = 0; i < 3; ++i) { const double dp = c[i] - ee[i], dq = wc * vt[i] - wt * vc[i] + cr[i]; dl += dp * dp; dr += dq * dq; }
    t_reach = t + fmax(sqrt(dr) / prm.target_rotation_velocity, sqrt(dl) / prm.target_displacement_velocity);   // estimateTimeToTarget (:24-41)
    for (int i = 0; i < 7; ++i) le[i] = c[i];
  }
  base_cur[2] = prm.com_height; base_cur[4] = 0.0; base_cur[5] = 0.0;   // targetPoseToTargetTrajectories (:44-68)
  double* tt = target_times + (size_t)b * KMAX; double* ts = target_states + (size_t)b * KMAX * TARGET_DIM;
  n_target[b] = 2; tt[0] = t; tt[1] = t_reach; for (int k = 2; k < KMAX; ++k) tt[k] = 0.0;
  for (int k = 0; k < 2; ++k) {
    double* s = ts + k * TARGET_DIM;
    for (int i = 0; i < 3; ++i) { s[i] = vel[i]; s[3 + i] = 0.0; }
    for (int i = 0; i < 6; ++i) s[6 + i] = k == 0 ? base_cur[i] : base_tgt[i];
    for (int j = 0; j < NJ; ++j) s[12 + j] = prm.default_joint_state[j];
    for (int i = 0; i < 7; ++i) s[30 + i] = k == 0 ? ee_cur[i] : ee_tgt[i];
  }
  for (int i = 2 * TARGET_DIM; i < KMAX * TARGET_DIM; ++i) ts[i] = 0.0;
}

// -----------------------------------------------------------------------------------------------------------------
// Control law: one thread per (robot, joint).  joint_cmd[b][j] = (posDes, velDes, kp, kd, ff) as HybridJointHandle::setCommand receives them.
// variant 0 (QMController): legs only once time > 10 (before that the handle keeps its previous command: the entry is left untouched);
//   arm joints (posDes, 0, arm_kp, arm_kd, torque).
// variant 1 (QMMpcController): legs always; the arm is position controlled at 100 Hz: arm_pos_cmd[b][j] = state(24+j) + velDes(12+j)/100
//   whenever time - last_time > 1/100 (last_time is then advanced); its hybrid entries are left untouched.
// status: bit 0 = SafetyChecker orientation check failed (|roll| > pi/2 → the reference calls stopRequest).
__global__ void __launch_bounds__(ControlLawParams::THREADS) ctrl_control_law_kernel(ControlLawParams prm, int B, const double* __restrict__ x_des, const double* __restrict__ u_des, const double* __restrict__ wbc_cmd,
                                                                                      const double* __restrict__ t_obs, const double* __restrict__ x_obs, double* __restrict__ joint_cmd /*[B][18][5]*/,
                                                                                      double* __restrict__ arm_pos_cmd /*[B][6]*/, double* __restrict__ last_time /*[B]*/, int32_t* __restrict__ status) {
  const int b = blockIdx.x * ControlLawParams::ROBOTS + threadIdx.x / NJ, j = threadIdx.x % NJ;
  const bool live = threadIdx.x < ControlLawParams::ROBOTS * NJ && b < B;
  double t = 0.0, lt = 0.0;
  if (live) {
    t = t_obs[b];
    const double pos_des = x_des[(size_t)b * NX + 12 + j], vel_des = u_des[(size_t)b * NU + 12 + j], tau = wbc_cmd[(size_t)b * 54 + 36 + j];
    double* jc = joint_cmd + ((size_t)b * NJ + j) * 5;
    if (j < 12) {
      if (prm.variant == 1 || t > 10.0) { jc[0] = pos_des; jc[1] = vel_des; jc[2] = 0.0; jc[3] = 3.0; jc[4] = tau; }
    } else if (prm.variant == 0) {
      jc[0] = pos_des; jc[1] = 0.0; jc[2] = prm.arm_kp; jc[3] = prm.arm_kd; jc[4] = tau;
    } else {
      lt = last_time[b];
      if (t - lt > 1.0 / 100.0) arm_pos_cmd[(size_t)b * 6 + j - 12] = x_obs[(size_t)b * NX + 12 + j] + vel_des * 1.0 / 100.0;
    }
    if (j == 0) { const double roll = x_obs[(size_t)b * NX + 11]; status[b] = (roll > 0.5 * PI || roll < -0.5 * PI) ? 1 : 0; }
  }
  __syncthreads();   // every arm thread has read last_time before it moves
  if (live && prm.variant == 1 && j == 12 && t - lt > 1.0 / 100.0) last_time[b] = t;
}

// -----------------------------------------------------------------------------------------------------------------
// QMHWSim::writeSim: per robot a FIFO of stamped commands (ring of HW_DEPTH entries, oldest at `tail`); the command applied is the oldest
// one that is not older than `delay`.  One thread per (robot, joint); all joints of a robot share the stamps.
__global__ void __launch_bounds__(ControlLawParams::THREADS) ctrl_hw_write_kernel(int B, double delay, const double* __restrict__ time, const double* __restrict__ period,
                                                                                   const double* __restrict__ joint_cmd /*[B][18][5]*/, const double* __restrict__ joint_pos /*[B][18]*/,
                                                                                   const double* __restrict__ joint_vel, double* __restrict__ ring_cmd /*[B][HW_DEPTH][18][5]*/,
                                                                                   double* __restrict__ ring_stamp /*[B][HW_DEPTH]*/, int32_t* __restrict__ ring_state /*[B][2] tail, count*/,
                                                                                   double* __restrict__ effort /*[B][18]*/, int32_t* __restrict__ status) {
  const int b = blockIdx.x * ControlLawParams::ROBOTS + threadIdx.x / NJ, j = threadIdx.x % NJ;
  const bool live = threadIdx.x < ControlLawParams::ROBOTS * NJ && b < B;
  int tail = 0, count = 0, st = 0; double t = 0.0;
  if (live) {
    t = time[b]; tail = ring_state[2 * b]; count = ring_state[2 * b + 1];
    const double* stamp = ring_stamp + (size_t)b * HW_DEPTH;
    if (t == period[b]) count = 0;                                                             // simulation reset (:101-103)
    while (count > 0 && stamp[tail] + delay < t) { tail = (tail + 1) % HW_DEPTH; --count; }    // pop_back of expired commands (:105-107)
    if (count == HW_DEPTH) { tail = (tail + 1) % HW_DEPTH; --count; st = 2; }                  // ring full (the reference deque is unbounded): drop the oldest, flag it
    const int head = (tail + count) % HW_DEPTH;
    const double* jc = joint_cmd + ((size_t)b * NJ + j) * 5; double* slot = ring_cmd + (((size_t)b * HW_DEPTH + head) * NJ + j) * 5;
    double c5[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) { c5[i] = jc[i]; slot[i] = c5[i]; }                            // push_front (:108-109)
    if (count > 0) { const double* old = ring_cmd + (((size_t)b * HW_DEPTH + tail) * NJ + j) * 5;
#pragma unroll
      for (int i = 0; i < 5; ++i) c5[i] = old[i]; }                                            // buffer.back() (:111)
    effort[(size_t)b * NJ + j] = c5[2] * (c5[0] - joint_pos[(size_t)b * NJ + j]) + c5[3] * (c5[1] - joint_vel[(size_t)b * NJ + j]) + c5[4];
  }
  __syncthreads();   // every joint thread has read the stamps / ring state of its robot
  if (live && j == 0) { ring_stamp[(size_t)b * HW_DEPTH + (tail + count) % HW_DEPTH] = t; ring_state[2 * b] = tail; ring_state[2 * b + 1] = count + 1; status[b] = st; }
}

// -----------------------------------------------------------------------------------------------------------------
int launch_observation(const DevModel* mdl, int B, const double* rbd, const double* period, double* t_obs, double* x_obs, cudaStream_t s) {
  ctrl_observation_kernel<<<(B + OBS_ROBOTS - 1) / OBS_ROBOTS, OBS_ROBOTS, 0, s>>>(mdl, B, rbd, period, t_obs, x_obs); return 1;
}
int launch_target(const TargetParams& prm, int kind, int B, const double* cmd, const double* t_obs, const double* x_obs, const double* ee_state, double* last_ee_target,
                  int32_t* n_target, double* target_times, double* target_states, cudaStream_t s) {
  ctrl_target_kernel<<<(B + 127) / 128, 128, 0, s>>>(prm, kind, B, cmd, t_obs, x_obs, ee_state, last_ee_target, n_target, target_times, target_states); return 1;
}
int launch_control_law(const ControlLawParams& prm, int B, const double* x_des, const double* u_des, const double* wbc_cmd, const double* t_obs, const double* x_obs,
                       double* joint_cmd, double* arm_pos_cmd, double* last_time, int32_t* status, cudaStream_t s) {
  ctrl_control_law_kernel<<<(B + ControlLawParams::ROBOTS - 1) / ControlLawParams::ROBOTS, ControlLawParams::THREADS, 0, s>>>(prm, B, x_des, u_des, wbc_cmd, t_obs, x_obs, joint_cmd, arm_pos_cmd, last_time, status); return 1;
}
int launch_hw_write(int B, double delay, const double* time, const double* period, const double* joint_cmd, const double* joint_pos, const double* joint_vel,
                    double* ring_cmd, double* ring_stamp, int32_t* ring_state, double* effort, int32_t* status, cudaStream_t s) {
  ctrl_hw_write_kernel<<<(B + ControlLawParams::ROBOTS - 1) / ControlLawParams::ROBOTS, ControlLawParams::THREADS, 0, s>>>(B, delay, time, period, joint_cmd, joint_pos, joint_vel, ring_cmd, ring_stamp, ring_state, effort, status); return 1;
}

}  // namespace qmb
