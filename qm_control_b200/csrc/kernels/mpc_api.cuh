// Host-visible interface of the MPC kernels: device buffers of one handle and the launchers.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "dev_common.cuh"

namespace qmb {

constexpr int EMAX = 32, KMAX = 4, TARGET_DIM = 37;
// per-node projected LQ stage as the LQ kernel hands it to the Riccati kernel (doubles)
constexpr int ST_A = 0, ST_B = 900, ST_b = 1440, ST_Q = 1470, ST_R = 2370, ST_S = 2694, ST_q = 3234, ST_r = 3264, ST_PXD = 3282, ST_PUD = 3762, ST_PED = 4050, ST_PERF = 4066, STAGE_DBL = 4072;
// ints: [0] type (0 intermediate, 1 event, 2 terminal), [1] m (projected input dim), [2] ndep, [3..19) dep input index, [19..37) free input index
constexpr int SI_TYPE = 0, SI_M = 1, SI_NDEP = 2, SI_DEP = 3, SI_FREE = 19, STAGE_INT = 40;
constexpr int GAIN_DBL = 18 * 30 + 18 + 6;   // feedback K (m x 30), feedforward k (m)
constexpr int ROBOT_DBL = 8;                 // armijo, base cost, base dyn SSE, base eq SSE, |dx|, |du|

// PrimalSolution of every robot: node count, node times, event annotation (0 none, 1 pre-event, 2 post-event), x, u
struct MpcSolutionDev { int32_t* n_nodes = nullptr; double* t = nullptr; int32_t* event = nullptr; double* x = nullptr; double* u = nullptr; };

struct MpcBuffers {
  int B = 0, nmax = 0, cur = 0;
  // inputs of the last solve (kept for policy evaluation: mode schedule)
  double *t0 = nullptr, *x0 = nullptr, *event_times = nullptr, *target_times = nullptr, *target_states = nullptr;
  int32_t *n_events = nullptr, *modes = nullptr, *n_target = nullptr;
  MpcSolutionDev sol[2];
  double *stage = nullptr, *gains = nullptr, *dx = nullptr, *du = nullptr, *robot = nullptr, *step_info = nullptr;
  int32_t *stage_i = nullptr, *status = nullptr;
};
bool mpc_alloc(MpcBuffers& m, int B, int nmax, std::string& err, std::vector<void*>& allocs, cudaStream_t stream);
int mpc_configure_device();   // per-device opt-in shared memory of the MPC kernels (qmb200_create, after cudaSetDevice)

struct MpcProblemDev { const double* t0; const double* x0; const int32_t* n_events; const double* event_times; const int32_t* modes; const int32_t* n_target; const double* target_times; const double* target_states; };

// One SQP iteration for robots [b0, b1) (4 kernels on `stream`): reads m.sol[m.cur], writes m.sol[1 - m.cur]; the caller
// flips m.cur after queueing every range.  Returns the number of kernels launched.
// `ev` (optional, 5 events): recorded before K1 and after each of K1..K4 for per-kernel timing.
int mpc_solve_launch(const DevModel* mdl, const DevModel& host_mdl, MpcBuffers& m, const MpcProblemDev& p, int b0, int b1, cudaStream_t stream, cudaEvent_t* ev = nullptr);
// fp64 FMA throughput microbenchmark (roofline denominator for the compute-bound kernels); returns TFLOP/s
double measure_fp64_peak(cudaStream_t stream);
// evaluatePolicy on m.sol[m.cur]; returns kernels launched
int mpc_policy_eval_launch(const MpcBuffers& m, const double* t, double* x_des, double* u_des, int32_t* mode, cudaStream_t stream, int b0 = 0, int b1 = -1);
// input fix-up after loading a solution from the host (inputs at pre-event / last nodes)
int mpc_fixup_launch(const MpcBuffers& m, cudaStream_t stream);

}  // namespace qmb
