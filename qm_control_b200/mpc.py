"""Python mirror of the MPC seam: ocs2::SqpMpc as QMController::setupMpc builds it (qm_controllers/src/QMController.cpp:286-306):
one multiple-shooting SQP iteration per advanceMpc(), warm-started from the previous PrimalSolution."""
from .interface import QMInterface, Solver


class SqpMpc:
    def __init__(self, interface=None, batch=1, device=0, time_horizon=0.0, dt=0.0, solver=None):
        self.solver = solver or Solver(interface or QMInterface(), batch=batch, device=device, time_horizon=time_horizon, dt=dt)
        self.batch = self.solver.batch

    def reset(self):
        """MPC_BASE::reset — forget the previous solution (cold start through QMInitializer)."""
        self.solver.mpc_reset()

    def run(self, prob):
        """MPC_BASE::run(t0, x0) for every robot; prob as in Solver.mpc_solve. Returns the PrimalSolution arrays."""
        return self.solver.mpc_solve(prob)

    def evaluatePolicy(self, t):
        """MPC_MRT_Interface::evaluatePolicy → (optimizedState, optimizedInput, plannedMode)."""
        return self.solver.policy_eval(t)
