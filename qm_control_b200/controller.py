"""Python mirror of qm::QMController / qm::QMMpcController (qm_controllers/include/qm_controllers/QMController.h:37-118) for the
numerical part of the plugin, batched: the members the reference keeps per controller (currentObservation_, joint handle commands,
lastEeTarget_ of the target publisher, last_time_) are arrays over robots here; every method is one C-ABI call on the device."""
import numpy as np

from .interface import QMInterface, Solver


class QMController:
    WBC_VARIANT = 0   # HierarchicalWbc (QMController::setupWbc, QMController.cpp:336-340)

    def __init__(self, interface=None, batch=1, device=0, time_horizon=0.0, dt=0.0, solver=None):
        self.solver = solver or Solver(interface or QMInterface(), batch=batch, device=device, time_horizon=time_horizon, dt=dt, wbc_variant=self.WBC_VARIANT)
        B = self.batch = self.solver.batch
        self.t_obs = np.zeros(B); self.x_obs = np.tile(self.solver.initial_state, (B, 1))   # currentObservation_ (QMController.cpp:100-104)
        self.joint_cmd = np.zeros((B, 18, 5)); self.arm_pos_cmd = np.zeros((B, 6)); self.last_time = np.zeros(B)
        self.last_ee_target = self.solver.initial_ee_target()                               # QmTargetTrajectoriesPublisher.h:55-57
        self.measured_rbd = np.zeros((B, 55))

    def starting(self, rbd, time=0.0):
        """QMController::starting (QMController.cpp:98-126): first observation from the measured state, last_time_ = observation time."""
        self.measured_rbd = np.asarray(rbd, dtype=np.float64).reshape(self.batch, 55)
        self.t_obs = np.full(self.batch, float(time)); self.x_obs = self.solver.centroidal_state_from_rbd(self.measured_rbd)
        self.last_time = self.t_obs.copy()

    def updateStateEstimation(self, rbd, period):
        self.measured_rbd = np.asarray(rbd, dtype=np.float64).reshape(self.batch, 55)
        self.t_obs, self.x_obs = self.solver.observation_update(self.measured_rbd, np.broadcast_to(np.asarray(period, dtype=np.float64), (self.batch,)), self.t_obs, self.x_obs)

    def targetTrajectories(self, kind, cmd):
        """The publisher node's callbacks (QmTargetTrajectoriesPublisher.h:75-103, .cpp:94-109) on the latest observation / EE state."""
        nt, tt, ts, self.last_ee_target = self.solver.target_trajectories(kind, cmd, self.t_obs, self.x_obs, self.measured_rbd[:, 48:55], self.last_ee_target)
        return nt, tt, ts

    def advanceMpc(self, prob):
        """mpcMrtInterface_->advanceMpc() (QMController.cpp:315-332) with the observation as initial condition."""
        p = dict(prob); p["t0"] = self.t_obs.copy(); p["x0"] = self.x_obs.copy()
        return self.solver.mpc_solve(p)

    def update(self, rbd, period):
        """QMController::update (QMController.cpp:128-175) → (WBC 54-vector, status); joint handle commands land in self.joint_cmd."""
        self.measured_rbd = np.asarray(rbd, dtype=np.float64).reshape(self.batch, 55)
        per = np.broadcast_to(np.asarray(period, dtype=np.float64), (self.batch,))
        self.t_obs, self.x_obs, self.joint_cmd, self.arm_pos_cmd, self.last_time, cmd, status = self.solver.update(self.measured_rbd, per, self.t_obs, self.x_obs, self.joint_cmd, self.arm_pos_cmd, self.last_time)
        return cmd, status


class QMMpcController(QMController):
    WBC_VARIANT = 1   # HierarchicalMpcWbc + position-controlled arm (QMController.cpp:409-445)
