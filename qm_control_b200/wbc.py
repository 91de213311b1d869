"""Python mirror of the WBC seam: qm::WbcBase::update / HierarchicalWbc / HierarchicalMpcWbc
(qm_wbc/include/qm_wbc/WbcBase.h:28-34, src/HierarchicalWbc.cpp:18-44, src/HierarchicalMpcWbc.cpp:18-34), batched."""
import numpy as np

from .interface import QMInterface, Solver


class _WbcBase:
    VARIANT = 0

    def __init__(self, interface=None, batch=1, device=0, solver=None):
        self.solver = solver or Solver(interface or QMInterface(), batch=batch, device=device, wbc_variant=self.VARIANT)
        self.batch = self.solver.batch

    def loadTasksSetting(self, taskFile=None, verbose=False):
        """Torque limits come from the URDF and the friction coefficient from task.info at construction (WbcBase.cpp:565-596)."""
        return None

    def update(self, stateDesired, inputDesired, rbdStateMeasured, mode, period, time):
        """→ ([B,54] = [vdot(24), F(12), tau(18)], status[B]).  Scalars broadcast over the batch."""
        B = self.batch
        x = np.asarray(stateDesired, dtype=np.float64).reshape(B, 30); u = np.asarray(inputDesired, dtype=np.float64).reshape(B, 30)
        rbd = np.asarray(rbdStateMeasured, dtype=np.float64).reshape(B, 55)
        mode = np.broadcast_to(np.asarray(mode, dtype=np.int32), (B,)); period = np.broadcast_to(np.asarray(period, dtype=np.float64), (B,)); time = np.broadcast_to(np.asarray(time, dtype=np.float64), (B,))
        return self.solver.wbc_update(x, u, rbd, mode, period, time)


class HierarchicalWbc(_WbcBase):
    VARIANT = 0


class HierarchicalMpcWbc(_WbcBase):
    VARIANT = 1
