"""Tiny end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): every kernel of the library once, 5 robots (mixed gaits)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qm_control_b200 as q  # noqa: E402
from qm_control_b200 import synthetic  # noqa: E402

B = 5
ctrl = q.QMController(batch=B, dt=0.015); s = ctrl.solver
prob, wbc = synthetic.make_batch(np.arange(B), config=5)
ctrl.starting(wbc["rbd"], time=12.0)
for tick in range(2):
    p = dict(prob); p["t0"] = ctrl.t_obs.copy(); p["x0"] = ctrl.x_obs.copy()
    nt, tt, ts = ctrl.targetTrajectories(0, np.tile([0.2, 0.0, 0.0, 0.1], (B, 1)))
    cmd, status = s.tick(p, p["t0"] + 0.002, wbc["rbd"], wbc["period"])
    cmd2, status2 = ctrl.update(wbc["rbd"], 0.002)
    eff, st = s.hw_write(ctrl.t_obs, np.full(B, 0.002), ctrl.joint_cmd, wbc["rbd"][:, 6:24], wbc["rbd"][:, 30:48])
# solver variants (IPM thresholds; DDP: rollout kernels) and the multi-GPU pack path on one rank
import torch  # noqa: E402
for name in ("ipm", "ddp", "sqp"):
    s.mpc_set_solver(name); out = s.mpc_solve(prob)
dev = torch.device("cuda", 0); cmd_d = torch.from_numpy(cmd).to(dev); all_d = torch.zeros((B, 18), dtype=torch.float64, device=dev)
perm = s.gait_bin_permutation(prob); s.allgather_torque(cmd_d, all_d, torch.from_numpy(perm).to(dev)); torch.cuda.synchronize()
s1 = q.Solver(batch=B, dt=0.015, wbc_variant=1); c1, st1 = s1.tick(prob, prob["t0"] + 0.002, wbc["rbd"], wbc["period"])
print("sanitize_small ok", status, status2, out["status"], st1)
