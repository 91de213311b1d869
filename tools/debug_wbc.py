import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
import qm_control_b200 as q
from qm_control_b200 import synthetic
from _oracle import Oracle
o = Oracle(); B = 192; variant = int(sys.argv[1]) if len(sys.argv) > 1 else 1
solver = q.Solver(batch=B, wbc_variant=variant)
prob, wbc = synthetic.make_batch(np.arange(B), config=5)
x_des, u_des, mode = synthetic.nominal_wbc_inputs(prob, solver.robot_mass)
u_des = u_des + synthetic.uniform(77, np.arange(B), 1, 30, -1.0, 1.0) * np.r_[np.full(12, 5.0), np.full(18, 0.2)]
for b in range(B):
    for f in range(4):
        if not (mode[b] >> (3 - f)) & 1: u_des[b, 3*f:3*f+3] = 0.0
il = u_des + synthetic.uniform(78, np.arange(B), 2, 30, -0.002, 0.002); tarr = np.full(B, 12.0)
solver.wbc_set_input_last(il)
cmd, status = solver.wbc_update(x_des, u_des, wbc["rbd"], mode, wbc["period"], tarr)
ref, _ = o.wbc_update_batch(x_des, u_des, wbc["rbd"], mode, wbc["period"], tarr, il, variant=variant, nthreads=8)
err = np.max(np.abs(cmd - ref), axis=1) / np.maximum(1, np.max(np.abs(ref), axis=1))
errd = np.max(np.abs(cmd[:, np.r_[0:18, 24:36]] - ref[:, np.r_[0:18, 24:36]]), axis=1) / np.maximum(1, np.max(np.abs(ref), axis=1))
for b in range(B):
    if status[b] or err[b] > 1e-5:
        st = int(status[b])
        try: _, _, it = o.wbc_update(x_des[b], u_des[b], wbc["rbd"][b], int(mode[b]), 0.002, 12.0, input_last=il[b], variant=variant)
        except Exception as e: it = str(e)
        print("robot %3d mode %2d status %d it1 %d it2 %d nw %d | err %.2e err_determined %.2e | oracle iters %s" % (b, mode[b], st & 255, (st >> 8) & 255, (st >> 16) & 255, (st >> 24) & 255, err[b], errd[b], it))
print("max err all", err.max(), "max err determined", errd.max(), "n bad status", int((status != 0).sum()))
