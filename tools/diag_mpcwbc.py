#!/usr/bin/env python3
"""HierarchicalMpcWbc: who is closer to the exact cascade optimum, the CUDA path or the oracle?  For a sample of robots: hierarchy objectives (level-1 and
level-2 residual norms), feasibility, KKT certificates (NNLS) and the distance of each result to the exact re-solve on its own active set (numpy SVD)."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import qm_control_b200 as q
from qm_control_b200 import synthetic
from _oracle import Oracle
import test_wbc_twin_cpu as tw


def lse(A, b, E, e):
    U, s, Vt = np.linalg.svd(E, full_matrices=True); k = int((s > s.max() * 1e-12).sum())
    xp = Vt[:k].T @ ((U[:, :k].T @ e) / s[:k]); N = Vt[k:].T
    if N.shape[1] == 0:
        return xp
    y = np.linalg.lstsq(A @ N, b - A @ xp, rcond=1e-13)[0]
    return xp + N @ y


o = Oracle(); g = tw._gains(); B = 96; ids = np.arange(B); solver = q.Solver(batch=B, wbc_variant=1)
prob, wbc = synthetic.make_batch(ids, config=3); x_des, u_des, mode = synthetic.nominal_wbc_inputs(prob, solver.robot_mass)
u_des = u_des + synthetic.uniform(77, ids, 1, 30, -1.0, 1.0) * np.r_[np.full(12, 5.0), np.full(18, 0.2)]
il = u_des + synthetic.uniform(78, ids, 2, 30, -0.002, 0.002); tarr = np.full(B, 12.0)
solver.wbc_set_input_last(il); cmd, status = solver.wbc_update(x_des, u_des, wbc["rbd"], mode, wbc["period"], tarr); diag = solver.wbc_get_diagnostics()
ref, _ = o.wbc_update_batch(x_des, u_des, wbc["rbd"], mode, wbc["period"], tarr, il, variant=1, nthreads=8)
rows = []
for b in range(0, B, 4):
    dbg = o.wbc_debug(x_des[b], u_des[b], wbc["rbd"][b], int(mode[b]), wbc["period"][b], 12.0, input_last=il[b], variant=1)
    (A0, b0, D0, f0), (A1, b1), (A2, b2), M = tw._tasks(o, dbg, u_des[b], int(mode[b]), 12.0, g)
    A1v = np.r_[A1[:4], A2[12:14]]; b1v = np.r_[b1[:4], b2[12:14]]; A2v = A2[:12]; b2v = b2[:12]
    rec = {"robot": b, "it": [int(diag[k][b]) for k in ("level0_passes", "level1_iterations", "level2_iterations", "working_set")]}
    for name, x in (("cuda", cmd[b, :36]), ("oracle", ref[b, :36])):
        viol = D0 @ x - f0; act = viol > -1e-7 * (1.0 + np.abs(f0))
        r1 = tw._certificate(A1v.T @ (A1v @ x - b1v), A0, D0[act]); r2 = tw._certificate(A2v.T @ (A2v @ x - b2v), np.r_[A0, A1v], D0[act])
        E = np.r_[A0, A1v, D0[act]]; e = np.r_[b0, A1v @ x, f0[act]]; xs = lse(A2v, b2v, E, e)
        rec[name] = dict(eq0=float(np.max(np.abs(A0 @ x - b0))), viol=float(viol.max()), nact=int(act.sum()), obj1=float(np.linalg.norm(A1v @ x - b1v)), obj2=float(np.linalg.norm(A2v @ x - b2v)),
                         kkt1=float(r1), kkt2=float(r2), dist_exact=float(np.max(np.abs(xs - x) / np.maximum(1.0, np.abs(xs)))), act=np.nonzero(act)[0].tolist())
    rec["same_active_set"] = rec["cuda"]["act"] == rec["oracle"]["act"]
    rec["dx_leg_torque"] = float(np.max(np.abs(cmd[b, 36:48] - ref[b, 36:48]))); rec["dx_force"] = float(np.max(np.abs(cmd[b, 24:36] - ref[b, 24:36])))
    rows.append(rec)
    print(b, rec["it"], "same_act", rec["same_active_set"], "| cuda obj1 %.9e obj2 %.9e kkt %.1e %.1e viol %.1e dist %.1e | orc obj1 %.9e obj2 %.9e kkt %.1e %.1e viol %.1e dist %.1e | dF %.1e dtau_leg %.1e" % (
        rec["cuda"]["obj1"], rec["cuda"]["obj2"], rec["cuda"]["kkt1"], rec["cuda"]["kkt2"], rec["cuda"]["viol"], rec["cuda"]["dist_exact"],
        rec["oracle"]["obj1"], rec["oracle"]["obj2"], rec["oracle"]["kkt1"], rec["oracle"]["kkt2"], rec["oracle"]["viol"], rec["oracle"]["dist_exact"], rec["dx_force"], rec["dx_leg_torque"]))
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "diag_mpcwbc.json"), "w"))
