#!/usr/bin/env python3
"""Run the bench workload for a few ticks and report robots whose status word is non-zero (besides the no-step bit)."""
import sys, numpy as np
sys.path.insert(0, ".")
import qm_control_b200 as q
from qm_control_b200 import synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192; ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 12
solver = q.Solver(batch=B, dt=0.01, time_horizon=1.0)
prob, wbc = synthetic.make_batch(np.arange(B), config=4, horizon=1.0)
te = prob["t0"] + 0.002
for k in range(ticks):
    cmd, st = solver.tick(prob, te, wbc["rbd"], wbc["period"]); prob["t0"] = prob["t0"] + 0.01; te = te + 0.01
    bad = np.nonzero(st & ~(16 << 8))[0]
    sol = solver.mpc_get_solution()
    print("tick", k, "bad", bad.tolist(), [hex(int(st[b])) for b in bad], "nostep", int(np.count_nonzero(st & (16 << 8))), "step", np.unique(sol["step_info"][:, 0]).tolist()[:6], "max|cmd|", float(np.abs(cmd).max()))
    for b in bad[:3]:
        print("   robot", b, "step_info", sol["step_info"][b].tolist(), "t0", prob["t0"][b], "n_nodes", sol["n_nodes"][b], "wbc it?", "tau max", float(np.abs(cmd[b, 36:]).max()))
