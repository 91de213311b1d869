import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
import qm_control_b200 as q
from qm_control_b200 import synthetic
from _oracle import Oracle
np.set_printoptions(linewidth=200, precision=5, suppress=True)
config = int(sys.argv[1]) if len(sys.argv) > 1 else 2; B = int(sys.argv[2]) if len(sys.argv) > 2 else 3; dt = 0.015
o = Oracle(); o.mpc_set(dt=dt, horizon=1.0)
solver = q.Solver(batch=B, dt=dt)
prob, wbc = synthetic.make_batch(np.arange(B), config=config)
out = solver.mpc_solve(prob); dx, du, robot = solver.debug_get_step()
ref = o.mpc_solve_batch(prob, solver.nmax, nthreads=8)
print("status", out["status"], "n_nodes", out["n_nodes"], ref["n_nodes"])
print("gpu step_info [alpha cost dyn eq]\n", out["step_info"]); print("gpu robot [armijo basecost dyn eq |dx| |du|]\n", robot[:, :6])
print("oracle dbg [alpha basecost basedyn baseeq stepcost stepdyn stepeq armijo trials]\n", ref["dbg"])
for b in range(B):
    sub = {k: v[b:b+1] for k, v in prob.items()}
    d = o.mpc_debug(sub, solver.nmax, max_k=solver.nmax)
    n = d["n_nodes"]
    edx = np.abs(dx[b, :n] - d["dx"][:n]).max(axis=1); edu = np.abs(du[b, :n-1] - d["du"][:n-1]).max(axis=1)
    print("robot", b, "max|ddx|", edx.max(), "at node", edx.argmax(), "max|ddu|", edu.max(), "at node", edu.argmax(), " |dx|max", np.abs(d["dx"][:n]).max(), "|du|max", np.abs(d["du"][:n-1]).max())
    k = int(edu.argmax()); print("  du gpu   ", du[b, k]); print("  du oracle", d["du"][k]); print("  event flags", ref["event"][b, max(0,k-2):k+3])
    n = ref["n_nodes"][b]
    ex = np.abs(out["x"][b, :n] - ref["x"][b, :n]).max(); eu = np.abs(out["u"][b, :n-1] - ref["u"][b, :n-1]).max(); print("  final traj err x", ex, "u", eu)
