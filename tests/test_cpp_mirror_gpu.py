"""The C++ mirrors (include/qmb200.hpp) give the same numbers as the Python mirrors of the same reference classes: HierarchicalWbc::update and the
QMController starting → advanceMpc → update sequence, through examples/plugin_demo (bit-identical: both are thin layers over the same C-ABI calls)."""
import os
import subprocess

import numpy as np
import pytest

from test_cpp_mirror_cpu import ROOT, build_demo

pytestmark = pytest.mark.gpu


def test_cpp_plugin_demo_matches_python_mirror(tmp_path):
    import qm_control_b200 as q
    from qm_control_b200 import synthetic
    prob, wbc = synthetic.make_batch(np.arange(1), config=3); solver = q.Solver(batch=1, dt=0.015, time_horizon=1.0)
    x_des, u_des, mode = synthetic.nominal_wbc_inputs(prob, solver.robot_mass); rbd = wbc["rbd"]; period, time, t_start = 0.002, 12.0, 12.0
    inp = tmp_path / "in.txt"; inp.write_text(" ".join("%.17g" % v for v in np.r_[x_des[0], u_des[0], rbd[0], float(mode[0]), period, time, t_start]) + "\n")
    out = subprocess.run([build_demo(), os.path.join(ROOT, "assets"), str(inp)], capture_output=True, text=True); assert out.returncode == 0, out.stdout + out.stderr
    rows = {ln.split()[0]: ln.split()[1:] for ln in out.stdout.strip().splitlines() if " " in ln}
    cmd, status = solver.wbc_update(x_des, u_des, rbd, mode, np.array([period]), np.array([time]))
    np.testing.assert_array_equal(np.array(rows["wbc"], dtype=float), cmd[0]); assert int(rows["wbc_status"][0]) == int(status[0])
    # controller sequence with the Python mirror of the same class
    ctrl = q.QMController(batch=1, dt=0.015, time_horizon=1.0); ctrl.starting(rbd, time=t_start)
    p = dict(prob); p["n_events"] = np.array([2], dtype=np.int32); p["event_times"] = np.zeros((1, 32)); p["event_times"][0, :2] = [t_start - 1.0, t_start + 5.0]; p["modes"] = np.full((1, 33), 15, dtype=np.int32)
    tgt = np.zeros(37); tgt[6:30] = ctrl.x_obs[0, 6:30]; tgt[30:37] = rbd[0, 48:55]
    p["n_target"] = np.array([2], dtype=np.int32); p["target_times"] = np.zeros((1, 4)); p["target_times"][0, :2] = [t_start, t_start + 1.0]; p["target_states"] = np.zeros((1, 4, 37)); p["target_states"][0, :2] = tgt
    sol = ctrl.advanceMpc(p)
    assert int(rows["mpc_nodes"][0]) == int(sol["n_nodes"][0]) and int(rows["mpc_nodes"][2]) == int(sol["status"][0]) and float(rows["mpc_nodes"][4]) == sol["step_info"][0, 0]
    np.testing.assert_array_equal(np.array(rows["mpc_x1"], dtype=float), sol["x"][0, 1])
    cmd2, st2 = ctrl.update(rbd, period)
    np.testing.assert_array_equal(np.array(rows["update"], dtype=float), cmd2[0]); assert int(rows["update_status"][0]) == int(st2[0]) and float(rows["update_status"][4]) == ctrl.t_obs[0]
    np.testing.assert_array_equal(np.array(rows["joint_cmd"], dtype=float).reshape(18, 5), ctrl.joint_cmd[0])
