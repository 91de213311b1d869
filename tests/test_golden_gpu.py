"""The CUDA path (through the C-ABI) against the committed golden fixtures (tests/golden/*.npz) — no oracle build involved.
Contract 1e-5 relative on trajectories and the WBC vector (BASELINE.json north_star); asserted 1e-8 per block (tests/_parity.py), 1e-12 on the target front-end."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import make_golden as mg  # noqa: E402
from _parity import CMD_BLOCKS, MPC_TOL, MPCWBC_TOL, WBC_TOL, assert_cmd, assert_traj  # noqa: E402


def test_wbc_against_golden():
    import qm_control_b200 as q
    g = np.load(os.path.join(GOLD, "wbc_config5.npz")); ids = g["ids"]; B = len(ids); solver = q.Solver(batch=B)
    x_des, u_des, mode, wbc, il = mg.wbc_inputs(ids, solver.robot_mass)
    solver.wbc_set_input_last(il)
    cmd, status = solver.wbc_update(x_des, u_des, wbc["rbd"], mode, wbc["period"], np.full(B, 12.0)); assert np.all(status == 0)
    assert_cmd(cmd, g["cmd"], WBC_TOL, tag="golden wbc_config5")
    np.testing.assert_array_equal(solver.wbc_get_input_last(), g["input_last"])


def test_wbc_mpc_variant_against_golden():
    """HierarchicalMpcWbc: the 12 leg torques QMMpcController consumes (QMController.cpp:427-431), and every other block, against the fixture."""
    import qm_control_b200 as q
    g = np.load(os.path.join(GOLD, "wbc_mpc_variant_config3.npz")); ids = g["ids"]; B = len(ids); solver = q.Solver(batch=B, wbc_variant=1)
    x_des, u_des, mode, wbc, il = mg.wbc_mpc_inputs(ids, solver.robot_mass)
    solver.wbc_set_input_last(il)
    cmd, status = solver.wbc_update(x_des, u_des, wbc["rbd"], mode, wbc["period"], np.full(B, 12.0)); assert np.all(status == 0)
    blocks = dict(CMD_BLOCKS); blocks["arm_acc"] = (18, 24, 1e3)
    assert_cmd(cmd, g["cmd"], MPCWBC_TOL, tag="golden wbc_mpc_variant", blocks=blocks)


def test_mpc_against_golden():
    import qm_control_b200 as q
    from qm_control_b200 import synthetic
    g = np.load(os.path.join(GOLD, "mpc_config5_dt015.npz")); ids = g["ids"]; B = len(ids); solver = q.Solver(batch=B, dt=0.015); assert solver.nmax == int(g["nmax"])
    prob, _ = synthetic.make_batch(ids, config=5)
    for tick, pre in ((1, "t1_"), (2, "t2_")):
        if tick == 2:
            prob = dict(prob); prob["t0"] = prob["t0"] + 0.01; prob["x0"] = g["x0_tick2"]
            solver.mpc_set_solution({k: g["t1_" + k] for k in ("n_nodes", "t", "event", "x", "u")})
        out = solver.mpc_solve(prob)
        np.testing.assert_array_equal(out["n_nodes"], g[pre + "n_nodes"]); np.testing.assert_array_equal(out["step_info"][:, 0], g[pre + "dbg"][:, 0])
        assert_traj(out, {k: g[pre + k] for k in ("n_nodes", "t", "event", "x", "u")}, MPC_TOL, tag="golden mpc_config5 tick %d" % tick)


def test_target_against_golden():
    import qm_control_b200 as q
    from qm_control_b200 import synthetic
    g = np.load(os.path.join(GOLD, "target_config4.npz")); solver = q.Solver(batch=4); prob, _ = synthetic.make_batch(np.arange(4), config=4)
    for kind in range(3):
        nt, tt, ts, le = solver.target_trajectories(kind, np.tile(g["cmds"][kind], (4, 1)), np.full(4, 12.0), prob["x0"], g["ee"], g["last"])
        np.testing.assert_allclose(tt[:, :2], g["times_%d" % kind], rtol=0, atol=1e-12); np.testing.assert_allclose(ts[:, :2], g["states_%d" % kind], rtol=0, atol=1e-12)
        np.testing.assert_array_equal(le, g["last_%d" % kind])
