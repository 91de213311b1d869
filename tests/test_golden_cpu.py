"""The oracle reproduces the committed golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py): guards the checker
against silent drift between rounds / compilers.  Tolerance 1e-9 relative (FMA contraction may differ between hosts)."""
import os
import sys

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import make_golden as mg  # noqa: E402


def _close(a, b, tol=1e-9):
    np.testing.assert_allclose(a, b, rtol=tol, atol=tol * max(1.0, float(np.max(np.abs(b)))))


def test_wbc_golden(oracle):
    g = np.load(os.path.join(GOLD, "wbc_config5.npz")); x_des, u_des, mode, wbc, il = mg.wbc_inputs(g["ids"], oracle.model_info()["mass"])
    np.testing.assert_array_equal(mode, g["mode"]); assert set(mode) >= {0, 15}
    cmd, il_out = oracle.wbc_update_batch(x_des, u_des, wbc["rbd"], mode, wbc["period"], np.full(len(mode), 12.0), il, variant=0, nthreads=2)
    _close(cmd, g["cmd"], 1e-8); _close(il_out, g["input_last"])


def test_wbc_mpc_variant_golden(oracle):
    g = np.load(os.path.join(GOLD, "wbc_mpc_variant_config3.npz")); x_des, u_des, mode, wbc, il = mg.wbc_mpc_inputs(g["ids"], oracle.model_info()["mass"])
    cmd, il_out = oracle.wbc_update_batch(x_des, u_des, wbc["rbd"], mode, wbc["period"], np.full(len(mode), 12.0), il, variant=1, nthreads=2)
    _close(cmd[:, 36:48], g["cmd"][:, 36:48], 1e-8); _close(cmd[:, 24:36], g["cmd"][:, 24:36], 1e-8); _close(cmd[:, :18], g["cmd"][:, :18], 1e-7)   # leg torques, forces, base / leg accelerations
    _close(cmd[:, 18:24], g["cmd"][:, 18:24], 1e-6)                                                                                         # arm accelerations of O(1e4)


def test_mpc_golden(oracle):
    from qm_control_b200 import synthetic
    g = np.load(os.path.join(GOLD, "mpc_config5_dt015.npz")); oracle.mpc_set(dt=0.015, horizon=1.0)
    prob, _ = synthetic.make_batch(g["ids"], config=5); nmax = int(g["nmax"])
    t1 = oracle.mpc_solve_batch(prob, nmax, nthreads=3)
    np.testing.assert_array_equal(t1["n_nodes"], g["t1_n_nodes"]); np.testing.assert_array_equal(t1["event"], g["t1_event"]); _close(t1["t"], g["t1_t"], 1e-13)
    _close(t1["x"], g["t1_x"], 1e-8); _close(t1["u"], g["t1_u"], 1e-8); np.testing.assert_array_equal(t1["dbg"][:, 0], g["t1_dbg"][:, 0])
    prob2 = dict(prob); prob2["t0"] = prob["t0"] + 0.01; prob2["x0"] = g["x0_tick2"]
    prev = {k: g["t1_" + k] for k in ("n_nodes", "t", "event", "x", "u")}
    t2 = oracle.mpc_solve_batch(prob2, nmax, prev=prev, nthreads=3)
    np.testing.assert_array_equal(t2["n_nodes"], g["t2_n_nodes"]); _close(t2["x"], g["t2_x"], 1e-8); _close(t2["u"], g["t2_u"], 1e-8)


def test_target_golden():
    from _oracle import TargetOracle
    from qm_control_b200 import synthetic
    g = np.load(os.path.join(GOLD, "target_config4.npz")); to = TargetOracle(); prob, _ = synthetic.make_batch(np.arange(4), config=4)
    for kind in range(3):
        for b in range(4):
            times, states, last = to.target(kind, g["cmds"][kind], 12.0, prob["x0"][b], g["ee"][b], g["last"][b])
            _close(times, g["times_%d" % kind][b], 1e-13); _close(states, g["states_%d" % kind][b], 1e-13); _close(last, g["last_%d" % kind][b], 1e-13)
