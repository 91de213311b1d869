"""The oracle's DDP variant (ddp{} block, discrete-time form): properties that do not depend on how the backward pass is computed."""
import numpy as np

from qm_control_b200 import synthetic

NMAX = 88


def test_ddp_step_is_a_rollout_and_decreases_the_merit(oracle):
    oracle.mpc_set(dt=0.015, horizon=1.0)
    try:
        oracle.mpc_set_solver(solver=2, iterations=1, ddp_penalty=20.0, ddp_min_step=1e-2, ddp_max_step=1.0)
        prob, _ = synthetic.make_batch(np.arange(6), config=5); out = oracle.mpc_solve_batch(prob, NMAX, nthreads=6)
        dbg = out["dbg"]; assert np.all(dbg[:, 0] > 0)                                               # a step length was accepted for every robot
        merit0 = dbg[:, 1] + 20.0 * np.sqrt(dbg[:, 3]); merit = dbg[:, 4] + 20.0 * np.sqrt(dbg[:, 6]); assert np.all(merit < merit0)
        assert np.all(dbg[:, 2] < 1e-20) and np.all(dbg[:, 5] < 1e-20)                               # single shooting: no dynamics defect before or after the step
        np.testing.assert_array_equal(out["x"][:, 0], prob["x0"])
        # the SQP step from the same start is a different point (multiple shooting moves the states freely)
        oracle.mpc_set_solver(solver=0); sq = oracle.mpc_solve_batch(prob, NMAX, nthreads=6); assert np.max(np.abs(sq["x"] - out["x"])) > 1e-6
    finally:
        oracle.mpc_set_solver(solver=0, iterations=1)
