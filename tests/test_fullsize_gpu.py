"""BASELINE.json's full sizes on the GPU, checked through size-independent properties (the oracle cannot run whole batches in seconds):
  * batch-position invariance: a robot's result does not depend on where it sits in the batch or on its neighbours (bit-identical),
  * a seeded sub-sample against the CPU oracle at the 1e-5 tolerance,
  * solver-independent identities: x[0] = x0, equality rows / defects of the accepted step, floating-base equation of motion, limits,
  * >= 20 consecutive warm-started ticks (config 4) tracked against the oracle tick by tick."""
import numpy as np
import pytest

from _parity import MPC_TOL, TICK_TOL, WBC_TOL, assert_cmd, assert_traj

pytestmark = pytest.mark.gpu


def test_config2_mpc_b1024_n100(oracle):
    """configs[1]: batched MPC only, state 30 / input 30, horizon 100, batch 1024."""
    import qm_control_b200 as q
    from qm_control_b200 import synthetic
    B = 1024; solver = q.Solver(batch=B, dt=0.01); oracle.mpc_set(dt=0.01, horizon=1.0)
    ids = np.arange(B); prob, _ = synthetic.make_batch(ids, config=2, horizon=1.0)
    out = solver.mpc_solve(prob)
    assert np.all((out["status"] & ~16) == 0), np.unique(out["status"])
    assert np.all(out["n_nodes"] >= 101)
    np.testing.assert_array_equal(out["x"][:, 0], prob["x0"])                         # the initial state is a hard constraint
    acc = out["step_info"][:, 0] > 0; assert acc.mean() > 0.99
    # multiple shooting: after an accepted full step from a cold start (state held, weight-compensating input) the dynamics defect drops
    dx, du, robot = solver.debug_get_step(); full = out["step_info"][:, 0] == 1.0
    assert full.sum() > 0 and np.mean(out["step_info"][full, 2] < robot[full, 2]) > 0.95
    # sub-sample vs the oracle
    sel = np.array([0, 7, 100, 511, 512, 777, 1000, 1023]); sub = {k: v[sel] for k, v in prob.items()}
    ref = oracle.mpc_solve_batch(sub, solver.nmax, nthreads=8)
    for i, b in enumerate(sel):
        assert out["step_info"][b, 0] == ref["dbg"][i, 0]; assert_traj(out, ref, MPC_TOL, tag="config2 b1024 robot %d" % b, b_out=b, b_ref=i)
    # batch-position invariance: reverse the batch
    solver2 = q.Solver(batch=B, dt=0.01); rev = ids[::-1].copy(); prob_r, _ = synthetic.make_batch(rev, config=2, horizon=1.0)
    out_r = solver2.mpc_solve(prob_r)
    for k in ("n_nodes", "t", "event", "x", "u", "status", "step_info"):
        assert np.array_equal(out[k], out_r[k][::-1]), k


def test_config3_wbc_b4096(oracle):
    """configs[2]: batched WBC HoQp, batch 4096."""
    import qm_control_b200 as q
    from qm_control_b200 import synthetic
    B = 4096; solver = q.Solver(batch=B); ids = np.arange(B)
    def run(idv, s):
        prob, wbc = synthetic.make_batch(idv, config=3)
        x_des, u_des, mode = synthetic.nominal_wbc_inputs(prob, s.robot_mass)
        u_des = u_des + synthetic.uniform(77, idv, 1, 30, -1.0, 1.0) * np.r_[np.full(12, 5.0), np.full(18, 0.2)]
        il = synthetic.uniform(78, idv, 2, 30, -0.1, 0.1); s.wbc_set_input_last(il)
        cmd, status = s.wbc_update(x_des, u_des, wbc["rbd"], mode, wbc["period"], np.full(len(idv), 12.0))
        return cmd, status, (x_des, u_des, mode, wbc, il)
    cmd, status, (x_des, u_des, mode, wbc, il) = run(ids, solver)
    assert np.all(status == 0), np.unique(status)
    sel = np.arange(0, B, 128)
    ref, _ = oracle.wbc_update_batch(x_des[sel], u_des[sel], wbc["rbd"][sel], mode[sel], wbc["period"][sel], np.full(len(sel), 12.0), il[sel], variant=0, nthreads=8)
    assert_cmd(cmd[sel], ref, WBC_TOL, tag="config3 wbc b4096")
    eff = oracle.model_info()["effort"]; lim = np.r_[np.tile(eff[:3], 4), eff[12:]]
    assert np.all(np.abs(cmd[:, 36:]) <= lim + 1e-6)                                  # torque limits, every robot
    F = cmd[:, 24:36].reshape(B, 4, 3); assert np.all(F[:, :, 2] >= -1e-7) and np.all(np.abs(F[:, :, :2]) <= 0.3 * F[:, :, 2:3] + 1e-6)   # friction pyramid
    for b in sel[::4]:                                                                # floating-base equation of motion
        rbd = wbc["rbd"][b]; qv = np.r_[rbd[3:6], rbd[0:3], rbd[6:24]]; z, y = qv[3], qv[4]
        T = np.array([[0, -np.sin(z), np.cos(y) * np.cos(z)], [0, np.cos(z), np.cos(y) * np.sin(z)], [1, 0, -np.sin(y)]])
        v = np.r_[rbd[27:30], np.linalg.solve(T, rbd[24:27]), rbd[30:48]]; r = oracle.rbd(qv, v)
        res = r["M"] @ cmd[b, :24] + r["nle"] - r["Jfoot"].T @ cmd[b, 24:36] - np.r_[np.zeros(6), cmd[b, 36:]]; assert np.max(np.abs(res)) < 1e-6
    cmd_r, status_r, _ = run(ids[::-1].copy(), q.Solver(batch=B))
    assert np.array_equal(cmd, cmd_r[::-1]) and np.array_equal(status, status_r[::-1])


def test_config4_full_tick_b8192_rows_do_not_depend_on_the_batch(oracle):
    """configs[3] shape on one GPU (the bench workload): 8192 robots, trot, N = 100.  Every row of the big batch equals the same robot
    solved in a small batch (bit-identical), a sub-sample matches the oracle, and the only flagged robot is the known indefinite one."""
    import qm_control_b200 as q
    from qm_control_b200 import synthetic
    B = 8192; solver = q.Solver(batch=B, dt=0.01); oracle.mpc_set(dt=0.01, horizon=1.0)
    prob, wbc = synthetic.make_batch(np.arange(B), config=4, horizon=1.0); t_eval = prob["t0"] + 0.002
    cmd, status = solver.tick(prob, t_eval, wbc["rbd"], wbc["period"])
    bad = np.nonzero(status & ~(16 << 8))[0]; assert set(bad) <= {1758}, bad             # tests/test_mpc_gpu.py::test_not_positive_definite_...
    sel = np.array([0, 1, 999, 4095, 4096, 8000, 8191]); small = q.Solver(batch=len(sel), dt=0.01)
    ps, ws = synthetic.make_batch(sel, config=4, horizon=1.0)
    cmd_s, status_s = small.tick(ps, ps["t0"] + 0.002, ws["rbd"], ws["period"])
    assert np.array_equal(cmd[sel], cmd_s) and np.array_equal(status[sel], status_s)
    ref = oracle.tick_batch(ps, small.nmax, ps["t0"] + 0.002, ws["rbd"], ws["period"], np.zeros((len(sel), 30)), nthreads=8)
    assert_cmd(cmd_s, ref["cmd"], TICK_TOL, tag="config4 b8192 tick cmd")
    assert_traj(small.mpc_get_solution(), ref, MPC_TOL, tag="config4 b8192 tick traj")


def test_config4_twenty_warm_started_ticks(oracle):
    """configs[3]: >= 20 consecutive ticks, MPC every 10 ms with warm start, WBC period 2 ms; both sides advance along the oracle's policy."""
    import qm_control_b200 as q
    from qm_control_b200 import synthetic
    B = 6; solver = q.Solver(batch=B, dt=0.015); oracle.mpc_set(dt=0.015, horizon=1.0)
    prob, wbc = synthetic.make_batch(np.arange(B), config=4); prev = None; il = np.zeros((B, 30)); worst = 0.0
    for tick in range(20):
        if tick > 0:
            prob = dict(prob); prob["t0"] = prob["t0"] + 0.01; x0 = np.zeros((B, 30))
            for b in range(B):
                n = prev["n_nodes"][b]; ne = prob["n_events"][b]
                x0[b], _, _ = oracle.evaluate_policy(prev["t"][b, :n], prev["event"][b, :n], prev["x"][b, :n], prev["u"][b, :n], prob["event_times"][b, :ne], prob["modes"][b, :ne + 1], prob["t0"][b])
            prob["x0"] = x0; solver.mpc_set_solution(prev)
        solver.wbc_set_input_last(il)
        t_eval = prob["t0"] + 0.002
        cmd, status = solver.tick(prob, t_eval, wbc["rbd"], wbc["period"]); assert np.all((status & ~(16 << 8)) == 0), (tick, np.unique(status))
        ref = oracle.tick_batch(prob, solver.nmax, t_eval, wbc["rbd"], wbc["period"], il, prev=prev, nthreads=6)
        lv = assert_traj(solver.mpc_get_solution(), ref, MPC_TOL, tag="config4 20 ticks: tick %d traj" % tick); worst = max(worst, max(lv.values()))
        assert_cmd(cmd, ref["cmd"], TICK_TOL, tag="config4 20 ticks: tick %d cmd" % tick)
        np.testing.assert_allclose(solver.wbc_get_input_last(), ref["input_last"], rtol=0, atol=1e-6)   # = the policy input of each side's own solution
        prev = {k: ref[k] for k in ("n_nodes", "t", "event", "x", "u")}; il = ref["input_last"]
    assert worst < MPC_TOL, worst


def test_config5_mixed_gaits_per_gpu_share(oracle):
    """configs[4]: mixed stance / trot / flying-trot batch, 16384 robots over 8 GPUs = 2048 per GPU (the shard one rank owns, global ids of rank 3),
    contact switches inside the horizon.  Rows equal the same robots solved alone (bit-identical), a sub-sample matches the oracle."""
    import qm_control_b200 as q
    from qm_control_b200 import synthetic
    from qm_control_b200.parallel import shard_range
    lo, hi = shard_range(16384, 3, 8); ids = np.arange(lo, hi); B = len(ids); assert B == 2048
    solver = q.Solver(batch=B, dt=0.01); oracle.mpc_set(dt=0.01, horizon=1.0)
    prob, wbc = synthetic.make_batch(ids, config=5, horizon=1.0); t_eval = prob["t0"] + 0.002
    cmd, status = solver.tick(prob, t_eval, wbc["rbd"], wbc["period"])
    flagged = np.nonzero(status & ~(16 << 8))[0]; assert len(flagged) <= 2, (ids[flagged], status[flagged])   # an indefinite projected Hessian is reported, never silent
    sol = solver.mpc_get_solution()
    assert {0, 15} <= set(np.unique(prob["modes"])) and sol["event"].max() == 2
    sel = np.array([0, 1, 2, 700, 701, 702, 2045, 2046, 2047]); sel = sel[~np.isin(sel, flagged)]
    small = q.Solver(batch=len(sel), dt=0.01); ps, ws = synthetic.make_batch(ids[sel], config=5, horizon=1.0)
    cmd_s, status_s = small.tick(ps, ps["t0"] + 0.002, ws["rbd"], ws["period"])
    assert np.array_equal(cmd[sel], cmd_s) and np.array_equal(status[sel], status_s)
    ref = oracle.tick_batch(ps, small.nmax, ps["t0"] + 0.002, ws["rbd"], ws["period"], np.zeros((len(sel), 30)), nthreads=9)
    assert_cmd(cmd_s, ref["cmd"], TICK_TOL, tag="config5 b2048 tick cmd")
    assert_traj(small.mpc_get_solution(), ref, MPC_TOL, tag="config5 b2048 tick traj")
