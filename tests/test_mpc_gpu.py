"""Parity of the CUDA MPC tick (through the C-ABI) against the CPU oracle: one multiple-shooting SQP iteration
(QMController.cpp:287-288 → SqpSolver::runImpl) on the OCP of qm_interface.  Contract: 1e-5 relative on the optimal
state / input trajectories (BASELINE.json north_star); ASSERTED: 1e-8 per block of like quantities (tests/_parity.py)."""
import numpy as np
import pytest

from _parity import MPC_TOL, TICK_TOL, assert_cmd, assert_traj

pytestmark = pytest.mark.gpu


def _solve_both(oracle, config, B, dt, ticks=2, gait=None, horizon=1.0):
    import qm_control_b200 as q
    from qm_control_b200 import synthetic
    solver = q.Solver(batch=B, dt=dt, time_horizon=horizon)
    oracle.mpc_set(dt=dt, horizon=horizon)
    prob, wbc = synthetic.make_batch(np.arange(B), config=config, gait=gait, horizon=horizon)
    prev = None; results = []
    for tick in range(ticks):
        if tick > 0:   # advance 10 ms along the oracle's own policy (same x0 for both)
            prob["t0"] = prob["t0"] + 0.01
            x0 = np.zeros((B, 30))
            for b in range(B):
                n = prev["n_nodes"][b]; ne = prob["n_events"][b]
                x0[b], _, _ = oracle.evaluate_policy(prev["t"][b, :n], prev["event"][b, :n], prev["x"][b, :n], prev["u"][b, :n], prob["event_times"][b, :ne], prob["modes"][b, :ne + 1], prob["t0"][b])
            prob["x0"] = x0
        out = solver.mpc_solve(prob)
        ref = oracle.mpc_solve_batch(prob, solver.nmax, prev=prev, nthreads=8)
        results.append((out, ref))
        prev = ref
        # hand the oracle's solution to the CUDA path as warm start so both ticks start from identical data
        ref_fixed = dict(ref); solver.mpc_set_solution(ref_fixed)
    return solver, prob, results


def _check(results, tag="mpc"):
    for tick, (out, ref) in enumerate(results):
        assert np.all((out["status"] & ~16) == 0), "tick %d status %s" % (tick, np.unique(out["status"]))
        np.testing.assert_allclose(out["step_info"][:, 0], ref["dbg"][:, 0], rtol=0, atol=0, err_msg="line-search step size differs (tick %d)" % tick)
        assert_traj(out, ref, MPC_TOL, tag="%s tick %d" % (tag, tick))
        acc = ref["dbg"][:, 0] > 0
        np.testing.assert_allclose(out["step_info"][acc, 1], ref["dbg"][acc, 4], rtol=1e-6, atol=1e-8)   # cost after the step


def test_mpc_stance_reference_grid(oracle):
    """config 1/2: stance, reference default grid (dt = 0.015 → 67 intervals + stance-template event nodes)."""
    _, _, results = _solve_both(oracle, config=2, B=8, dt=0.015)
    _check(results, "mpc_stance_reference_grid")


def test_mpc_stance_n100(oracle):
    """config 2: horizon 100 nodes (dt = 0.01)."""
    _, _, results = _solve_both(oracle, config=2, B=8, dt=0.01, ticks=2)
    _check(results, "mpc_stance_n100")


def test_mpc_trot_contact_switches(oracle):
    """config 4: trot — swing legs (zero force + normal velocity rows), event nodes inside the horizon."""
    _, prob, results = _solve_both(oracle, config=4, B=8, dt=0.015)
    assert results[0][1]["event"].max() == 2
    _check(results, "mpc_trot_contact_switches")


def test_mpc_mixed_gaits(oracle):
    """config 5: stance / trot / flying trot (n_c in {4,2,0})."""
    _, _, results = _solve_both(oracle, config=5, B=12, dt=0.015)
    _check(results, "mpc_mixed_gaits")


def test_policy_eval_matches_oracle(oracle):
    solver, prob, results = _solve_both(oracle, config=4, B=6, dt=0.015, ticks=1)
    out, ref = results[0]
    solver.mpc_set_solution(ref)
    tq = prob["t0"] + 0.0123
    xd, ud, mode = solver.policy_eval(tq)
    for b in range(6):
        n = ref["n_nodes"][b]; ne = prob["n_events"][b]
        x, u, m = oracle.evaluate_policy(ref["t"][b, :n], ref["event"][b, :n], ref["x"][b, :n], ref["u"][b, :n], prob["event_times"][b, :ne], prob["modes"][b, :ne + 1], tq[b])
        np.testing.assert_allclose(xd[b], x, rtol=0, atol=1e-12); np.testing.assert_allclose(ud[b], u, rtol=0, atol=1e-10); assert mode[b] == m


def test_full_tick_matches_oracle(oracle):
    """config 4 loop: mpc_solve → evaluatePolicy → wbc update through qmb200_tick vs the same chain on the oracle."""
    import qm_control_b200 as q
    from qm_control_b200 import synthetic
    B = 8; solver = q.Solver(batch=B, dt=0.015); oracle.mpc_set(dt=0.015, horizon=1.0)
    prob, wbc = synthetic.make_batch(np.arange(B), config=4)
    t_eval = prob["t0"] + 0.002
    cmd, status = solver.tick(prob, t_eval, wbc["rbd"], wbc["period"])
    assert np.all((status & ~(16 << 8)) == 0), np.unique(status)
    ref = oracle.mpc_solve_batch(prob, solver.nmax, nthreads=8)
    for b in range(B):
        n = ref["n_nodes"][b]; ne = prob["n_events"][b]
        x, u, m = oracle.evaluate_policy(ref["t"][b, :n], ref["event"][b, :n], ref["x"][b, :n], ref["u"][b, :n], prob["event_times"][b, :ne], prob["modes"][b, :ne + 1], t_eval[b])
        c, _, _ = oracle.wbc_update(x, u, wbc["rbd"][b], m, wbc["period"][b], t_eval[b], input_last=np.zeros(30))
        assert_cmd(cmd[b], c, TICK_TOL, tag="full_tick robot %d" % b)   # chain MPC -> policy -> WBC


def test_pipelined_tick_is_bit_identical():
    """qmb200_set_pipeline only changes how robot ranges are scheduled (streams): commands, status and the stored
    solution must be bit-identical to the single-chain tick, over two consecutive ticks (warm start crosses the flip)."""
    import qm_control_b200 as q
    from qm_control_b200 import synthetic
    B = 37; prob, wbc = synthetic.make_batch(np.arange(B), config=4)
    outs = []
    for chunks in (1, 3, 8):
        solver = q.Solver(batch=B, dt=0.015); solver.set_pipeline(chunks)
        res = []
        for tick in range(2):
            p = dict(prob); p["t0"] = prob["t0"] + 0.01 * tick
            cmd, status = solver.tick(p, p["t0"] + 0.002, wbc["rbd"], wbc["period"])
            sol = solver.mpc_get_solution()
            res.append((cmd.copy(), status.copy(), sol["x"].copy(), sol["u"].copy()))
        outs.append(res)
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            for xa, xb in zip(a, b):
                assert np.array_equal(xa, xb)


def test_not_positive_definite_is_reported_like_the_oracle(oracle):
    """Synthetic robot 1758 of the bench workload (dt 0.01, trot) has an indefinite projected input Hessian: the oracle raises
    (HPIPM would fail in the reference), the CUDA path must flag MST_NOT_PD | MST_NO_STEP for that robot only and keep its warm start."""
    import qm_control_b200 as q
    from qm_control_b200 import synthetic
    ids = np.array([1758, 5]); solver = q.Solver(batch=2, dt=0.01); oracle.mpc_set(dt=0.01, horizon=1.0)
    prob, _ = synthetic.make_batch(ids, config=4, horizon=1.0)
    out = solver.mpc_solve(prob)
    assert out["status"][0] == (8 | 16 | 64), hex(int(out["status"][0]))    # NOT_PD | NO_STEP | NEG_DT: the root cause is named (tests/test_neg_interval_cpu.py)
    assert out["status"][1] & ~16 == 0
    cmd, st = solver.tick(prob, prob["t0"] + 0.002, *[synthetic.make_batch(ids, config=4, horizon=1.0)[1][k] for k in ("rbd", "period")])
    assert st[0] == (8 | 16 | 64) << 8 and (st[1] & ~(16 << 8)) == 0, [hex(int(v)) for v in st]   # merged word: MPC flags in bits 8..15, WBC byte clean
    with pytest.raises(RuntimeError, match="not positive definite"):
        oracle.mpc_solve_batch({k: v[:1] for k, v in prob.items()}, solver.nmax, nthreads=1)
    oracle.mpc_solve_batch({k: v[1:] for k, v in prob.items()}, solver.nmax, nthreads=1)


def test_dense_state_weight_falls_back_to_the_general_path(tmp_path):
    """task.info's Q is diagonal and the kernels use that (DevModel::q_is_diag); a Q with off-diagonal entries is still a legal input of
    QMInterface (loadEigenMatrix reads a dense 30x30): the general path must give the oracle's answer for it."""
    import qm_control_b200 as q
    from qm_control_b200 import synthetic, _lib
    from _oracle import Oracle, URDF, REFERENCE, GAINS
    import re
    txt = open(_lib.asset("qm_task.info")).read()
    m = re.search(r"(?m)^Q\s*\n?\{", txt); assert m
    txt = txt[:m.end()] + "\n  (0,1) 3.0\n  (1,0) 3.0\n  (6,9) 20.0\n  (9,6) 20.0\n" + txt[m.end():]
    task = tmp_path / "task_dense_q.info"; task.write_text(txt)
    B = 4; solver = q.Solver(q.QMInterface(taskFile=str(task)), batch=B, dt=0.015); orc = Oracle(URDF, str(task), REFERENCE, GAINS); orc.mpc_set(dt=0.015, horizon=1.0)
    Q, _ = orc.mpc_weights(); assert Q[0, 1] != 0 and Q[6, 9] != 0
    prob, _ = synthetic.make_batch(np.arange(B), config=4)
    out = solver.mpc_solve(prob); ref = orc.mpc_solve_batch(prob, solver.nmax, nthreads=4)
    assert np.all((out["status"] & ~16) == 0); np.testing.assert_array_equal(out["step_info"][:, 0], ref["dbg"][:, 0])
    assert_traj(out, ref, MPC_TOL, tag="mpc_dense_Q")


def test_multiple_sqp_iterations_and_convergence_exit(oracle):
    """sqp.sqpIteration > 1 (SqpSolver::runImpl loop): three iterations from a cold start track the oracle's loop; with loose tolerances the
    convergence test (checkConvergence: step size / metrics / primal step) ends the loop early for the same robots on both sides."""
    import qm_control_b200 as q
    from qm_control_b200 import synthetic
    B = 9; solver = q.Solver(batch=B, dt=0.015); oracle.mpc_set(dt=0.015, horizon=1.0)
    prob, _ = synthetic.make_batch(np.arange(B), config=5)
    try:
        solver.mpc_set_iterations(3); oracle.mpc_set_sqp(3)
        out = solver.mpc_solve(prob); ref = oracle.mpc_solve_batch(prob, solver.nmax, nthreads=8)
        assert np.all(ref["dbg"][:, 9] >= 2)                              # the cold start is far from converged: more than one iteration everywhere
        assert np.all((out["status"] & ~(16 | 32)) == 0), np.unique(out["status"])
        np.testing.assert_array_equal(((out["status"] & 32) != 0), ref["dbg"][:, 9] < 3)
        np.testing.assert_array_equal(out["step_info"][:, 0], ref["dbg"][:, 0])
        assert_traj(out, ref, 10 * MPC_TOL, tag="mpc_3_sqp_iterations")   # three chained iterations
        one = q.Solver(batch=B, dt=0.015).mpc_solve(prob)                  # and the extra iterations did move the solution
        assert max(np.max(np.abs(one["x"] - out["x"])), np.max(np.abs(one["u"] - out["u"]))) > 1e-6
        # early exit: robot 1758 of the bench workload has an indefinite projected Hessian -> no step -> checkConvergence(STEPSIZE) ends its loop after the
        # first iteration (status NOT_PD | NO_STEP | CONVERGED, warm start kept) while its neighbour runs all three iterations and matches the oracle
        ids = np.array([1758, 5]); s2 = q.Solver(batch=2, dt=0.01); s2.mpc_set_iterations(3); oracle.mpc_set(dt=0.01, horizon=1.0)
        p2, _ = synthetic.make_batch(ids, config=4, horizon=1.0); o2 = s2.mpc_solve(p2)
        assert o2["status"][0] == (8 | 16 | 32 | 64) and (o2["status"][1] & ~16) == 0, o2["status"]
        r2 = oracle.mpc_solve_batch({k: v[1:] for k, v in p2.items()}, s2.nmax, nthreads=1); assert r2["dbg"][0, 9] == 3
        assert_traj(o2, r2, 10 * MPC_TOL, tag="mpc_3_sqp_iterations neighbour", b_out=1, b_ref=0)
    finally:
        oracle.mpc_set_sqp(1)
