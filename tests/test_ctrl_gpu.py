"""Parity of the controller-side CUDA kernels (observation update, target front-end, control law, plant law, fused
QMController::update) against the CPU oracle (oracle/src/ctrl.cpp), through the C-ABI.  These are pure maps:
the bar is 1e-12 absolute (libm vs CUDA sincos/fmod round-off), bit-exact where no transcendental is involved."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from _parity import CMD_BLOCKS, MPCWBC_TOL, WBC_TOL, assert_cmd

pytestmark = pytest.mark.gpu


def _batch(B, config=4):
    from qm_control_b200 import synthetic
    return synthetic.make_batch(np.arange(B), config=config)


def _ee_states(B, seed):
    rng = np.random.default_rng(seed)
    return np.c_[rng.uniform(-1, 1, (B, 3)), Rotation.random(B, random_state=seed).as_quat()]


def test_observation_update_matches_oracle(oracle):
    import qm_control_b200 as q
    B = 333; solver = q.Solver(batch=B); prob, wbc = _batch(B)
    rng = np.random.default_rng(1); rbd = wbc["rbd"].copy(); rbd[:, 0] = rng.uniform(-np.pi, np.pi, B)             # yaw anywhere on the circle
    x_prev = prob["x0"].copy(); x_prev[:, 9] = rbd[:, 0] + rng.choice([-4, -2, 0, 2, 4], B) * np.pi + rng.uniform(-3.0, 3.0, B)   # previous unwrapped yaw, several turns away
    t_prev = rng.uniform(0, 20, B); period = rng.uniform(0.001, 0.003, B)
    t, x = solver.observation_update(rbd, period, t_prev, x_prev)
    for b in range(B):
        tr, xr = oracle.observation_update(rbd[b], period[b], t_prev[b], x_prev[b])
        assert t[b] == tr; np.testing.assert_allclose(x[b], xr, rtol=0, atol=1e-12)
        assert abs(x[b, 9] - x_prev[b, 9]) <= np.pi + 1e-9


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_target_trajectories_match_oracle(kind):
    import qm_control_b200 as q
    from _oracle import TargetOracle
    B = 257; solver = q.Solver(batch=B); prob, wbc = _batch(B); to = TargetOracle()
    rng = np.random.default_rng(10 + kind); ee = _ee_states(B, 20 + kind); last = np.c_[ee[:, :3] + rng.uniform(-0.12, 0.12, (B, 3)), Rotation.random(B, random_state=30 + kind).as_quat()]
    cmd = rng.uniform(-0.5, 0.5, (B, 7))
    if kind == 2:
        cmd[:, 3:7] = Rotation.random(B, random_state=40).as_quat()
    t_obs = rng.uniform(0, 30, B)
    nt, tt, ts, le = solver.target_trajectories(kind, cmd, t_obs, prob["x0"], ee, last)
    assert np.all(nt == 2) and np.all(tt[:, 2:] == 0) and np.all(ts[:, 2:] == 0)
    for b in range(B):
        times, states, lr = to.target(kind, cmd[b], t_obs[b], prob["x0"][b], ee[b], last[b])
        np.testing.assert_allclose(tt[b, :2], times, rtol=0, atol=1e-12); np.testing.assert_allclose(ts[b, :2], states, rtol=0, atol=1e-12); np.testing.assert_array_equal(le[b], lr)
    # the produced targets are accepted by the solver as they are (same layout as qmb200_mpc_solve's target arguments)
    p = dict(prob); p["n_target"] = nt; p["target_times"] = tt + 0.0; p["target_times"][:, 0] = prob["t0"]; p["target_times"][:, 1] = np.maximum(tt[:, 1] - tt[:, 0], 0.2) + prob["t0"]; p["target_states"] = ts
    out = solver.mpc_solve(p); assert np.all((out["status"] & ~16) == 0)


@pytest.mark.parametrize("variant", [0, 1])
def test_control_law_matches_oracle(oracle, variant):
    import qm_control_b200 as q
    B = 200; solver = q.Solver(batch=B, wbc_variant=variant); solver.set_arm_gains(1.5, 0.75)
    rng = np.random.default_rng(variant); xd = rng.normal(size=(B, 30)); ud = rng.normal(size=(B, 30)); w = rng.normal(size=(B, 54)); xo = rng.normal(size=(B, 30)); xo[:, 11] = rng.uniform(-2.0, 2.0, B)
    t = rng.uniform(9.0, 11.0, B); jc0 = rng.normal(size=(B, 18, 5)); ap0 = rng.normal(size=(B, 6)); lt0 = t - rng.uniform(0.0, 0.02, B)
    jc, ap, lt, st = solver.control_law(xd, ud, w, t, xo, jc0, ap0, lt0)
    for b in range(B):
        jr, ar, lr, safe = oracle.control_law(variant, 1.5, 0.75, xd[b], ud[b], w[b], t[b], xo[b], jc0[b], ap0[b], lt0[b])
        np.testing.assert_array_equal(jc[b], jr); np.testing.assert_array_equal(ap[b], ar); assert lt[b] == lr and bool(st[b]) == (not safe)
    assert st.sum() > 0 and (st == 0).sum() > 0


def test_hw_write_delay_fifo_matches_oracle():
    import qm_control_b200 as q
    from _oracle import HwSimOracle
    B = 23; solver = q.Solver(batch=B); solver.hw_set_delay(0.009); sims = [HwSimOracle(0.009) for _ in range(B)]
    rng = np.random.default_rng(7); period = np.where(np.arange(B) % 2 == 0, 0.001, 0.002); start = rng.integers(1, 4, B)   # robots whose first stamp equals the period reset their FIFO
    for k in range(50):
        t = (start + k) * period; jc = rng.normal(size=(B, 18, 5)); pos = rng.normal(size=(B, 18)); vel = rng.normal(size=(B, 18))
        eff, st = solver.hw_write(t, period, jc, pos, vel); assert np.all(st == 0)
        for b in range(B):
            np.testing.assert_array_equal(eff[b], sims[b].write(t[b], period[b], jc[b], pos[b], vel[b]))
    solver.hw_set_delay(1.0)                                      # 1 s window at 1 kHz overflows the 32-entry ring: flagged, never silent
    flagged = 0
    for k in range(40):
        eff, st = solver.hw_write(np.full(B, 100.0 + 0.001 * k), np.full(B, 0.001), np.zeros((B, 18, 5)), np.zeros((B, 18)), np.zeros((B, 18))); flagged += int(np.any(st == 2))
    assert flagged > 0


@pytest.mark.parametrize("variant", [0, 1])
def test_controller_update_matches_oracle_chain(oracle, variant):
    """QMController::update = observation update → evaluatePolicy → WbcBase::update → safety + control law, three consecutive RT ticks on one policy."""
    import qm_control_b200 as q
    B = 12; ctrl = (q.QMMpcController if variant else q.QMController)(batch=B, dt=0.015); solver = ctrl.solver; oracle.mpc_set(dt=0.015, horizon=1.0)
    prob, wbc = _batch(B, config=3 if variant else 4); rbd = wbc["rbd"].copy()   # HierarchicalMpcWbc: stance + realistic joint accelerations (see tests/test_wbc_gpu.py)
    ctrl.starting(rbd, time=12.0)
    np.testing.assert_allclose(ctrl.x_obs, np.stack([oracle.centroidal_state_from_rbd(r) for r in rbd]), atol=1e-12)
    p = dict(prob); p["t0"] = ctrl.t_obs.copy(); p["x0"] = ctrl.x_obs.copy()
    ref = oracle.mpc_solve_batch(p, solver.nmax, nthreads=8); solver.mpc_solve(p); solver.mpc_set_solution(ref)      # both sides evaluate the same policy
    t_o = ctrl.t_obs.copy(); x_o = ctrl.x_obs.copy(); il = np.zeros((B, 30)); jc_o = np.zeros((B, 18, 5)); ap_o = np.zeros((B, 6)); lt_o = ctrl.last_time.copy()
    rng = np.random.default_rng(4)
    if variant:   # WbcBase::inputLast_ = the policy's input at the start time, so that (u - inputLast_)/period is a realistic joint acceleration
        for b in range(B):
            n = ref["n_nodes"][b]; ne = p["n_events"][b]
            _, il[b], _ = oracle.evaluate_policy(ref["t"][b, :n], ref["event"][b, :n], ref["x"][b, :n], ref["u"][b, :n], p["event_times"][b, :ne], p["modes"][b, :ne + 1], t_o[b])
        solver.wbc_set_input_last(il)
    for tick in range(3):
        rbd = rbd + rng.normal(size=rbd.shape) * 1e-3; period = 0.002
        cmd, status = ctrl.update(rbd, period); assert np.all(status == 0), np.unique(status)
        for b in range(B):
            t_o[b], x_o[b] = oracle.observation_update(rbd[b], period, t_o[b], x_o[b])
            n = ref["n_nodes"][b]; ne = p["n_events"][b]
            xd, ud, mode = oracle.evaluate_policy(ref["t"][b, :n], ref["event"][b, :n], ref["x"][b, :n], ref["u"][b, :n], p["event_times"][b, :ne], p["modes"][b, :ne + 1], t_o[b])
            cb, ilb = oracle.wbc_update_batch(xd[None], ud[None], rbd[b][None], [mode], [period], [t_o[b]], il[b][None], variant=variant); c = cb[0]; il[b] = ilb[0]   # (the batch entry tolerates the oracle's QP iteration cap in the free arm directions)
            jc_o[b], ap_o[b], lt_o[b], safe = oracle.control_law(variant, 0.0, 0.5, xd, ud, c, t_o[b], x_o[b], jc_o[b], ap_o[b], lt_o[b])
            blocks = dict(CMD_BLOCKS)
            if variant:
                blocks["arm_acc"] = (18, 24, 1e3)                       # HierarchicalMpcWbc: arm accelerations of O(1e4) behind a 3e3-conditioned block (tests/test_wbc_gpu.py)
            assert_cmd(cmd[b], c, MPCWBC_TOL if variant else WBC_TOL, tag="controller update variant %d tick %d" % (variant, tick), blocks=blocks)
            np.testing.assert_allclose(ctrl.x_obs[b], x_o[b], atol=1e-12); assert ctrl.t_obs[b] == t_o[b]
            legs = np.abs(ctrl.joint_cmd[b, :12] - jc_o[b, :12]); assert legs[:, :4].max() < 1e-9 and (variant or legs[:, 4].max() < 1e-3)
            if variant == 0:
                arm = np.abs(ctrl.joint_cmd[b, 12:] - jc_o[b, 12:]); assert arm[:, :4].max() < 1e-9 and arm[:, 4].max() < 1e-3
            else:
                np.testing.assert_allclose(ctrl.arm_pos_cmd[b], ap_o[b], atol=1e-9); assert ctrl.last_time[b] == lt_o[b]
