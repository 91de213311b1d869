"""Pins for the oracle's restatement of the controller-side steps (oracle/src/ctrl.cpp) against an independent
numpy/scipy transliteration of the reference sources, plus known-answer cases.  The reference functions are reference-owned code
(qm_controllers/src/QmTargetTrajectoriesPublisher_node.cpp, QMController.cpp, qm_gazebo/src/QMHWSim.cpp), so these pins are exact."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from _oracle import HwSimOracle, TargetOracle
from qm_control_b200 import synthetic

COM_HEIGHT, DISP_VEL, ROT_VEL, T_TARGET = 0.4, 0.3, 0.1, 1.0   # reference.info:1-4, task.info mpc.timeHorizon


def _djs():
    from qm_control_b200 import _lib
    return synthetic._info_vector(_lib.asset("qm_reference.info"), "defaultJointState", 18)


def _twin(kind, cmd, t, x, ee, last):
    """Direct transliteration with scipy rotations (Eigen quaternion = scipy xyzw; zyx euler = intrinsic 'ZYX')."""
    last = last.copy(); base = x[6:12].copy(); vel = np.zeros(3)
    if kind == 0:
        vel = Rotation.from_euler("ZYX", base[3:6]).apply(cmd[:3])
        bt = np.array([base[0] + vel[0] * T_TARGET, base[1] + vel[1] * T_TARGET, COM_HEIGHT, base[3] + cmd[3] * T_TARGET, 0, 0])
        if np.linalg.norm(last[:3] - ee[:3]) > 0.1:
            last[:3] = ee[:3]
        et = last.copy(); ec = et.copy(); tr = t + T_TARGET
    elif kind == 1:
        v = (Rotation.from_quat(ee[3:7]) * Rotation.from_quat([0.5, -0.5, 0.5, -0.5]).inv()).apply(cmd[:3])
        et = np.r_[ee[0] + v[0] * T_TARGET, ee[1] + v[1] * T_TARGET, last[2:7]]; ec = ee.copy()
        bt = np.array([et[0] - 0.52, et[1] - 0.09, COM_HEIGHT, base[3], 0, 0]); tr = t + T_TARGET
    else:
        et = cmd[:7].copy(); ec = ee.copy(); bt = np.array([cmd[0] - 0.52, cmd[1] - 0.09, COM_HEIGHT, base[3], 0, 0])
        qc, qt = ee[3:7], cmd[3:7]
        dq = qc[3] * qt[:3] - qt[3] * qc[:3] + np.cross(qc[:3], qt[:3])
        tr = t + max(np.linalg.norm(dq) / ROT_VEL, np.linalg.norm(cmd[:3] - ee[:3]) / DISP_VEL); last = cmd[:7].copy()
    bc = base.copy(); bc[2] = COM_HEIGHT; bc[4] = bc[5] = 0
    s0 = np.r_[vel, np.zeros(3), bc, _djs(), ec]; s1 = np.r_[vel, np.zeros(3), bt, _djs(), et]
    return np.array([t, tr]), np.stack([s0, s1]), last


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_target_trajectories_match_scipy_twin(kind):
    to = TargetOracle(); rng = np.random.default_rng(5 + kind)
    for trial in range(20):
        x = rng.uniform(-0.5, 0.5, 30); ee = np.r_[rng.uniform(-1, 1, 3), Rotation.random(random_state=trial).as_quat()]
        last = np.r_[ee[:3] + rng.uniform(-0.12, 0.12, 3), Rotation.random(random_state=100 + trial).as_quat()]
        cmd = rng.uniform(-0.5, 0.5, 7)
        if kind == 2:
            cmd[3:7] = Rotation.random(random_state=200 + trial).as_quat()
        t = 3.0 + trial
        times, states, le = to.target(kind, cmd, t, x, ee, last)
        rt, rs, rl = _twin(kind, cmd, t, x, ee, last)
        np.testing.assert_allclose(times, rt, rtol=0, atol=1e-12); np.testing.assert_allclose(states, rs, rtol=0, atol=1e-12); np.testing.assert_allclose(le, rl, rtol=0, atol=0)


def test_initial_controller_target_is_reproduced():
    """Known answer: with the robot at its nominal pose and zero cmd_vel the first knot equals the second and the EE target is lastEeTarget_."""
    to = TargetOracle(); x = np.zeros(30); x[8] = 0.4; ee = np.array([0.52, 0.09, 0.44, 0.5, -0.5, 0.5, -0.5])
    times, states, le = to.target(0, np.zeros(4), 2.0, x, ee, ee)
    np.testing.assert_allclose(times, [2.0, 3.0]); np.testing.assert_allclose(states[0], states[1]); np.testing.assert_allclose(states[0, 30:], ee); np.testing.assert_allclose(states[0, 12:30], _djs())


def test_observation_yaw_unwrap_and_time(oracle):
    """Yaw keeps counting through +-pi (QMController.cpp:241); the other coordinates are the centroidal conversion."""
    rbd = np.zeros(55); rbd[3:6] = [0, 0, 0.4]; rbd[6:24] = oracle.model_info()["q_nominal"][6:]
    x = np.zeros(30); t = 1.0; yaws = []
    for k, yaw in enumerate(np.arange(0.0, 9.0, 0.7)):           # true yaw grows past 2 pi; the measured one wraps into (-pi, pi]
        rbd[0] = np.arctan2(np.sin(yaw), np.cos(yaw))
        t, x = oracle.observation_update(rbd, 0.002, t, x); yaws.append(x[9])
        np.testing.assert_allclose(x[10:12], 0.0); np.testing.assert_allclose(x[12:30], rbd[6:24])
    np.testing.assert_allclose(yaws, np.arange(0.0, 9.0, 0.7), atol=1e-12); assert abs(t - (1.0 + 0.002 * len(yaws))) < 1e-12
    ref = oracle.centroidal_state_from_rbd(rbd); ref[9] = x[9]
    np.testing.assert_allclose(x, ref, atol=1e-14)


def test_control_law_branches(oracle):
    rng = np.random.default_rng(3); xd = rng.normal(size=30); ud = rng.normal(size=30); w = rng.normal(size=54); xo = rng.normal(size=30) * 0.1
    jc0 = rng.normal(size=(18, 5))
    jc, ap, lt, safe = oracle.control_law(0, 0.0, 0.5, xd, ud, w, 5.0, xo, jc0, np.zeros(6), 0.0)          # t < 10: legs untouched
    np.testing.assert_array_equal(jc[:12], jc0[:12]); np.testing.assert_allclose(jc[12:], np.c_[xd[24:30], np.zeros(6), np.zeros(6), np.full(6, 0.5), w[48:54]]); assert safe
    jc, ap, lt, safe = oracle.control_law(0, 0.0, 0.5, xd, ud, w, 10.5, xo, jc0, np.zeros(6), 0.0)
    np.testing.assert_allclose(jc[:12], np.c_[xd[12:24], ud[12:24], np.zeros(12), np.full(12, 3.0), w[36:48]])
    jc, ap, lt, safe = oracle.control_law(1, 0.0, 0.5, xd, ud, w, 10.5, xo, jc0, np.zeros(6), 10.495)        # 100 Hz gate closed
    np.testing.assert_array_equal(jc[12:], jc0[12:]); np.testing.assert_array_equal(ap, 0.0); assert lt == 10.495
    jc, ap, lt, safe = oracle.control_law(1, 0.0, 0.5, xd, ud, w, 10.5, xo, jc0, np.zeros(6), 10.48)
    np.testing.assert_allclose(ap, xo[24:30] + ud[24:30] / 100.0); assert lt == 10.5
    xo[11] = 1.6
    assert not oracle.control_law(0, 0.0, 0.5, xd, ud, w, 10.5, xo, jc0, np.zeros(6), 0.0)[3]


def test_hw_sim_delay_known_answer():
    """With delay d the applied command is the oldest one stamped >= t - d (QMHWSim.cpp:105-111); time == period resets."""
    hw = HwSimOracle(0.009); pos = np.zeros(18); vel = np.zeros(18); period = 0.001
    stamps = []
    for k in range(1, 40):
        t = k * period; jc = np.zeros((18, 5)); jc[:, 4] = k; eff = hw.write(t, period, jc, pos, vel); stamps.append(t)
        expect = min(i for i in range(1, k + 1) if i * period + 0.009 >= t)
        assert eff[0] == expect, (k, eff[0], expect)
    jc = np.zeros((18, 5)); jc[:, 0] = 1.0; jc[:, 2] = 10.0; jc[:, 3] = 2.0; jc[:, 4] = 0.5     # PD part: kp (1 - 0.25) + kd (0 - 0.5) + ff
    hw2 = HwSimOracle(0.0); eff = hw2.write(0.5, 0.001, jc, np.full(18, 0.25), np.full(18, 0.5))
    np.testing.assert_allclose(eff, 10.0 * 0.75 - 2.0 * 0.5 + 0.5)
