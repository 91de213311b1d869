"""CPU tests that PIN the oracle (parity is unpinned by the reference: it ships no tests or golden vectors).
Tier 1: analytic known-answer tests derivable from the reference inputs.  Tier 2: self-consistency (finite differences,
energy identities) against an independent numpy twin that parses the URDF itself."""
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from _oracle import URDF


# ---------------------------------------------------------------- independent numpy twin (own URDF parse, own FK)
def _rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr], [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr], [-sp, cp * sr, cp * cr]])


class Twin:
    """Un-lumped link tree straight from the URDF; q = [p, zyx, joints in oracle order]."""

    def __init__(self, joint_names):
        root = ET.parse(URDF).getroot(); f = lambda s: np.array([float(v) for v in s.split()])
        self.links = {}
        for l in root.findall("link"):
            i = l.find("inertial")
            if i is not None:
                o = i.find("origin"); xyz = f(o.get("xyz", "0 0 0")) if o is not None else np.zeros(3); rpy = f(o.get("rpy", "0 0 0")) if o is not None else np.zeros(3)
                a = i.find("inertia").attrib; I = np.array([[a["ixx"], a["ixy"], a["ixz"]], [a["ixy"], a["iyy"], a["iyz"]], [a["ixz"], a["iyz"], a["izz"]]], dtype=float)
                R = _rpy(*rpy); self.links[l.get("name")] = (float(i.find("mass").get("value")), xyz, R @ I @ R.T)
            else:
                self.links[l.get("name")] = (0.0, np.zeros(3), np.zeros((3, 3)))
        self.joints = []
        for j in root.findall("joint"):
            if j.get("type") is None:
                continue
            o = j.find("origin"); xyz = f(o.get("xyz", "0 0 0")) if o is not None else np.zeros(3); rpy = f(o.get("rpy", "0 0 0")) if o is not None else np.zeros(3)
            ax = f(j.find("axis").get("xyz")) if j.find("axis") is not None else np.array([1.0, 0, 0])
            self.joints.append(dict(name=j.get("name"), type=j.get("type"), parent=j.find("parent").get("link"), child=j.find("child").get("link"), xyz=xyz, R=_rpy(*rpy), axis=ax))
        self.qidx = {n: 6 + k for k, n in enumerate(joint_names)}
        children = {j["child"] for j in self.joints}; self.root = [n for n in self.links if n not in children][0]

    def fk(self, q):
        """world pose of every link frame."""
        cz, sz, cy, sy, cx, sx = np.cos(q[3]), np.sin(q[3]), np.cos(q[4]), np.sin(q[4]), np.cos(q[5]), np.sin(q[5])
        Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]); Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]); Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        pose = {self.root: (Rz @ Ry @ Rx, q[0:3].copy())}; todo = [self.root]
        while todo:
            p = todo.pop()
            for j in self.joints:
                if j["parent"] != p:
                    continue
                Rp, pp = pose[p]; R = Rp @ j["R"]; pos = pp + Rp @ j["xyz"]
                if j["type"] != "fixed":
                    a = j["axis"]; th = q[self.qidx[j["name"]]]; K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
                    R = R @ (np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K)
                pose[j["child"]] = (R, pos); todo.append(j["child"])
        return pose

    def com_and_bodies(self, q):
        pose = self.fk(q); out = []
        for n, (m, c, I) in self.links.items():
            R, p = pose[n]; out.append((m, p + R @ c, R @ I @ R.T, R))
        return out

    def potential(self, q):
        return sum(m * 9.81 * c[2] for m, c, _, _ in self.com_and_bodies(q))

    def kinetic(self, q, v, eps=1e-6):
        a = self.com_and_bodies(q + eps * v); b = self.com_and_bodies(q - eps * v); T = 0.0
        for (m, ca, Ia, Ra), (_, cb, _, Rb) in zip(a, b):
            cd = (ca - cb) / (2 * eps); W = (Ra - Rb) / (2 * eps) @ (0.5 * (Ra + Rb)).T; w = np.array([W[2, 1], W[0, 2], W[1, 0]])
            T += 0.5 * m * cd @ cd + 0.5 * w @ (Ia @ w)
        return T


@pytest.fixture(scope="module")
def twin(oracle):
    return Twin(oracle.model_info()["joint_names"])


def _rand_q(oracle, seed):
    rng = np.random.default_rng(seed); q = oracle.model_info()["q_nominal"].copy()
    q[0:3] = rng.uniform(-0.2, 0.2, 3) + [0, 0, 0.4]; q[3:6] = rng.uniform(-0.4, 0.4, 3); q[6:] += rng.uniform(-0.3, 0.3, 18)
    return q, rng.uniform(-0.5, 0.5, 24)


# ---------------------------------------------------------------- tier 1
def test_total_mass_and_joint_order(oracle):
    mi = oracle.model_info()
    assert abs(mi["mass"] - 27.371574) < 1e-9   # sum of <mass> in robot.urdf
    assert mi["joint_names"] == ["LF_HAA", "LF_HFE", "LF_KFE", "LH_HAA", "LH_HFE", "LH_KFE", "RF_HAA", "RF_HFE", "RF_KFE", "RH_HAA", "RH_HFE", "RH_KFE"] + ["j2n6s300_joint_%d" % i for i in range(1, 7)]   # task.info:168-188
    np.testing.assert_allclose(mi["effort"][:3], [35.278, 35.278, 44.4]); np.testing.assert_allclose(mi["effort"][12:], [40, 80, 40, 20, 20, 20])   # WbcBase.cpp:567-572


def test_nominal_end_effector_matches_controller_target(oracle):
    """QMController::starting commands the EE to (0.52, 0.09, 0.38 + z_base) (QMController.cpp:106) — the pose of the default joint state."""
    q = oracle.model_info()["q_nominal"].copy(); q[2] = 0.4
    ee = oracle.rbd(q, np.zeros(24))["ee_pos"]
    assert np.linalg.norm(ee - [0.52, 0.09, 0.38 + 0.4]) < 0.03


def test_gravity_and_mass_matrix_basics(oracle):
    mi = oracle.model_info(); q = mi["q_nominal"].copy(); q[2] = 0.4; r = oracle.rbd(q, np.zeros(24))
    assert abs(r["nle"][2] - mi["mass"] * 9.81) < 1e-9 and np.allclose(r["nle"][:2], 0)
    np.testing.assert_allclose(r["M"][:3, :3], mi["mass"] * np.eye(3), atol=1e-12)
    assert np.linalg.eigvalsh(r["M"]).min() > 0 and np.abs(r["M"] - r["M"].T).max() == 0


def test_weight_compensation_static_stance(oracle):
    """Static stance with weight-compensating forces: friction rows hold, torques within URDF limits, EoM residual zero."""
    mi = oracle.model_info(); q = mi["q_nominal"].copy(); q[2] = 0.4
    x = np.zeros(30); x[6:] = q; u = np.zeros(30); u[[2, 5, 8, 11]] = mi["mass"] * 9.81 / 4
    assert abs(u[2] - 67.1288) < 1e-3   # m g / 4 (QMInitializer.cpp:35-36)
    rbd = np.zeros(55); rbd[3:6] = q[:3]; rbd[0:3] = q[3:6]; rbd[6:24] = q[6:]
    cmd, _, _ = oracle.wbc_update(x, u, rbd, 15, 0.002, 20.0, input_last=u)
    r = oracle.rbd(q, np.zeros(24)); xx = cmd[:36]; tau = cmd[36:]
    res = r["M"] @ xx[:24] + r["nle"] - r["Jfoot"].T @ xx[24:] - np.r_[np.zeros(6), tau]
    assert np.abs(res).max() < 1e-8
    F = xx[24:].reshape(4, 3); assert np.all(F[:, 2] > 0) and np.all(np.abs(F[:, :2]) <= 0.3 * F[:, 2:3] + 1e-9)
    lim = np.r_[np.tile(mi["effort"][:3], 4), mi["effort"][12:]]; assert np.all(np.abs(tau) <= lim + 1e-9)


# ---------------------------------------------------------------- tier 2: independent twin
def test_forward_kinematics_matches_independent_twin(oracle, twin):
    for seed in range(3):
        q, v = _rand_q(oracle, seed); r = oracle.rbd(q, v); pose = twin.fk(q)
        for i, n in enumerate(["LF_FOOT", "RF_FOOT", "LH_FOOT", "RH_FOOT"]):
            np.testing.assert_allclose(r["foot_pos"][i], pose[n][1], atol=1e-12)
        np.testing.assert_allclose(r["ee_pos"], pose["j2n6s300_end_effector"][1], atol=1e-12); np.testing.assert_allclose(r["ee_rot"], pose["j2n6s300_end_effector"][0], atol=1e-12)
        bodies = twin.com_and_bodies(q); com = sum(m * c for m, c, _, _ in bodies) / sum(m for m, _, _, _ in bodies)
        np.testing.assert_allclose(r["com"], com, atol=1e-12)   # fixed-joint lumping preserves the mass distribution


def test_mass_matrix_is_the_kinetic_energy_form(oracle, twin):
    q, _ = _rand_q(oracle, 7); M = oracle.rbd(q, np.zeros(24))["M"]; rng = np.random.default_rng(1)
    for _ in range(4):
        v = rng.uniform(-1, 1, 24); assert abs(0.5 * v @ M @ v - twin.kinetic(q, v)) < 1e-6 * max(1.0, 0.5 * v @ M @ v)


def test_gravity_vector_is_potential_gradient(oracle, twin):
    q, _ = _rand_q(oracle, 11); g = oracle.rbd(q, np.zeros(24))["nle"]; eps = 1e-6
    fd = np.array([(twin.potential(q + eps * e) - twin.potential(q - eps * e)) / (2 * eps) for e in np.eye(24)])
    np.testing.assert_allclose(g, fd, atol=1e-6)


def test_coriolis_power_identity(oracle, twin):
    """v'(C v) = 1/2 v' Mdot v  (skew-symmetry of Mdot - 2C), with Mdot by finite differences of the oracle's M."""
    q, v = _rand_q(oracle, 13); eps = 1e-6
    r = oracle.rbd(q, v); g = oracle.rbd(q, np.zeros(24))["nle"]; cv = r["nle"] - g
    Md = (oracle.rbd(q + eps * v, v)["M"] - oracle.rbd(q - eps * v, v)["M"]) / (2 * eps)
    assert abs(v @ cv - 0.5 * v @ Md @ v) < 1e-6


def test_jacobians_and_time_derivatives_by_finite_differences(oracle):
    q, v = _rand_q(oracle, 17); eps = 1e-6; r = oracle.rbd(q, v)
    Jfd = np.stack([(oracle.rbd(q + eps * e, v)["foot_pos"].ravel() - oracle.rbd(q - eps * e, v)["foot_pos"].ravel()) / (2 * eps) for e in np.eye(24)], axis=1)
    np.testing.assert_allclose(r["Jfoot"], Jfd, atol=1e-8)
    np.testing.assert_allclose(r["foot_vel"].ravel(), r["Jfoot"] @ v, atol=1e-12)
    dJ = (oracle.rbd(q + eps * v, v)["Jfoot"] - oracle.rbd(q - eps * v, v)["Jfoot"]) / (2 * eps)
    np.testing.assert_allclose(r["dJfoot"], dJ, atol=1e-7)
    dJe = (oracle.rbd(q + eps * v, v)["Jee"] - oracle.rbd(q - eps * v, v)["Jee"]) / (2 * eps)
    np.testing.assert_allclose(r["dJee"], dJe, atol=1e-7)
    # centroidal momentum matrix: Ag v = [m com_dot; angular momentum about the COM], and its rate
    comd = (oracle.rbd(q + eps * v, v)["com"] - oracle.rbd(q - eps * v, v)["com"]) / (2 * eps)
    np.testing.assert_allclose((r["Ag"] @ v)[:3], oracle.model_info()["mass"] * comd, atol=1e-6)
    hd = (oracle.rbd(q + eps * v, v)["Ag"] @ v - oracle.rbd(q - eps * v, v)["Ag"] @ v) / (2 * eps)
    np.testing.assert_allclose(r["dAg_v"], hd, atol=1e-6)


def test_flow_map_jacobians_by_finite_differences(oracle):
    rng = np.random.default_rng(3); mi = oracle.model_info()
    x = np.r_[rng.uniform(-0.2, 0.2, 6), 0.05, -0.03, 0.41, rng.uniform(-0.3, 0.3, 3), mi["q_nominal"][6:] + rng.uniform(-0.2, 0.2, 18)]
    u = np.r_[rng.uniform(-20, 20, 12) + np.tile([0, 0, 67.0], 4), rng.uniform(-0.5, 0.5, 18)]
    f, A, B = oracle.flow_map(x, u); eps = 1e-6
    Afd = np.stack([(oracle.flow_map(x + eps * e, u)[0] - oracle.flow_map(x - eps * e, u)[0]) / (2 * eps) for e in np.eye(30)], axis=1)
    Bfd = np.stack([(oracle.flow_map(x, u + eps * e)[0] - oracle.flow_map(x, u - eps * e)[0]) / (2 * eps) for e in np.eye(30)], axis=1)
    np.testing.assert_allclose(A, Afd, atol=1e-6); np.testing.assert_allclose(B, Bfd, atol=1e-7)
    assert abs(f[2] - (u[2] + u[5] + u[8] + u[11]) / mi["mass"] + 9.81) < 1e-12   # vcom_z dot = sum F_z / m - g
    np.testing.assert_allclose(f[12:], u[12:])                                     # joint positions integrate the joint-velocity inputs


def test_centroidal_state_round_trip(oracle):
    """computeCentroidalStateFromRbdModel followed by the SRBD velocity mapping returns the measured base velocity."""
    q, v = _rand_q(oracle, 23); z, y = q[3], q[4]
    T = np.array([[0, -np.sin(z), np.cos(y) * np.cos(z)], [0, np.cos(z), np.cos(y) * np.sin(z)], [1, 0, -np.sin(y)]])
    rbd = np.r_[q[3:6], q[0:3], q[6:], T @ v[3:6], v[0:3], v[6:]]
    x = oracle.centroidal_state_from_rbd(rbd); np.testing.assert_allclose(x[6:], q)
    u = np.r_[np.zeros(12), v[6:]]; f, _, _ = oracle.flow_map(x, u)
    np.testing.assert_allclose(f[6:12], v[:6], atol=1e-12)


def test_swing_reference_spline(oracle):
    """SwingTrajectoryPlanner knots: lift-off 0.05 m/s, apex 0.15 m, touch-down -0.1 m/s (task.info:23-30), swing 0.35 s ≥ swingTimeScale."""
    ev = [0.0, 0.35, 0.7, 1.05]; modes = [15, 6, 9, 6, 15]   # LF (foot 0) swings during RF_LH = mode 6
    zp, zv = oracle.swing_reference(ev, modes, 0, 1e-9); assert abs(zp) < 1e-9 and abs(zv - 0.05) < 1e-6
    zp, zv = oracle.swing_reference(ev, modes, 0, 0.175); assert abs(zp - 0.15) < 1e-12 and abs(zv) < 1e-12
    zp, zv = oracle.swing_reference(ev, modes, 0, 0.35 - 1e-9); assert abs(zp) < 1e-8 and abs(zv + 0.1) < 1e-6
    assert oracle.swing_reference(ev, modes, 1, 0.2) == (0.0, 0.0)   # RF is in stance during mode 6


def test_cold_start_mpc_grid_and_feasibility(oracle):
    """One SQP iteration from the QMInitializer guess: grid = 67 intervals of 0.015 s (+ stance-template event nodes),
    stance legs keep zero joint velocity (zero-velocity rows), the accepted step improves cost or violation."""
    from qm_control_b200 import synthetic
    oracle.mpc_set(dt=0.015, horizon=1.0)
    prob, _ = synthetic.make_batch(np.arange(2), config=2)
    out = oracle.mpc_solve_batch(prob, 140, nthreads=2)
    for b in range(2):
        n = out["n_nodes"][b]; t = out["t"][b, :n]; ev = out["event"][b, :n]
        assert abs(t[0] - 12.0) < 1e-12 and abs(t[-1] - 13.0) < 1e-12 and 68 + int((ev == 1).sum()) <= n <= 68 + 2 * int((ev == 1).sum())
        assert np.all(np.diff(t) <= 0.015 + 1e-12)
        alpha, base_cost, base_dyn, base_eq, step_cost, step_dyn, step_eq = out["dbg"][b, :7]
        assert alpha > 0 and (step_cost < base_cost or step_dyn + step_eq < base_dyn + base_eq)   # filter line search: cost or violation improves
        assert step_eq < 0.1 * base_eq and step_dyn < 0.1 * base_dyn   # the projected QP step satisfies the linearised constraints: residuals drop to second order


def test_srbd_flow_map_matches_independent_twin(oracle, twin):
    """The continuous dynamics of the OCP (QMDynamicsAD.cpp:22-33 → PinocchioCentroidalDynamicsAD, SRBD) restated with the twin's own FK and its own
    nominal constants: hdot/m = [g + sum F/m ; sum (p_foot - r_com) x F / m], qdot_base = A_b^{-1} m h, A_b = [[m I, m S(R c) T], [0, R I_nom R' T]], r_com = p - R c."""
    info = oracle.model_info(); qn = info["q_nominal"]; bodies = twin.com_and_bodies(qn); mass = sum(b[0] for b in bodies)
    com = sum(b[0] * b[1] for b in bodies) / mass; c_nom = qn[0:3] - com                                     # comToBasePositionNominal (base orientation is zero at the nominal pose)
    I_nom = sum(b[2] + b[0] * ((b[1] - com) @ (b[1] - com) * np.eye(3) - np.outer(b[1] - com, b[1] - com)) for b in bodies)
    np.testing.assert_allclose(mass, info["mass"], rtol=1e-12); np.testing.assert_allclose(c_nom, info["com_to_base"], atol=1e-12); np.testing.assert_allclose(I_nom, info["inertia_nominal"], atol=1e-10)
    feet = ("LF_FOOT", "RF_FOOT", "LH_FOOT", "RH_FOOT")                                                       # contact order (ModelSettings.h:38)
    for seed in range(5):
        q, _ = _rand_q(oracle, 40 + seed); rng = np.random.default_rng(seed); h = rng.uniform(-0.3, 0.3, 6); u = np.r_[rng.uniform(-30, 60, 12), rng.uniform(-1, 1, 18)]
        x = np.r_[h, q]; f, _, _ = oracle.flow_map(x, u)
        pose = twin.fk(q); R = pose[twin.root][0]; z, y = q[3], q[4]
        T = np.array([[0, -np.sin(z), np.cos(y) * np.cos(z)], [0, np.cos(z), np.cos(y) * np.sin(z)], [1, 0, -np.sin(y)]])
        c = R @ c_nom; r_com = q[0:3] - c; S = np.array([[0, -c[2], c[1]], [c[2], 0, -c[0]], [-c[1], c[0], 0]])
        Ab = np.block([[mass * np.eye(3), mass * S @ T], [np.zeros((3, 3)), R @ I_nom @ R.T @ T]])
        F = u[:12].reshape(4, 3); lin = np.array([0, 0, -9.81]) + F.sum(0) / mass; ang = sum(np.cross(pose[n][1] - r_com, F[i]) for i, n in enumerate(feet)) / mass
        ref = np.r_[lin, ang, np.linalg.solve(Ab, mass * h), u[12:]]
        np.testing.assert_allclose(f, ref, rtol=0, atol=1e-10)


def test_equality_constraint_values_match_independent_twin(oracle, twin):
    """ZeroVelocityConstraintCppAd (stance: foot velocity = 0), ZeroForceConstraint (swing: F = 0) and NormalVelocityConstraintCppAd (swing: v_z - zdot_ref = 0,
    NormalVelocityConstraintCppAd.cpp:37-66, positionErrorGain 0): the residuals the oracle stacks per node against foot velocities obtained by differentiating the
    twin's own FK along qdot = [A_b^{-1} m h ; joint velocities]."""
    from qm_control_b200 import synthetic
    info = oracle.model_info(); mass = info["mass"]; c_nom = info["com_to_base"]; I_nom = info["inertia_nominal"]; feet = ("LF_FOOT", "RF_FOOT", "LH_FOOT", "RH_FOOT")
    prob, _ = synthetic.make_batch(np.array([2]), config=5); ne = int(prob["n_events"][0]); et = prob["event_times"][0, :ne]; md = prob["modes"][0, :ne + 1]
    tt = prob["target_times"][0, :2]; ts = prob["target_states"][0, :2]; checked = set()
    for j, t in enumerate(np.linspace(12.01, 12.9, 12)):
        q, _ = _rand_q(oracle, 60 + j); rng = np.random.default_rng(j); h = rng.uniform(-0.3, 0.3, 6); u = np.r_[rng.uniform(-30, 60, 12), rng.uniform(-1, 1, 18)]; x = np.r_[h, q]
        mode = md[int(np.searchsorted(et, t, side="left"))]; flags = [(mode >> (3 - f)) & 1 for f in range(4)]
        _, _, _, g = oracle.stage_probe(et, md, tt, ts, t, x, u, want_grad=False)
        R = twin.fk(q)[twin.root][0]; z, y = q[3], q[4]; T = np.array([[0, -np.sin(z), np.cos(y) * np.cos(z)], [0, np.cos(z), np.cos(y) * np.sin(z)], [1, 0, -np.sin(y)]])
        c = R @ c_nom; S = np.array([[0, -c[2], c[1]], [c[2], 0, -c[0]], [-c[1], c[0], 0]]); Ab = np.block([[mass * np.eye(3), mass * S @ T], [np.zeros((3, 3)), R @ I_nom @ R.T @ T]])
        v = np.r_[np.linalg.solve(Ab, mass * h), u[12:]]; eps = 1e-6; pa = twin.fk(q + eps * v); pb = twin.fk(q - eps * v)
        ref = []
        for i, n in enumerate(feet):                                        # constraint order: per foot in contact order, stance → 3 velocity rows; swing → 3 force rows + 1 normal-velocity row
            vf = (pa[n][1] - pb[n][1]) / (2 * eps)
            if flags[i]:
                ref += list(vf)
            else:
                zp, zv = oracle.swing_reference(et, md, i, t); ref += list(u[3 * i:3 * i + 3]) + [vf[2] - zv]
        ref = np.array(ref); assert len(g) == len(ref) == 3 * sum(flags) + 4 * (4 - sum(flags))
        # the oracle stacks the rows by constraint type, the twin by foot: compare the residuals as multisets
        np.testing.assert_allclose(np.sort(g), np.sort(ref), rtol=0, atol=2e-8); checked.add(sum(flags))
    assert checked >= {0, 2}


def _barrier(mu, delta, h):   # ocs2 RelaxedBarrierPenalty [upstream]
    return -mu * np.log(h) if h > delta else mu * (-np.log(delta) + 0.5 * ((h - 2.0 * delta) / delta) ** 2 - 0.5)


def test_input_weight_and_stage_cost_value_match_independent_twin(oracle, twin):
    """(a) initializeInputCostWeight (QMInterface.cpp:274-299): R[12:24,12:24] = J' R_task J with J = d(foot positions)/d(leg joints) at initialState, J from central
    differences of the twin's FK.  (b) the intermediate cost VALUE at random points from its published pieces: quadratic tracking (LeggedRobotQuadraticTrackingCost.h:34-40),
    end-effector penalty 1/2 mu |e|^2 (task.info:118-122), relaxed barriers on the arm joint box (task.info:165-200, URDF limits) and on the friction cone
    (task.info:159-164, regularisation 25 [upstream default]); EE pose from the twin's FK, quaternions from scipy."""
    from scipy.spatial.transform import Rotation
    from qm_control_b200 import synthetic, _lib
    info = oracle.model_info(); Q, R = oracle.mpc_weights(); qn = info["q_nominal"]; feet = ("LF_FOOT", "RF_FOOT", "LH_FOOT", "RH_FOOT")
    txt = open(_lib.asset("qm_task.info")).read(); import re as _re
    blk = _re.search(r"(?m)^R\s*\n?\{(.*?)^\}", txt, _re.S).group(1); scaling = float(_re.search(r"scaling\s+([-+.\deE]+)", blk).group(1))
    Rtask = np.zeros(30)
    for i, j, v in _re.findall(r"\((\d+),(\d+)\)\s+([-+.\deE]+)", blk):
        assert i == j; Rtask[int(i)] = scaling * float(v)
    eps = 1e-6; J = np.zeros((12, 12))
    for k in range(12):
        d = np.zeros(24); d[6 + k] = eps; pa = twin.fk(qn + d); pb = twin.fk(qn - d)
        J[:, k] = np.concatenate([(pa[n][1] - pb[n][1]) / (2 * eps) for n in feet])
    Rref = np.diag(Rtask); Rref[12:24, 12:24] = J.T @ np.diag(Rtask[12:24]) @ J
    np.testing.assert_allclose(R, Rref, rtol=0, atol=1e-9)
    Qblk = _re.search(r"(?m)^Q\s*\n?\{(.*?)^\}", txt, _re.S).group(1); Qs = float(_re.search(r"scaling\s+([-+.\deE]+)", Qblk).group(1)); Qref = np.zeros((30, 30))
    for i, j, v in _re.findall(r"\((\d+),(\d+)\)\s+([-+.\deE]+)", Qblk):
        Qref[int(i), int(j)] = Qs * float(v)
    np.testing.assert_allclose(Q, Qref, rtol=0, atol=0)
    # (b) cost value
    prob, _ = synthetic.make_batch(np.array([4]), config=4); ne = int(prob["n_events"][0]); et = prob["event_times"][0, :ne]; md = prob["modes"][0, :ne + 1]
    tt = prob["target_times"][0, :2]; ts = prob["target_states"][0, :2]; lo = info["lower"][12:]; hi = info["upper"][12:]; vlim = np.array([0.628, 0.628, 0.628, 0.837, 0.837, 0.837])
    for j, t in enumerate(np.linspace(12.05, 12.95, 6)):
        q, _ = _rand_q(oracle, 80 + j); rng = np.random.default_rng(100 + j); q[18:24] = np.clip(q[18:24], lo + 0.05, hi - 0.05)
        x = np.r_[rng.uniform(-0.2, 0.2, 6), q]; u = np.r_[np.tile([3.0, -4.0, 60.0], 4) + rng.uniform(-2, 2, 12), rng.uniform(-0.5, 0.5, 18)]
        mode = md[int(np.searchsorted(et, t, side="left"))]; flags = [(mode >> (3 - f)) & 1 for f in range(4)]
        f_or, _, _, _ = oracle.stage_probe(et, md, tt, ts, t, x, u, want_grad=False)
        a = (tt[1] - t) / (tt[1] - tt[0]); xref = a * ts[0, :30] + (1 - a) * ts[1, :30]; un = np.zeros(30)
        for f in range(4):
            if flags[f]:
                un[3 * f + 2] = info["mass"] * 9.81 / sum(flags)
        val = 0.5 * (x - xref) @ Q @ (x - xref) + 0.5 * (u - un) @ R @ (u - un)
        Ree, pee = twin.fk(q)["j2n6s300_end_effector"]; pref = ts[0, 30:33]; qref = ts[0, 33:37]                     # constant EE target over the horizon
        qe = Rotation.from_matrix(Ree).as_quat(); dist = qe[3] * qref[:3] - qref[3] * qe[:3] + np.cross(qe[:3], qref[:3])    # quaternionDistance(q_ee, q_ref)
        val += 0.5 * 2000.0 * np.sum((pee - pref) ** 2) + 0.5 * 1000.0 * np.sum(dist ** 2)
        for i in range(6):
            val += _barrier(0.1, 1e-3, x[24 + i] - lo[i]) + _barrier(0.1, 1e-3, hi[i] - x[24 + i]) + _barrier(0.1, 1e-3, u[24 + i] + vlim[i]) + _barrier(0.1, 1e-3, vlim[i] - u[24 + i])
        for f in range(4):
            if flags[f]:
                F = u[3 * f:3 * f + 3]; val += _barrier(0.1, 5.0, 0.3 * F[2] - np.sqrt(F[0] ** 2 + F[1] ** 2 + 25.0))
        np.testing.assert_allclose(f_or, val, rtol=1e-10, atol=1e-8)
