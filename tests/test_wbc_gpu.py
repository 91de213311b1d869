"""Parity of the CUDA WBC path (through the C-ABI) against the CPU oracle — the reference's WbcBase::update →
HierarchicalWbc::update → HoQp → updateCmd chain (qm_wbc/src/*.cpp).  Contract: 1e-5 relative (BASELINE.json north_star);
ASSERTED: 1e-8 per robot and per block of like quantities (accelerations, forces, torques: tests/_parity.py)."""
import numpy as np
import pytest

from _parity import CMD_BLOCKS, MPCWBC_TOL, WBC_TOL, assert_cmd

pytestmark = pytest.mark.gpu


def _run(oracle, config, B, variant, time, gait=None, perturb_u=True):
    import qm_control_b200 as q
    from qm_control_b200 import synthetic
    solver = q.Solver(batch=B, wbc_variant=variant)
    prob, wbc = synthetic.make_batch(np.arange(B), config=config, gait=gait)
    x_des, u_des, mode = synthetic.nominal_wbc_inputs(prob, solver.robot_mass)
    if perturb_u:   # desired forces / joint velocities away from the nominal so every task row is exercised
        u_des = u_des + synthetic.uniform(77, np.arange(B), 1, 30, -1.0, 1.0) * np.r_[np.full(12, 5.0), np.full(18, 0.2)]
        for b in range(B):
            for f in range(4):
                if not (mode[b] >> (3 - f)) & 1:
                    u_des[b, 3 * f:3 * f + 3] = 0.0
    il = synthetic.uniform(78, np.arange(B), 2, 30, -0.1, 0.1)
    tarr = np.full(B, time)
    solver.wbc_set_input_last(il)
    cmd, status = solver.wbc_update(x_des, u_des, wbc["rbd"], mode, wbc["period"], tarr)
    ref, il_ref = oracle.wbc_update_batch(x_des, u_des, wbc["rbd"], mode, wbc["period"], tarr, il, variant=variant, nthreads=8)
    assert np.all(status == 0), "status flags: %s" % np.unique(status)
    np.testing.assert_array_equal(solver.wbc_get_input_last(), u_des)   # WbcBase::inputLast_ update (WbcBase.cpp:213)
    assert_cmd(cmd, ref, WBC_TOL, tag="wbc config %d variant %d t=%g B=%d" % (config, variant, time, B))
    return cmd, ref, mode


def test_wbc_stance_matches_oracle(oracle):
    _run(oracle, config=3, B=256, variant=0, time=12.0)


def test_wbc_init_branch_matches_oracle(oracle):
    """time < 10 → task0 → taskInit → task2 (HierarchicalWbc.cpp:32-37)."""
    _run(oracle, config=3, B=64, variant=0, time=3.0)


def test_wbc_trot_matches_oracle(oracle):
    cmd, ref, mode = _run(oracle, config=4, B=256, variant=0, time=12.0)
    assert set(np.unique(mode)) <= {6, 9, 15}


def test_wbc_mixed_gaits_match_oracle(oracle):
    """stance / trot / flying trot incl. zero-contact modes (n_c in {4,2,0})."""
    cmd, ref, mode = _run(oracle, config=5, B=384, variant=0, time=12.0)
    assert 0 in mode and 15 in mode


def _mpc_variant_case(B):
    from qm_control_b200 import synthetic
    ids = np.arange(B); prob, wbc = synthetic.make_batch(ids, config=3)
    return ids, prob, wbc


def test_wbc_mpc_variant_matches_oracle(oracle):
    """HierarchicalMpcWbc (HierarchicalMpcWbc.cpp:18-34): task1 = height + base angular + base linear + 100 swing, task2 = contact force.  The six arm
    accelerations carry no task of their own, but the cascade optimum is still unique: the floating-base rows couple them to the contact forces through the
    6x6 block M[base, arm] (condition number ~3e3), so level 2 spends them - up to the arm torque limits, which end up ACTIVE (7-19 rows in the working
    set, arm accelerations of 1e4 rad/s^2) - on pulling F towards the MPC's forces.  What QMMpcController consumes are the 12 LEG torques
    (QMController.cpp:427-431); they do not see the arm accelerations directly (M[leg, arm] = 0: different branches of the tree), only through F and the base.
    Asserted: every block against the oracle, the leg torques among them; and, solver-independently, that the CUDA result is a KKT point of the literal
    level problems of HoQp.cpp:53-124 built by the numpy twin of the task formulators (tests/test_wbc_twin_cpu.py)."""
    import qm_control_b200 as q
    from qm_control_b200 import synthetic
    import test_wbc_twin_cpu as tw
    B = 96; solver = q.Solver(batch=B, wbc_variant=1); ids, prob, wbc = _mpc_variant_case(B)
    x_des, u_des, mode = synthetic.nominal_wbc_inputs(prob, solver.robot_mass)
    u_des = u_des + synthetic.uniform(77, ids, 1, 30, -1.0, 1.0) * np.r_[np.full(12, 5.0), np.full(18, 0.2)]
    il = u_des + synthetic.uniform(78, ids, 2, 30, -0.002, 0.002); tarr = np.full(B, 12.0)
    solver.wbc_set_input_last(il)
    cmd, status = solver.wbc_update(x_des, u_des, wbc["rbd"], mode, wbc["period"], tarr)
    ref, _ = oracle.wbc_update_batch(x_des, u_des, wbc["rbd"], mode, wbc["period"], tarr, il, variant=1, nthreads=8)
    assert np.all(status == 0)
    # (1) against the oracle, per block; the arm accelerations are O(1e4) and enter through a 3e3-conditioned 6x6 solve: their own floor is their size
    blocks = dict(CMD_BLOCKS); blocks["arm_acc"] = (18, 24, 1e3)
    assert_cmd(cmd, ref, MPCWBC_TOL, tag="wbc HierarchicalMpcWbc vs oracle", blocks=blocks)
    # (2) KKT certificate of the CUDA result for a sample of robots (NNLS multipliers, no solver shared with either side)
    g = tw._gains(); worst = 0.0
    for b in range(0, B, 8):
        dbg = oracle.wbc_debug(x_des[b], u_des[b], wbc["rbd"][b], int(mode[b]), wbc["period"][b], 12.0, input_last=il[b], variant=1)
        (A0, b0, D0, f0), (A1, b1), (A2, b2), M = tw._tasks(oracle, dbg, u_des[b], int(mode[b]), 12.0, g)
        A1v = np.r_[A1[:4], A2[12:14]]; b1v = np.r_[b1[:4], b2[12:14]]; A2v = A2[:12]; b2v = b2[:12]      # HierarchicalMpcWbc.cpp:23-28 stacks
        x = cmd[b, :36]
        assert np.max(np.abs(A0 @ x - b0)) < 1e-7 * (1.0 + np.max(np.abs(b0))), (b, "level 0 equalities")   # feasible level 0 (no slack needed in stance)
        viol = D0 @ x - f0; assert viol.max() < 1e-6, (b, "inequalities", viol.max())
        active = viol > -1e-7 * (1.0 + np.abs(f0))
        r1 = tw._certificate(A1v.T @ (A1v @ x - b1v), A0, D0[active]); r2 = tw._certificate(A2v.T @ (A2v @ x - b2v), np.r_[A0, A1v], D0[active])
        worst = max(worst, r1, r2); assert r1 < 1e-9 and r2 < 1e-9, (b, "KKT residual level 1 / 2", r1, r2)   # observed 1e-11 .. 1e-15 (profiles/r02_mpcwbc_certificates.txt)
        tau = M["M"][6:] @ x[:24] - M["Jfoot"].T[6:] @ x[24:] + M["nle"][6:]
        np.testing.assert_allclose(cmd[b, 36:], tau, rtol=1e-10, atol=1e-9)                                 # updateCmd (WbcBase.cpp:548-563)
    print("HierarchicalMpcWbc: worst KKT residual of the CUDA result %.2e" % worst)


def test_wbc_equation_of_motion_and_limits(oracle):
    """Solver-independent properties: floating-base EoM residual, contact constraint, friction pyramid, torque limits."""
    cmd, ref, mode = _run(oracle, config=4, B=128, variant=0, time=12.0)
    from qm_control_b200 import synthetic
    prob, wbc = synthetic.make_batch(np.arange(128), config=4)
    eff = oracle.model_info()["effort"]; lim = np.r_[np.tile(eff[:3], 4), eff[12:]]
    for b in range(0, 128, 8):
        rbd = wbc["rbd"][b]; q = np.r_[rbd[3:6], rbd[0:3], rbd[6:24]]
        z, y = q[3], q[4]; T = np.array([[0, -np.sin(z), np.cos(y) * np.cos(z)], [0, np.cos(z), np.cos(y) * np.sin(z)], [1, 0, -np.sin(y)]])
        v = np.r_[rbd[27:30], np.linalg.solve(T, rbd[24:27]), rbd[30:48]]
        r = oracle.rbd(q, v); x = cmd[b, :36]; tau = cmd[b, 36:]
        res = r["M"] @ x[:24] + r["nle"] - r["Jfoot"].T @ x[24:] - np.r_[np.zeros(6), tau]
        assert np.max(np.abs(res)) < 1e-6
        assert np.all(np.abs(tau) <= lim + 1e-6)
        for f in range(4):
            F = x[24 + 3 * f:27 + 3 * f]
            if (mode[b] >> (3 - f)) & 1:
                assert F[2] >= -1e-7 and abs(F[0]) <= 0.3 * F[2] + 1e-6 and abs(F[1]) <= 0.3 * F[2] + 1e-6
                acc = r["Jfoot"][3 * f:3 * f + 3] @ x[:24] + r["dJfoot"][3 * f:3 * f + 3] @ v
                assert np.max(np.abs(acc)) < 1e-6
            else:
                assert np.max(np.abs(F)) < 1e-9


def test_wbc_batch_one(oracle):
    """B = 1 must work (plugin use, config 1)."""
    _run(oracle, config=1, B=1, variant=0, time=12.0)


def test_wbc_dynamic_reconfigure_matches_an_oracle_built_with_those_gains(tmp_path):
    """WbcBase::dynamicCallback (WbcBase.cpp:69-117): gains replaced at run time through qmb200_wbc_set_gains give the result of a controller
    constructed with those gains; reading them back returns what was set."""
    import qm_control_b200 as q
    from qm_control_b200 import synthetic
    from _oracle import Oracle, URDF, TASK, REFERENCE
    new = dict(kp_swing=300.0, kd_swing=30.0, base_height_kp=350.0, base_height_kd=120.0, kp_base_linear=380.0, kd_base_linear=90.0, kp_base_angular=420.0, kd_base_angular=150.0,
               kp_arm_joint=[3000, 3100, 3200, 3300, 3400, 3500], kd_arm_joint=[60, 61, 62, 63, 64, 65], kp_ee_linear=[2500, 2600, 2700], kd_ee_linear=[70, 71, 72],
               kp_ee_angular=[1500, 1600, 1700], kd_ee_angular=[50, 51, 52])
    lines = ["wbcGains", "{"]
    for k in ("kp_swing", "kd_swing"):
        lines.append("  %s %r" % (k, new[k]))
    lines += ["  baseHeightKp %r" % new["base_height_kp"], "  baseHeightKd %r" % new["base_height_kd"], "  kp_base_linear %r" % new["kp_base_linear"], "  kd_base_linear %r" % new["kd_base_linear"],
              "  kp_base_angular %r" % new["kp_base_angular"], "  kd_base_angular %r" % new["kd_base_angular"]]
    for i in range(6):
        lines += ["  kp_arm_joint_%d %r" % (i + 1, float(new["kp_arm_joint"][i])), "  kd_arm_joint_%d %r" % (i + 1, float(new["kd_arm_joint"][i]))]
    for i, ax in enumerate("xyz"):
        lines += ["  kp_ee_linear_%s %r" % (ax, float(new["kp_ee_linear"][i])), "  kd_ee_linear_%s %r" % (ax, float(new["kd_ee_linear"][i])),
                  "  kp_ee_angular_%s %r" % (ax, float(new["kp_ee_angular"][i])), "  kd_ee_angular_%s %r" % (ax, float(new["kd_ee_angular"][i]))]
    lines.append("}"); gfile = tmp_path / "gains.info"; gfile.write_text("\n".join(lines) + "\n")
    orc = Oracle(URDF, TASK, REFERENCE, str(gfile))
    B = 48; solver = q.Solver(batch=B); prob, wbc = synthetic.make_batch(np.arange(B), config=5)
    x_des, u_des, mode = synthetic.nominal_wbc_inputs(prob, solver.robot_mass)
    u_des = u_des + synthetic.uniform(77, np.arange(B), 1, 30, -1.0, 1.0) * np.r_[np.full(12, 5.0), np.full(18, 0.2)]
    for b in range(B):
        for f in range(4):
            if not (mode[b] >> (3 - f)) & 1:
                u_des[b, 3 * f:3 * f + 3] = 0.0
    il = synthetic.uniform(78, np.arange(B), 2, 30, -0.1, 0.1); tarr = np.full(B, 12.0)
    solver.wbc_set_input_last(il); before, _ = solver.wbc_update(x_des, u_des, wbc["rbd"], mode, wbc["period"], tarr)
    solver.wbc_set_gains(**new); got = solver.wbc_get_gains()
    for k, v in new.items():
        assert np.allclose(got[k], v, rtol=0, atol=0), k
    solver.wbc_set_input_last(il); after, status = solver.wbc_update(x_des, u_des, wbc["rbd"], mode, wbc["period"], tarr); assert np.all(status == 0)
    ref, _ = orc.wbc_update_batch(x_des, u_des, wbc["rbd"], mode, wbc["period"], tarr, il, variant=0, nthreads=8)
    assert_cmd(after, ref, WBC_TOL, tag="wbc dynamic reconfigure")
    assert np.max(np.abs(after - before)) > 1e-3     # the new gains did change the command
