"""Status words are pure flag sets (include/qmb200.h: WBC flags in bits 0..7, MPC flags << 8 in the merged tick word, QMB200_ST_SAFETY = bit 16); the WBC's
iteration counts live in qmb200_wbc_get_diagnostics.  A forced QMB200_ST_ITER_CAP must decode as exactly that on every path that merges words
(round 1 OR-ed iteration counts into bits 8..31 of a failed robot: a level-2 count of 1 read as QMB200_ST_SAFETY, i.e. "stop the controller")."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ST_ITER_CAP, ST_SAFETY, MPC_SHIFT = 1, 0x10000, 8


def test_forced_wbc_iteration_cap_decodes_cleanly_on_every_path():
    import qm_control_b200 as q
    from qm_control_b200 import synthetic
    B = 64; ctrl = q.QMController(batch=B, dt=0.015); s = ctrl.solver
    prob, wbc = synthetic.make_batch(np.arange(B), config=4)
    x_des, u_des, mode = synthetic.nominal_wbc_inputs(prob, s.robot_mass)
    u_des = u_des + synthetic.uniform(77, np.arange(B), 1, 30, -1.0, 1.0) * np.r_[np.full(12, 5.0), np.full(18, 0.2)]
    # reference run with the default caps: clean status, diagnostics separate
    cmd0, st0 = s.wbc_update(x_des, u_des, wbc["rbd"], mode, wbc["period"], np.full(B, 12.0)); d0 = s.wbc_get_diagnostics()
    assert np.all(st0 == 0) and np.all(d0["level0_passes"] >= 1) and np.all(d0["level1_iterations"] >= 1)
    needs_more = d0["level1_iterations"] > 1; assert needs_more.any()
    # cap the active set at ONE iteration: robots that needed more carry ITER_CAP and nothing else
    s.wbc_set_iteration_caps(0, 1)
    cmd1, st1 = s.wbc_update(x_des, u_des, wbc["rbd"], mode, wbc["period"], np.full(B, 12.0))
    assert set(np.unique(st1)) <= {0, ST_ITER_CAP} and np.all((st1 == ST_ITER_CAP)[needs_more]), np.unique(st1)
    assert np.array_equal(cmd1[st1 == 0], cmd0[st1 == 0])
    # qmb200_tick: merged word = WBC byte | MPC flags << 8
    cmd, st = s.tick(prob, prob["t0"] + 0.002, wbc["rbd"], wbc["period"])
    assert np.all((st & 0xFF & ~ST_ITER_CAP) == 0) and np.any(st & ST_ITER_CAP), np.unique(st & 0xFF)
    assert np.all(((st >> MPC_SHIFT) & ~16) == 0), np.unique(st >> MPC_SHIFT)               # MPC side: at most NO_STEP; no iteration counts leaking into bits 8..31
    # qmb200_update: WBC byte | SAFETY; a capped QP must not read as a safety stop
    ctrl.starting(wbc["rbd"], time=12.0)
    cmd, status = ctrl.update(wbc["rbd"], 0.002)
    assert np.all((status & ~(ST_ITER_CAP | ST_SAFETY)) == 0), np.unique(status)
    s.wbc_set_iteration_caps(30, 80)
    cmd, status2 = ctrl.update(wbc["rbd"], 0.002)
    assert np.array_equal(status & ST_SAFETY, status2 & ST_SAFETY) and np.all((status2 & 0xFF) == 0)   # the safety verdict never depended on the QP's iteration counts


def test_out_of_range_counts_are_rejected_or_flagged():
    """qmb200_mpc_solve / qmb200_tick (host pointers) reject n_events / n_target outside [0, EMAX] / [1, KMAX]; the _dev entry points clamp on the device and
    flag QMB200_ST_OVERFLOW instead of reading out of bounds."""
    import torch
    import qm_control_b200 as q
    from qm_control_b200 import synthetic
    B = 4; s = q.Solver(batch=B, dt=0.015); prob, wbc = synthetic.make_batch(np.arange(B), config=4)
    for key, val in (("n_events", 33), ("n_events", -1), ("n_target", 0), ("n_target", 5)):
        bad = dict(prob); bad[key] = prob[key].copy(); bad[key][2] = val
        with pytest.raises(q.QmbError, match=key):
            s.mpc_solve(bad)
    good = s.mpc_solve(prob); assert np.all((good["status"] & ~16) == 0)
    bad = dict(prob); bad["n_events"] = prob["n_events"].copy(); bad["n_events"][1] = 1000; bad["n_target"] = prob["n_target"].copy(); bad["n_target"][3] = 77
    dev = torch.device("cuda", 0); keys = ("t0", "x0", "n_events", "event_times", "modes", "n_target", "target_times", "target_states")
    pdev = {k: torch.from_numpy(np.ascontiguousarray(bad[k])).to(dev) for k in keys}
    s.mpc_reset(); s.mpc_solve_dev(pdev); torch.cuda.synchronize(); sol = s.mpc_get_solution()
    assert sol["status"][1] & 2 and sol["status"][3] & 2, sol["status"]
    assert np.all((sol["status"][[0, 2]] & ~16) == 0)
    for b in (0, 2):                                                                          # the well-formed robots are untouched by their neighbours
        n = int(good["n_nodes"][b]); assert np.array_equal(sol["x"][b, :n], good["x"][b, :n])
