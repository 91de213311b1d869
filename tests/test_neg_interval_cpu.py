"""Root cause of the one flagged robot of the bench workload (synthetic robot 1758, trot, dt = 0.01): NOT cancellation in the Riccati recursion and NOT a
singular stance Jacobian, but a time-grid interval with a NEGATIVE duration.

timeDiscretizationWithEvents [upstream ocs2_oc/oc_data/TimeDiscretization.cpp, recalled] merges a node into its predecessor only when they are closer than
dt_min = 10 * limitEpsilon = 1e-8 s, while getIntervalEnd / getIntervalStart shift pre-/post-event nodes by weakEpsilon = 1e-6 s.  An event that falls between
1e-8 and 1e-6 s after a grid node therefore produces an interval whose duration getIntervalEnd - getIntervalStart is negative: the stage cost and the RK2 step
of that interval are scaled by dt < 0, R*dt is NEGATIVE definite and the QP is genuinely non-convex - every exact solver must reject it (HPIPM's Cholesky of
R + B'PB hits non-positive pivots).  The schedule of robot 1758 has its first event 6.8e-7 s after the node t0 + 14*dt (probability ~1e-6/dt per event and tick).
Both sides report it instead of returning a step: the oracle raises, the CUDA path sets QMB200_ST_NEG_DT | QMB200_ST_NOT_PD | QMB200_ST_NO_STEP
(tests/test_mpc_gpu.py::test_not_positive_definite_is_reported_like_the_oracle)."""
import numpy as np
import pytest

from qm_control_b200 import synthetic

NMAX = 121


def test_robot_1758_has_an_interval_of_negative_duration(oracle):
    oracle.mpc_set(dt=0.01, horizon=1.0); prob, _ = synthetic.make_batch(np.array([1758]), config=4, horizon=1.0)
    ne = int(prob["n_events"][0]); ev = prob["event_times"][0, :ne]; t0 = float(prob["t0"][0]); grid = t0 + 0.01 * np.arange(101)
    inside = ev[(ev > t0) & (ev < t0 + 1.0)]; off = np.array([e - grid[np.searchsorted(grid, e) - 1] for e in inside])
    assert len(inside) == 3 and np.all(off > 1e-8) and np.all(off < 1e-6), off           # later than dt_min (no merge), earlier than weakEpsilon
    qp = oracle.mpc_qp(prob, NMAX); N = qp["n_nodes"] - 1                                 # the stage blocks are exported although the sweep rejects the QP
    assert np.all(qp["dx"] == 0.0) and np.all(qp["du"] == 0.0)
    R_ok = qp["R"][0]; assert np.linalg.eigvalsh(R_ok).min() > 0
    neg = [k for k in range(N) if not qp["is_event"][k] and np.linalg.eigvalsh(qp["R"][k]).max() < 0]
    assert neg == [14], neg                                                               # the first event; after it the grid restarts AT the event, so the later ones (one template period apart) coincide with nodes and are merged
    for k in neg:
        assert qp["is_event"][k + 1] == 1                                                 # the interval that ends in the pre-event node
        ratio = qp["R"][k][24, 24] / R_ok[24, 24]                                         # arm-joint weight: R*dt with the same R => ratio of the durations
        dur = ratio * 0.01; assert abs(dur - (off[0] - 1e-6)) < 1e-9, (dur, off[0] - 1e-6)  # = (t_event - weakEpsilon) - t_node < 0
    # every other interval is fine: with that stage left out the backward sweep in plain numpy (float64) keeps H = R~ + B~'P B~ positive definite
    P = qp["QN"].copy(); hmin = np.inf
    for k in range(N - 1, -1, -1):
        if qp["is_event"][k] or k in neg:
            continue
        ng = qp["ng"][k]; Ck, Dk = qp["C"][k, :ng], qp["D"][k, :ng]; A, B, Q, R, S = qp["A"][k], qp["B"][k], qp["Q"][k], qp["R"][k], qp["P"][k]
        _, sv, Vt = np.linalg.svd(Dk); assert sv.min() > 1e-3; Z = Vt[ng:].T; Px = -np.linalg.pinv(Dk) @ Ck
        At = A + B @ Px; Bt = B @ Z; Qt = Q + Px.T @ S + S.T @ Px + Px.T @ R @ Px; Rt = Z.T @ R @ Z; St = Z.T @ (S + R @ Px)
        H = Rt + Bt.T @ P @ Bt; H = 0.5 * (H + H.T); G = St + Bt.T @ P @ At; hmin = min(hmin, np.linalg.eigvalsh(H).min())
        P = Qt + At.T @ P @ At - G.T @ np.linalg.solve(H, G); P = 0.5 * (P + P.T)
    assert hmin > 1e-5, hmin
    with pytest.raises(RuntimeError, match="not positive definite"):
        oracle.mpc_solve_batch(prob, NMAX, nthreads=1)


def test_neighbour_robots_have_no_such_interval(oracle):
    oracle.mpc_set(dt=0.01, horizon=1.0); prob, _ = synthetic.make_batch(np.array([1757, 1759]), config=4, horizon=1.0)
    out = oracle.mpc_solve_batch(prob, NMAX, nthreads=2); assert np.all(out["dbg"][:, 0] > 0)
