"""Regenerates the golden fixtures in this directory from the CPU oracle (oracle/_build/liboracle.so).

The reference itself cannot run in this container (ROS/OCS2/Pinocchio/qpOASES absent), so these vectors are
outputs of the oracle at the commit that wrote them: they pin the oracle against silent drift and give the CUDA path a
fixture that does not depend on building the oracle.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _oracle import Oracle, TargetOracle  # noqa: E402
from qm_control_b200 import synthetic  # noqa: E402

WBC_IDS = np.array([0, 1, 2, 3, 4, 5, 23, 32, 44])       # config 5: stance / trot / flying trot
MPC_IDS = np.array([0, 1, 2])                            # config 5, dt 0.015 (reference grid), two ticks (cold, warm)
NMAX = 88                                                # node capacity of a dt = 0.015 handle: ceil(1/0.015) + 1 + 20


def wbc_inputs(ids, mass):
    prob, wbc = synthetic.make_batch(ids, config=5)
    x_des, u_des, mode = synthetic.nominal_wbc_inputs(prob, mass)
    u_des = u_des + synthetic.uniform(77, ids, 1, 30, -1.0, 1.0) * np.r_[np.full(12, 5.0), np.full(18, 0.2)]
    for b in range(len(ids)):
        for f in range(4):
            if not (mode[b] >> (3 - f)) & 1:
                u_des[b, 3 * f:3 * f + 3] = 0.0
    il = synthetic.uniform(78, ids, 2, 30, -0.1, 0.1)
    return x_des, u_des, mode, wbc, il


WBC_MPC_IDS = np.arange(16)


def wbc_mpc_inputs(ids, mass):
    prob, wbc = synthetic.make_batch(ids, config=3)
    x_des, u_des, mode = synthetic.nominal_wbc_inputs(prob, mass)
    u_des = u_des + synthetic.uniform(77, ids, 1, 30, -1.0, 1.0) * np.r_[np.full(12, 5.0), np.full(18, 0.2)]
    il = u_des + synthetic.uniform(78, ids, 2, 30, -0.002, 0.002)
    return x_des, u_des, mode, wbc, il


def main():
    o = Oracle(); mass = o.model_info()["mass"]
    x_des, u_des, mode, wbc, il = wbc_inputs(WBC_IDS, mass)
    cmd, il_out = o.wbc_update_batch(x_des, u_des, wbc["rbd"], mode, wbc["period"], np.full(len(WBC_IDS), 12.0), il, variant=0, nthreads=4)
    np.savez_compressed(os.path.join(HERE, "wbc_config5.npz"), ids=WBC_IDS, mode=mode, cmd=cmd, input_last=il_out)
    # HierarchicalMpcWbc (QMMpcController's WBC): stance, realistic joint accelerations; the 12 leg torques cmd[:, 36:48] are what the controller consumes
    x_des, u_des, mode, wbc, il = wbc_mpc_inputs(WBC_MPC_IDS, mass)
    cmd, il_out = o.wbc_update_batch(x_des, u_des, wbc["rbd"], mode, wbc["period"], np.full(len(WBC_MPC_IDS), 12.0), il, variant=1, nthreads=4)
    np.savez_compressed(os.path.join(HERE, "wbc_mpc_variant_config3.npz"), ids=WBC_MPC_IDS, mode=mode, cmd=cmd, input_last=il_out)

    o.mpc_set(dt=0.015, horizon=1.0)
    prob, _ = synthetic.make_batch(MPC_IDS, config=5)
    t1 = o.mpc_solve_batch(prob, NMAX, nthreads=3)
    prob2 = dict(prob); prob2["t0"] = prob["t0"] + 0.01
    x0 = np.zeros((len(MPC_IDS), 30))
    for b in range(len(MPC_IDS)):
        n = t1["n_nodes"][b]; ne = prob["n_events"][b]
        x0[b], _, _ = o.evaluate_policy(t1["t"][b, :n], t1["event"][b, :n], t1["x"][b, :n], t1["u"][b, :n], prob["event_times"][b, :ne], prob["modes"][b, :ne + 1], prob2["t0"][b])
    prob2["x0"] = x0
    t2 = o.mpc_solve_batch(prob2, NMAX, prev=t1, nthreads=3)
    np.savez_compressed(os.path.join(HERE, "mpc_config5_dt015.npz"), ids=MPC_IDS, nmax=NMAX,
                        **{"t1_" + k: t1[k] for k in ("n_nodes", "t", "event", "x", "u", "dbg")}, x0_tick2=x0, **{"t2_" + k: t2[k] for k in ("n_nodes", "t", "event", "x", "u", "dbg")})

    to = TargetOracle(); prob, wbc = synthetic.make_batch(np.arange(4), config=4); ee = np.tile([0.6, 0.1, 0.45, 0.5, -0.5, 0.5, -0.5], (4, 1)); ee[:, 0] += 0.05 * np.arange(4)
    last = np.tile([0.52, 0.09, 0.44, 0.5, -0.5, 0.5, -0.5], (4, 1)); cmds = np.array([[0.3, 0.0, 0.0, 0.2, 0, 0, 0], [0.1, -0.1, 0.05, 0, 0, 0, 0], [0.7, 0.2, 0.5, 0.0, 0.0, 0.0, 1.0]])
    out = {}
    for kind in range(3):
        res = [to.target(kind, cmds[kind], 12.0, prob["x0"][b], ee[b], last[b]) for b in range(4)]
        out["times_%d" % kind] = np.stack([r[0] for r in res]); out["states_%d" % kind] = np.stack([r[1] for r in res]); out["last_%d" % kind] = np.stack([r[2] for r in res])
    np.savez_compressed(os.path.join(HERE, "target_config4.npz"), ee=ee, last=last, cmds=cmds, **out)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
