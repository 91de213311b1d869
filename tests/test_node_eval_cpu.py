"""The thread-per-node evaluator of the product (qm_control_b200/csrc/kernels/node_eval.cuh: what K2a and the line search run, one CUDA thread per node)
compiled for the HOST from the same header (tests/nodeeval_host.cpp) and checked against the oracle: flow map and its Jacobian blocks, cost value,
equality residuals; constraint and end-effector Jacobians against central differences of the evaluator's own values."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from _oracle import ASSETS, GAINS, REFERENCE, ROOT, TASK, URDF, Oracle, _d, _i, f64, i32
from qm_control_b200 import synthetic

NMAX = 100
SRC = os.path.join(ROOT, "tests", "nodeeval_host.cpp")
LIB = os.path.join(ROOT, "tests", "_build", "libnodeeval.so")
CSRC = os.path.join(ROOT, "qm_control_b200", "csrc")


@pytest.fixture(scope="module")
def nev():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-attributes", "-Wno-unknown-pragmas", "-I/usr/local/cuda/include", "-I" + CSRC, "-o", LIB, SRC,
                           os.path.join(CSRC, "host", "qm_config.cpp")])
    lib = C.CDLL(LIB); lib.nev_create.restype = C.c_void_p
    h = lib.nev_create(TASK.encode(), URDF.encode(), REFERENCE.encode(), GAINS.encode()); assert h
    yield lib, C.c_void_p(h)
    lib.nev_destroy(C.c_void_p(h))


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def _flow(nev, x, u):
    lib, h = nev; f = np.zeros(30); A = np.zeros((30, 30)); B = np.zeros((30, 30)); lib.nev_flow(h, _d(f64(x)), _d(f64(u)), _d(f), _d(A), _d(B)); return f, A, B


def _stage(nev, et, md, tt, ts, t, x, u, terminal=0):
    lib, h = nev; et = f64(et); md = i32(md); tt = f64(tt); ts = f64(ts); cost = C.c_double(); eq = C.c_double(); fe = np.zeros((4, 3)); ee = np.zeros(6); Cm = np.zeros((4, 3, 12)); Je = np.zeros((6, 12))
    rc = lib.nev_stage(h, C.c_int(len(et)), _d(et), _i(md), C.c_int(len(tt)), _d(tt), _d(ts), C.c_double(float(t)), _d(f64(x)), _d(f64(u)), C.c_int(terminal), C.byref(cost), C.byref(eq), _d(fe), _d(ee), _d(Cm), _d(Je))
    return dict(rc=rc, cost=cost.value, eq=eq.value, foot_e=fe, ee=ee, C=Cm, Je=Je)


def _rand_point(oracle, seed):
    rng = np.random.default_rng(seed); mi = oracle.model_info()
    x = np.r_[rng.uniform(-0.2, 0.2, 6), 0.05, -0.03, 0.41, rng.uniform(-0.3, 0.3, 3), mi["q_nominal"][6:] + rng.uniform(-0.2, 0.2, 18)]
    u = np.r_[rng.uniform(-20, 20, 12) + np.tile([0, 0, 67.0], 4), rng.uniform(-0.5, 0.5, 18)]
    return x, u


@pytest.mark.parametrize("seed", [3, 4, 5])
def test_flow_map_and_jacobian_blocks_match_the_oracle(nev, oracle, seed):
    x, u = _rand_point(oracle, seed)
    f, A, B = _flow(nev, x, u); fo, Ao, Bo = oracle.flow_map(x, u)
    np.testing.assert_allclose(f, fo, rtol=0, atol=1e-12 * max(1.0, np.max(np.abs(fo))))
    np.testing.assert_allclose(A, Ao, rtol=0, atol=1e-11 * max(1.0, np.max(np.abs(Ao))))
    np.testing.assert_allclose(B, Bo, rtol=0, atol=1e-12 * max(1.0, np.max(np.abs(Bo))))


@pytest.mark.parametrize("config,robot", [(4, 1), (5, 2), (3, 0)])
def test_cost_and_equality_residuals_match_the_oracle_along_a_horizon(nev, oracle, config, robot):
    oracle.mpc_set(dt=0.015, horizon=1.0); prob, _ = synthetic.make_batch(np.array([robot]), config=config)
    sol = oracle.mpc_solve_batch(prob, NMAX, nthreads=1); n = int(sol["n_nodes"][0]); t = sol["t"][0, :n]; ev = sol["event"][0, :n]
    ne = int(prob["n_events"][0]); et = prob["event_times"][0, :ne]; md = prob["modes"][0, :ne + 1]; nk = int(prob["n_target"][0]); tt = prob["target_times"][0, :nk]; ts = prob["target_states"][0, :nk]
    checked = 0
    for k in range(0, n - 1, 5):
        if ev[k] == 1:
            continue
        tk = t[k] + (1e-6 if ev[k] == 2 else 0.0); x = sol["x"][0, k]; u = sol["u"][0, k]
        fo, _, _, go = oracle.stage_probe(et, md, tt, ts, tk, x, u, want_grad=False)
        r = _stage(nev, et, md, tt, ts, tk, x, u); assert r["rc"] == 0
        assert abs(r["cost"] - fo) <= 1e-11 * max(1.0, abs(fo)), (k, r["cost"], fo)
        assert abs(r["eq"] - float(go @ go)) <= 1e-11 * max(1.0, float(go @ go)), (k, r["eq"], float(go @ go))
        checked += 1
    assert checked >= 10


def test_constraint_and_end_effector_jacobians_by_central_differences(nev, oracle):
    oracle.mpc_set(dt=0.015, horizon=1.0); prob, _ = synthetic.make_batch(np.array([2]), config=5)
    ne = int(prob["n_events"][0]); et = prob["event_times"][0, :ne]; md = prob["modes"][0, :ne + 1]; nk = int(prob["n_target"][0]); tt = prob["target_times"][0, :nk]; ts = prob["target_states"][0, :nk]
    x, u = _rand_point(oracle, 11); t = float(prob["t0"][0]) + 0.3; r0 = _stage(nev, et, md, tt, ts, t, x, u); h = 1e-6
    info = oracle.model_info(); foot_leg = [0, 6, 3, 9]   # contact order LF, RF, LH, RH -> first joint of the leg in joint order LF, LH, RF, RH
    for i in range(4):
        cols = list(range(6)) + [9, 10, 11] + [12 + foot_leg[i] + j for j in range(3)]
        for pos, c in enumerate(cols):
            d = np.zeros(30); d[c] = h
            fd = (_stage(nev, et, md, tt, ts, t, x + d, u)["foot_e"][i] - _stage(nev, et, md, tt, ts, t, x - d, u)["foot_e"][i]) / (2 * h)
            np.testing.assert_allclose(r0["C"][i, :, pos], fd, atol=2e-7 * (1.0 + np.max(np.abs(r0["C"][i]))))
    for pos in range(12):
        c = 6 + pos if pos < 6 else 18 + pos; d = np.zeros(30); d[c] = h
        fd = (_stage(nev, et, md, tt, ts, t, x + d, u)["ee"] - _stage(nev, et, md, tt, ts, t, x - d, u)["ee"]) / (2 * h)
        np.testing.assert_allclose(r0["Je"][:, pos], fd, atol=2e-7 * (1.0 + np.max(np.abs(r0["Je"]))))
