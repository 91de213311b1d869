"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
ASSETS = os.path.join(ROOT, "assets")
URDF = os.path.join(ASSETS, "qm_robot.urdf")
TASK = os.path.join(ASSETS, "qm_task.info")
REFERENCE = os.path.join(ASSETS, "qm_reference.info")
GAIT = os.path.join(ASSETS, "qm_gait.info")
GAINS = os.path.join(ASSETS, "qm_wbc_gains.info")

EMAX, KMAX = 32, 4

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def _d(a):
    return a.ctypes.data_as(dp)


def _i(a):
    return a.ctypes.data_as(ip)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Oracle:
    def __init__(self, urdf=URDF, task=TASK, reference=REFERENCE, gains=GAINS):
        if not os.path.exists(LIB_PATH):
            build_oracle()
        self.lib = C.CDLL(LIB_PATH)
        self.lib.orc_create.restype = C.c_void_p
        self.lib.orc_last_error.restype = C.c_char_p
        self.h = self.lib.orc_create(urdf.encode(), task.encode(), reference.encode(), gains.encode() if gains else None)
        if not self.h:
            raise RuntimeError("oracle: " + self.lib.orc_last_error().decode())
        self.h = C.c_void_p(self.h)

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("oracle: " + self.lib.orc_last_error().decode())

    def __del__(self):
        try:
            self.lib.orc_destroy(self.h)
        except Exception:
            pass

    def model_info(self):
        mass = C.c_double()
        I = np.zeros(9); c = np.zeros(3); q = np.zeros(24); eff = np.zeros(18); lo = np.zeros(18); hi = np.zeros(18)
        self.lib.orc_model_info(self.h, C.byref(mass), _d(I), _d(c), _d(q), _d(eff), _d(lo), _d(hi))
        names = []
        for j in range(18):
            buf = C.create_string_buffer(64)
            self.lib.orc_joint_name(self.h, j, buf, 64)
            names.append(buf.value.decode())
        return dict(mass=mass.value, inertia_nominal=I.reshape(3, 3), com_to_base=c, q_nominal=q, effort=eff, lower=lo, upper=hi, joint_names=names)

    def rbd(self, q, v):
        q = f64(q); v = f64(v)
        o = dict(M=np.zeros((24, 24)), nle=np.zeros(24), Jfoot=np.zeros((12, 24)), dJfoot=np.zeros((12, 24)), Jbase=np.zeros((6, 24)), dJbase=np.zeros((6, 24)),
                 Jee=np.zeros((6, 24)), dJee=np.zeros((6, 24)), Ag=np.zeros((6, 24)), dAg_v=np.zeros(6), com=np.zeros(3), foot_pos=np.zeros((4, 3)), foot_vel=np.zeros((4, 3)),
                 ee_pos=np.zeros(3), ee_rot=np.zeros((3, 3)))
        self._chk(self.lib.orc_rbd(self.h, _d(q), _d(v), *[_d(o[k]) for k in ("M", "nle", "Jfoot", "dJfoot", "Jbase", "dJbase", "Jee", "dJee", "Ag", "dAg_v", "com", "foot_pos", "foot_vel", "ee_pos", "ee_rot")]))
        return o

    def centroidal_state_from_rbd(self, rbd):
        rbd = f64(rbd); x = np.zeros(30)
        self._chk(self.lib.orc_centroidal_state_from_rbd(self.h, _d(rbd), _d(x)))
        return x

    def wbc_update(self, x_des, u_des, rbd, mode, period, time, input_last=None, variant=0):
        x_des = f64(x_des); u_des = f64(u_des); rbd = f64(rbd)
        il = np.zeros(30) if input_last is None else f64(input_last).copy()
        cmd = np.zeros(54); iters = np.zeros(3, dtype=np.int32)
        self._chk(self.lib.orc_wbc_update(self.h, _d(x_des), _d(u_des), _d(rbd), C.c_int(mode), C.c_double(period), C.c_double(time), _d(il), C.c_int(variant), _d(cmd), _i(iters)))
        return cmd, il, iters

    def wbc_debug(self, x_des, u_des, rbd, mode, period, time, input_last=None, variant=0):
        x_des = f64(x_des); u_des = f64(u_des); rbd = f64(rbd)
        il = np.zeros(30) if input_last is None else f64(input_last)
        o = dict(q_meas=np.zeros(24), v_meas=np.zeros(24), q_des=np.zeros(24), v_des=np.zeros(24), base_acc=np.zeros(6), levels=np.zeros((3, 36)))
        self._chk(self.lib.orc_wbc_debug(self.h, _d(x_des), _d(u_des), _d(rbd), C.c_int(mode), C.c_double(period), C.c_double(time), _d(il), C.c_int(variant),
                                         *[_d(o[k]) for k in ("q_meas", "v_meas", "q_des", "v_des", "base_acc", "levels")]))
        return o

    def wbc_update_batch(self, x_des, u_des, rbd, mode, period, time, input_last, variant=0, nthreads=1):
        x_des = f64(x_des); u_des = f64(u_des); rbd = f64(rbd); mode = i32(mode); period = f64(period); time = f64(time)
        B = x_des.shape[0]; il = f64(input_last).copy(); cmd = np.zeros((B, 54))
        self._chk(self.lib.orc_wbc_update_batch(self.h, C.c_int(B), _d(x_des), _d(u_des), _d(rbd), _i(mode), _d(period), _d(time), _d(il), C.c_int(variant), _d(cmd), C.c_int(nthreads)))
        return cmd, il

    # ---- MPC ----
    def mpc_set(self, dt=-1.0, horizon=-1.0, rk=(-1.0, 0.0, 0.0)):
        self.lib.orc_mpc_set(self.h, C.c_double(dt), C.c_double(horizon), C.c_double(rk[0]), C.c_double(rk[1]), C.c_double(rk[2]))

    def mpc_set_sqp(self, sqp_iterations=0, cost_tol=0.0):
        self.lib.orc_mpc_set_sqp(self.h, C.c_int(int(sqp_iterations)), C.c_double(float(cost_tol)))

    def mpc_set_solver(self, solver=-1, iterations=0, delta_tol=0.0, g_max=0.0, g_min=0.0, ddp_penalty=0.0, ddp_min_step=0.0, ddp_max_step=0.0):
        """solver: 0 SQP, 1 IPM, 2 DDP; tolerances <= 0 keep their value (the tests read the task file's block themselves)."""
        self.lib.orc_mpc_set_solver(self.h, C.c_int(int(solver)), C.c_int(int(iterations)), *[C.c_double(float(v)) for v in (delta_tol, g_max, g_min, ddp_penalty, ddp_min_step, ddp_max_step)])

    def mpc_weights(self):
        Q = np.zeros((30, 30)); R = np.zeros((30, 30))
        self.lib.orc_mpc_get_weights(self.h, _d(Q), _d(R))
        return Q, R

    def flow_map(self, x, u):
        x = f64(x); u = f64(u); f = np.zeros(30); A = np.zeros((30, 30)); B = np.zeros((30, 30))
        self._chk(self.lib.orc_flow_map(self.h, _d(x), _d(u), _d(f), _d(A), _d(B)))
        return f, A, B

    def mpc_solve_batch(self, prob, nmax, prev=None, nthreads=1, want_dbg=True):
        """prob: dict(t0[B], x0[B,30], n_events[B], event_times[B,EMAX], modes[B,EMAX+1], n_target[B], target_times[B,KMAX], target_states[B,KMAX,37]);
        prev: dict(n_nodes[B], t[B,nmax], event[B,nmax], x[B,nmax,30], u[B,nmax,30]) or None."""
        B = prob["t0"].shape[0]
        out = dict(n_nodes=np.zeros(B, dtype=np.int32), t=np.zeros((B, nmax)), event=np.zeros((B, nmax), dtype=np.int32), x=np.zeros((B, nmax, 30)), u=np.zeros((B, nmax, 30)))
        dbg = np.zeros((B, 11)) if want_dbg else None
        pa = [None] * 5 if prev is None else [_i(i32(prev["n_nodes"])), _d(f64(prev["t"])), _i(i32(prev["event"])), _d(f64(prev["x"])), _d(f64(prev["u"]))]
        keep = [f64(prob["t0"]), f64(prob["x0"]), i32(prob["n_events"]), f64(prob["event_times"]), i32(prob["modes"]), i32(prob["n_target"]), f64(prob["target_times"]), f64(prob["target_states"])]
        emax = keep[3].shape[1]; kmax = keep[6].shape[1]
        self._chk(self.lib.orc_mpc_solve_batch(self.h, C.c_int(B), C.c_int(emax), C.c_int(kmax), C.c_int(nmax), _d(keep[0]), _d(keep[1]), _i(keep[2]), _d(keep[3]), _i(keep[4]), _i(keep[5]), _d(keep[6]), _d(keep[7]),
                                               *pa, _i(out["n_nodes"]), _d(out["t"]), _i(out["event"]), _d(out["x"]), _d(out["u"]), _d(dbg) if want_dbg else None, C.c_int(nthreads)))
        if want_dbg:
            out["dbg"] = dbg
        return out

    def tick_batch(self, prob, nmax, t_eval, rbd, period, input_last, prev=None, variant=0, nthreads=1):
        """mpc_solve → evaluatePolicy(t_eval) → wbc update per robot (thread pool) — the CPU baseline of bench.py."""
        B = prob["t0"].shape[0]
        out = dict(n_nodes=np.zeros(B, dtype=np.int32), t=np.zeros((B, nmax)), event=np.zeros((B, nmax), dtype=np.int32), x=np.zeros((B, nmax, 30)), u=np.zeros((B, nmax, 30)), cmd=np.zeros((B, 54)))
        pa = [None] * 5 if prev is None else [_i(i32(prev["n_nodes"])), _d(f64(prev["t"])), _i(i32(prev["event"])), _d(f64(prev["x"])), _d(f64(prev["u"]))]
        keep = [f64(prob["t0"]), f64(prob["x0"]), i32(prob["n_events"]), f64(prob["event_times"]), i32(prob["modes"]), i32(prob["n_target"]), f64(prob["target_times"]), f64(prob["target_states"])]
        te = f64(t_eval); rb = f64(rbd); pe = f64(period); il = f64(input_last).copy(); emax = keep[3].shape[1]; kmax = keep[6].shape[1]
        self._chk(self.lib.orc_tick_batch(self.h, C.c_int(B), C.c_int(emax), C.c_int(kmax), C.c_int(nmax), _d(keep[0]), _d(keep[1]), _i(keep[2]), _d(keep[3]), _i(keep[4]), _i(keep[5]), _d(keep[6]), _d(keep[7]),
                                          *pa, _d(te), _d(rb), _d(pe), _d(il), C.c_int(variant), _i(out["n_nodes"]), _d(out["t"]), _i(out["event"]), _d(out["x"]), _d(out["u"]), _d(out["cmd"]), C.c_int(nthreads)))
        out["input_last"] = il
        return out

    def mpc_debug(self, prob, nmax, prev=None, max_k=200):
        keep = [f64(prob["t0"][:1]), f64(prob["x0"][:1]), i32(prob["n_events"][:1]), f64(prob["event_times"][:1]), i32(prob["modes"][:1]), i32(prob["n_target"][:1]), f64(prob["target_times"][:1]), f64(prob["target_states"][:1])]
        emax = keep[3].shape[1]; kmax = keep[6].shape[1]
        pa = [None] * 5 if prev is None else [_i(i32(prev["n_nodes"][:1])), _d(f64(prev["t"][:1])), _i(i32(prev["event"][:1])), _d(f64(prev["x"][:1])), _d(f64(prev["u"][:1]))]
        A = np.zeros((max_k, 30, 30)); Bm = np.zeros((max_k, 30, 30)); b = np.zeros((max_k, 30)); dx = np.zeros((max_k + 1, 30)); du = np.zeros((max_k, 30)); n = C.c_int()
        self._chk(self.lib.orc_mpc_debug(self.h, C.c_int(emax), C.c_int(kmax), C.c_int(nmax), _d(keep[0]), _d(keep[1]), _i(keep[2]), _d(keep[3]), _i(keep[4]), _i(keep[5]), _d(keep[6]), _d(keep[7]), *pa,
                                         C.c_int(max_k), _d(A), _d(Bm), _d(b), _d(dx), _d(du), C.byref(n)))
        return dict(A=A, B=Bm, b=b, dx=dx, du=du, n_nodes=n.value)

    def evaluate_policy(self, t, event, x, u, event_times, modes, tq):
        t = f64(t); event = i32(event); x = f64(x); u = f64(u); et = f64(event_times); md = i32(modes)
        xd = np.zeros(30); ud = np.zeros(30); mode = C.c_int()
        self._chk(self.lib.orc_evaluate_policy(self.h, C.c_int(len(t)), _d(t), _i(event), _d(x), _d(u), C.c_int(len(et)), _d(et), _i(md), C.c_double(tq), _d(xd), _d(ud), C.byref(mode)))
        return xd, ud, mode.value

    def swing_reference(self, event_times, modes, leg, t):
        et = f64(event_times); md = i32(modes); zp = C.c_double(); zv = C.c_double()
        self._chk(self.lib.orc_swing_reference(self.h, C.c_int(len(et)), _d(et), _i(md), C.c_int(leg), C.c_double(t), C.byref(zp), C.byref(zv)))
        return zp.value, zv.value

    # ---- controller side (oracle/src/ctrl.cpp) ----
    def observation_update(self, rbd, period, t_obs, x_obs):
        t = C.c_double(float(t_obs)); x = f64(x_obs).copy()
        self._chk(self.lib.orc_observation_update(self.h, _d(f64(rbd)), C.c_double(float(period)), C.byref(t), _d(x)))
        return t.value, x

    def control_law(self, variant, arm_kp, arm_kd, x_des, u_des, wbc_cmd, time, x_obs, joint_cmd, arm_pos, last_time):
        jc = f64(joint_cmd).copy(); ap = f64(arm_pos).copy(); lt = C.c_double(float(last_time))
        safe = self.lib.orc_control_law(C.c_int(variant), C.c_double(arm_kp), C.c_double(arm_kd), _d(f64(x_des)), _d(f64(u_des)), _d(f64(wbc_cmd)), C.c_double(float(time)), _d(f64(x_obs)), _d(jc), _d(ap), C.byref(lt))
        return jc, ap, lt.value, bool(safe)


class TargetOracle:
    """QmTargetTrajectoriesPublisher_node.cpp restated (oracle/src/ctrl.cpp)."""

    def __init__(self, task=TASK, reference=REFERENCE):
        if not os.path.exists(LIB_PATH):
            build_oracle()
        self.lib = C.CDLL(LIB_PATH); self.lib.orc_ctrl_create.restype = C.c_void_p
        self.c = C.c_void_p(self.lib.orc_ctrl_create(task.encode(), reference.encode()))
        assert self.c

    def target(self, kind, cmd, t_obs, x_obs, ee_state, last_ee):
        c7 = np.zeros(7); cmd = f64(cmd); c7[:len(cmd)] = cmd; le = f64(last_ee).copy(); times = np.zeros(2); states = np.zeros((2, 37))
        self.lib.orc_target_trajectories(self.c, C.c_int(kind), _d(c7), C.c_double(float(t_obs)), _d(f64(x_obs)), _d(f64(ee_state)), _d(le), _d(times), _d(states))
        return times, states, le


class HwSimOracle:
    """QMHWSim::writeSim restated with a std::deque (oracle/src/ctrl.cpp)."""

    def __init__(self, delay):
        if not os.path.exists(LIB_PATH):
            build_oracle()
        self.lib = C.CDLL(LIB_PATH); self.lib.orc_hw_create.restype = C.c_void_p
        self.h = C.c_void_p(self.lib.orc_hw_create(C.c_double(delay)))

    def write(self, time, period, joint_cmd, pos, vel):
        eff = np.zeros(18)
        self.lib.orc_hw_write(self.h, C.c_double(float(time)), C.c_double(float(period)), _d(f64(joint_cmd)), _d(f64(pos)), _d(f64(vel)), _d(eff))
        return eff


def _mpc_qp(self, prob, nmax, max_k=200):
    """QP of the tick for robot 0 of `prob` (cold start): dict of per-interval blocks, see orc_mpc_qp."""
    keep = [f64(prob["t0"][:1]), f64(prob["x0"][:1]), i32(prob["n_events"][:1]), f64(prob["event_times"][:1]), i32(prob["modes"][:1]), i32(prob["n_target"][:1]), f64(prob["target_times"][:1]), f64(prob["target_states"][:1])]
    emax = keep[3].shape[1]; kmax = keep[6].shape[1]; K = max_k
    o = dict(A=np.zeros((K, 30, 30)), B=np.zeros((K, 30, 30)), b=np.zeros((K, 30)), Q=np.zeros((K, 30, 30)), R=np.zeros((K, 30, 30)), P=np.zeros((K, 30, 30)), q=np.zeros((K, 30)), r=np.zeros((K, 30)),
             C=np.zeros((K, 16, 30)), D=np.zeros((K, 16, 30)), e=np.zeros((K, 16)), ng=np.zeros(K, dtype=np.int32), is_event=np.zeros(K, dtype=np.int32), QN=np.zeros((30, 30)), qN=np.zeros(30),
             dx=np.zeros((K + 1, 30)), du=np.zeros((K, 30)))
    n = C.c_int()
    self._chk(self.lib.orc_mpc_qp(self.h, C.c_int(emax), C.c_int(kmax), C.c_int(nmax), _d(keep[0]), _d(keep[1]), _i(keep[2]), _d(keep[3]), _i(keep[4]), _i(keep[5]), _d(keep[6]), _d(keep[7]), C.c_int(K),
                                  *[_d(o[k]) for k in ("A", "B", "b", "Q", "R", "P", "q", "r", "C", "D", "e")], _i(o["ng"]), _i(o["is_event"]), _d(o["QN"]), _d(o["qN"]), _d(o["dx"]), _d(o["du"]), C.byref(n)))
    o["n_nodes"] = n.value
    return o


def _stage_probe(self, event_times, modes, target_times, target_states, t, x, u, want_grad=True):
    et = f64(event_times); md = i32(modes); tt = f64(target_times); ts = f64(target_states); f = C.c_double(); q = np.zeros(30); r = np.zeros(30); g = np.zeros(16); ng = C.c_int()
    self._chk(self.lib.orc_stage_probe(self.h, C.c_int(len(et)), _d(et), _i(md), C.c_int(len(tt)), _d(tt), _d(ts), C.c_double(float(t)), _d(f64(x)), _d(f64(u)), C.byref(f),
                                       _d(q) if want_grad else None, _d(r) if want_grad else None, _d(g), C.byref(ng)))
    return f.value, q, r, g[:ng.value].copy()


Oracle.mpc_qp = _mpc_qp
Oracle.stage_probe = _stage_probe
