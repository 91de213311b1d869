"""Synthetic batches for the BASELINE.json configs: identical bytes go to the CUDA path and to the
CPU oracle.  Counter-based PRNG = splitmix64 keyed by (seed, robot, stream) so any robot's inputs can be regenerated
independently (sharding across ranks never changes a robot's data)."""
import re

import numpy as np

from . import _lib
from ._lib import EMAX, KMAX, TARGET

MASK = np.uint64(0xFFFFFFFFFFFFFFFF)
SEED0 = 20230221
GAITS = ("stance", "trot", "flying_trot")
MODE = {"STANCE": 15, "FLY": 0}


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & MASK
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & MASK
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & MASK
    return z ^ (z >> np.uint64(31))


def uniform(seed, robot, stream, n, lo, hi):
    """U(lo,hi) draws, shape [len(robot), n]; value depends only on (seed, robot id, stream, index)."""
    robot = np.asarray(robot, dtype=np.uint64)[:, None]
    idx = np.arange(n, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        key = (np.uint64(seed) * np.uint64(0x100000001B3) + robot * np.uint64(0x10001) + np.uint64(stream) * np.uint64(0x1000193)) & MASK
        z = _splitmix64(_splitmix64(key) + idx)
    u = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return lo + (hi - lo) * u


def _info_vector(path, key, n):
    txt = open(path).read()
    m = re.search(r"(?m)^\s*" + re.escape(key) + r"\s*\{(.*?)^\}", txt, re.S)
    vals = np.zeros(n)
    for i, v in re.findall(r"\((\d+),0\)\s+([-+.\deE]+)", m.group(1)):
        vals[int(i)] = float(v)
    return vals


def _gait_template(path, name):
    txt = open(path).read()
    m = re.search(r"(?m)^" + re.escape(name) + r"\s*\{(.*?)^\}", txt, re.S)
    body = m.group(1)
    modes = re.findall(r"\[\d+\]\s+([A-Z_]+)", re.search(r"modeSequence\s*\{(.*?)\}", body, re.S).group(1))
    times = [float(v) for v in re.findall(r"\[\d+\]\s+([-+.\deE]+)", re.search(r"switchingTimes\s*\{(.*?)\}", body, re.S).group(1))]

    def num(s):
        if s in MODE:
            return MODE[s]
        bits = {"LF": 8, "RF": 4, "LH": 2, "RH": 1}
        return sum(bits[p] for p in s.split("_"))
    return np.array(times), [num(s) for s in modes]


def tile_schedule(times, modes, t_start, lo, hi):
    """Stateless GaitSchedule::getModeSchedule [upstream]: STANCE, t_start, template tiled past `hi`, final STANCE; trim before `lo` keeping one event."""
    ev = [t_start]; md = [15]
    while ev[-1] < hi:
        for i, m in enumerate(modes):
            md.append(m); ev.append(ev[-1] + (times[i + 1] - times[i]))
    md.append(15)
    idx = int(np.searchsorted(ev, lo, side="left"))
    if idx > 0:
        ev = ev[idx - 1:]; md = md[idx - 1:]; md[0] = 15
    return ev, md


def make_batch(robot_ids, config=2, t0=12.0, task_file=None, reference_file=None, gait_file=None, gait=None, horizon=1.0):
    """Inputs for robots `robot_ids` (global ids).  config ∈ {1..5} selects seed and gait assignment:
       1-3 stance, 4 trot, 5 mixed (id mod 3 → stance/trot/flying_trot).  `gait` overrides."""
    task_file = task_file or _lib.asset("qm_task.info"); reference_file = reference_file or _lib.asset("qm_reference.info"); gait_file = gait_file or _lib.asset("qm_gait.info")
    ids = np.asarray(robot_ids, dtype=np.int64); B = len(ids); seed = SEED0 + config
    nominal = _info_vector(task_file, "initialState", 30); djs = _info_vector(reference_file, "defaultJointState", 18)
    U = lambda stream, n, lo, hi: uniform(seed, ids, stream, n, lo, hi)
    q = np.zeros((B, 24))
    q[:, 0:2] = U(1, 2, -0.05, 0.05); q[:, 2] = 0.4 + U(2, 1, -0.02, 0.02)[:, 0]
    q[:, 3] = U(3, 1, -0.3, 0.3)[:, 0]; q[:, 4:6] = U(4, 2, -0.05, 0.05)
    q[:, 6:24] = nominal[12:30] + U(5, 18, -0.1, 0.1)
    x0 = np.zeros((B, 30)); x0[:, 0:3] = U(6, 3, -0.1, 0.1); x0[:, 3:6] = U(7, 3, -0.05, 0.05); x0[:, 6:30] = q
    rbd = np.zeros((B, 55)); rbd[:, 0:3] = q[:, 3:6]; rbd[:, 3:6] = q[:, 0:3]; rbd[:, 6:24] = q[:, 6:24]
    rbd[:, 24:27] = U(8, 3, -0.1, 0.1); rbd[:, 27:30] = U(9, 3, -0.1, 0.1); rbd[:, 30:48] = U(10, 18, -0.2, 0.2)
    # targets (QmTargetTrajectoriesPublisher_node.cpp:44-113): cmd_vel → two knots {t0, t0 + T}
    vx = U(11, 1, 0.0, 0.3)[:, 0]; wz = U(12, 1, -0.2, 0.2)[:, 0]; yaw = q[:, 3]
    # world-frame commanded velocity: R(zyx of the current base pose) * (vx, 0, 0)
    cy, sy = np.cos(yaw), np.sin(yaw); cp, sp = np.cos(q[:, 4]), np.sin(q[:, 4])
    vel = np.stack([cy * cp * vx, sy * cp * vx, -sp * vx], axis=1)
    tt = np.zeros((B, KMAX)); ts = np.zeros((B, KMAX, TARGET)); tt[:, 0] = t0; tt[:, 1] = t0 + horizon
    ee = np.zeros((B, 7)); ee[:, 0] = q[:, 0] + 0.52; ee[:, 1] = q[:, 1] + 0.09; ee[:, 2] = q[:, 2] + 0.036; ee[:, 3:7] = [0.5, -0.5, 0.5, -0.5]   # QMController.cpp:106-108
    for k in range(2):
        ts[:, k, 0:3] = vel; ts[:, k, 6:8] = q[:, 0:2] + (vel[:, 0:2] * horizon if k == 1 else 0.0); ts[:, k, 8] = 0.4
        ts[:, k, 9] = yaw + (wz * horizon if k == 1 else 0.0); ts[:, k, 12:30] = djs; ts[:, k, 30:37] = ee
    # mode schedules
    ev = np.zeros((B, EMAX)); md = np.full((B, EMAX + 1), 15, dtype=np.int32); ne = np.zeros(B, dtype=np.int32)
    phase = U(13, 1, 0.0, 1.0)[:, 0]
    tmpl = {g: _gait_template(gait_file, g) for g in GAITS}
    for b in range(B):
        g = gait or ("stance" if config <= 3 else ("trot" if config == 4 else GAITS[int(ids[b]) % 3]))
        times, modes = tmpl[g]; dur = times[-1]
        t_start = t0 - 2.0 * horizon - phase[b] * dur
        e, m = tile_schedule(times, modes, t_start, t0 - horizon, t0 + 2.0 * horizon)
        if len(e) > EMAX:
            raise ValueError("schedule window needs more than EMAX events")
        ne[b] = len(e); ev[b, :len(e)] = e; md[b, :len(m)] = m
    prob = dict(t0=np.full(B, float(t0)), x0=x0, n_events=ne, event_times=ev, modes=md, n_target=np.full(B, 2, dtype=np.int32), target_times=tt, target_states=ts)
    wbc = dict(rbd=rbd, period=np.full(B, 0.002), time=np.full(B, float(t0)))
    return prob, wbc


def nominal_wbc_inputs(prob, robot_mass):
    """WBC-only inputs (config 3): x_des = x0, u_des = weight-compensating forces for the mode at t0."""
    B = prob["x0"].shape[0]; u = np.zeros((B, 30)); mode = np.zeros(B, dtype=np.int32)
    for b in range(B):
        n = prob["n_events"][b]; idx = int(np.searchsorted(prob["event_times"][b, :n], prob["t0"][b], side="left")); mode[b] = prob["modes"][b, idx]
        flags = [(mode[b] >> (3 - f)) & 1 for f in range(4)]; nc = sum(flags)
        for f in range(4):
            if flags[f]:
                u[b, 3 * f + 2] = robot_mass * 9.81 / nc
    return prob["x0"].copy(), u, mode
