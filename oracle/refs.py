"""CPU restatement of the reference's per-solve reference preprocessing. TEST INFRASTRUCTURE ONLY.

Unlike the product's flattened segment lists (csrc/hb_planner.h) this follows the reference's own object structure: one spline
object per *phase index* and per axis, looked up by lower_bound on the event times, so that the two implementations only agree
if both are right.

  GaitSchedule::{insert,tile}ModeSequenceTemplate     legged_interface/src/gait/GaitSchedule.cpp:57-161
  SwingTrajectoryPlanner::update/calNextFootPos/...   legged_interface/src/foot_planner/SwingTrajectoryPlanner.cpp:164-358, 394-458
  CubicSpline / MultiCubicSpline                      legged_interface/src/foot_planner/{CubicSpline,MultiCubicSpline}.cpp
  cmdVelToTargetTrajectories                          legged_controllers/src/TargetTrajectoriesPublisher.cpp:41-130
  constants                                           legged_controllers/config/hunter/{reference,task}.info

parity: pinned only through the reference's published constants and the properties tested in tests/test_planner.py (the
reference has no fixtures for this layer and it cannot be imported: it is C++ against OCS2).
"""
import bisect
import math

import numpy as np

COM_HEIGHT = 0.57
NEXT_Z = 0.02
SWING_HEIGHT = 0.08
SWING_TIME_SCALE = 0.15
FEET_BIAS = [(0.11, 0.12, -0.57), (0.11, -0.12, -0.57), (-0.06, 0.12, -0.57), (-0.06, -0.12, -0.57)]
DEFAULT_JOINTS = None  # filled from the generated header below

GAITS = {
    "stance": ([3], [0.0, 0.5]),
    "trot": ([2, 1], [0.0, 0.3, 0.6]),
    "standing_trot": ([2, 3, 1, 3], [0.0, 0.25, 0.3, 0.55, 0.6]),
    "flying_trot": ([2, 0, 1, 0], [0.0, 0.15, 0.2, 0.35, 0.4]),
}


def _header_array(name):
    import os, re
    src = open(os.path.join(os.path.dirname(__file__), "..", "include", "hunter_model_constants.h")).read()
    m = re.search(name + r"\[[^\]]*\]\s*=\s*\{([^}]*)\}", src)
    return [float(v) for v in m.group(1).replace("\n", " ").split(",") if v.strip()]


def _header_value(name):
    import os, re
    src = open(os.path.join(os.path.dirname(__file__), "..", "include", "hunter_model_constants.h")).read()
    m = re.search(r"#define\s+" + name + r"\s+\(?([-+0-9.eE]+)", src)
    return float(m.group(1))


def constants_from_header():
    """The generated header carries the values parsed from reference.info / task.info; use them so the restatement follows the files."""
    global DEFAULT_JOINTS, COM_HEIGHT, NEXT_Z, SWING_HEIGHT, SWING_TIME_SCALE, FEET_BIAS
    DEFAULT_JOINTS = np.array(_header_array("HB_DEFAULT_JOINT_STATE"))
    COM_HEIGHT = _header_value("HB_COM_HEIGHT"); NEXT_Z = _header_value("HB_NEXT_POSITION_Z")
    SWING_HEIGHT = _header_value("HB_SWING_HEIGHT"); SWING_TIME_SCALE = _header_value("HB_SWING_TIME_SCALE")
    x1, x2, y, z = (_header_value("HB_FEET_BIAS_X1"), _header_value("HB_FEET_BIAS_X2"), _header_value("HB_FEET_BIAS_Y"), _header_value("HB_FEET_BIAS_Z"))
    FEET_BIAS = [(x1, y, z), (x1, -y, z), (x2, y, z), (x2, -y, z)]


constants_from_header()


def rot_zyx(e):
    z, y, x = e
    cz, sz, cy, sy, cx, sx = math.cos(z), math.sin(z), math.cos(y), math.sin(y), math.cos(x), math.sin(x)
    return np.array([[cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx],
                     [sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx],
                     [-sy, cy * sx, cy * cx]])


def stance_legs(mode):
    """modeNumber2StanceLeg (MotionPhaseDefinition.h:55-87): contacts l_toe, r_toe, l_heel, r_heel."""
    left, right = mode in (2, 3), mode in (1, 3)
    return [left, right, left, right]


class ModeSchedule:
    def __init__(self, events, modes):
        self.events, self.modes = list(events), list(modes)
        assert len(self.modes) == len(self.events) + 1

    def mode_at(self, t):
        return self.modes[bisect.bisect_left(self.events, t)]


def gait_schedule(gait, prev_event, start, final_time):
    """Initial {STANCE, STANCE} schedule, template inserted at `start`, tiled until final_time, closed by STANCE."""
    tm, tt = GAITS[gait]
    events, modes = [prev_event, start], [3, 3]
    while events[-1] < final_time:
        for i, m in enumerate(tm):
            modes.append(m)
            events.append(events[-1] + (tt[i + 1] - tt[i]))
    modes.append(3)
    return ModeSchedule(events, modes)


class Target:
    """TargetTrajectories with two samples; getDesiredState = clamped linear interpolation."""

    def __init__(self, times, states):
        self.times, self.states = list(times), [np.array(s, dtype=float) for s in states]

    def state(self, t):
        if t <= self.times[0]:
            return self.states[0].copy()
        if t >= self.times[-1]:
            return self.states[-1].copy()
        a = (t - self.times[0]) / (self.times[1] - self.times[0])
        return (1 - a) * self.states[0] + a * self.states[1]


def cmd_vel_to_target(cmd, time, state, time_to_target):
    pose = np.array(state[6:12], dtype=float)
    v = rot_zyx(pose[3:6]) @ np.array(cmd[:3], dtype=float)
    if abs(v[0]) < 0.06:
        v[0] = 0.0
    elif abs(v[1]) < 0.06:
        v[1] = 0.0
    target = np.array([pose[0] + v[0] * time_to_target, pose[1] + v[1] * time_to_target, COM_HEIGHT, pose[3] + cmd[3] * time_to_target, 0.0, 0.0])
    cur = pose.copy(); cur[4] = 0.0; cur[5] = 0.0
    cur[2] = pose[2] + max(-0.04, min(0.04, COM_HEIGHT - pose[2]))
    states = []
    for p in (cur, target):
        s = np.zeros(22); s[6:12] = p; s[12:22] = DEFAULT_JOINTS; s[0:3] = v
        states.append(s)
    return Target([time, time + time_to_target], states)


class CubicSpline:
    def __init__(self, n0, n1):
        (self.t0, p0, v0), (self.t1, p1, v1) = n0, n1
        self.dt = self.t1 - self.t0
        dp, dv = p1 - p0, v1 - v0
        self.c0 = p0; self.c1 = v0 * self.dt
        self.c2 = -(3.0 * v0 + dv) * self.dt + 3.0 * dp
        self.c3 = (2.0 * v0 + dv) * self.dt - 2.0 * dp

    def position(self, t):
        tn = (t - self.t0) / self.dt
        return self.c3 * tn ** 3 + self.c2 * tn ** 2 + self.c1 * tn + self.c0

    def velocity(self, t):
        tn = (t - self.t0) / self.dt
        return (3.0 * self.c3 * tn ** 2 + 2.0 * self.c2 * tn + self.c1) / self.dt


class MultiCubicSpline:
    def __init__(self, nodes):
        self.nodes = nodes
        self.splines = [CubicSpline(nodes[i], nodes[i + 1]) for i in range(len(nodes) - 1)]

    def _pick(self, t):
        for i in range(len(self.nodes) - 1):
            if self.nodes[i][0] <= t < self.nodes[i + 1][0]:
                return self.splines[i]
        if t < self.nodes[0][0]:
            return self.splines[0]
        return self.splines[-1]

    def position(self, t):
        return self._pick(t).position(t)

    def velocity(self, t):
        return self._pick(t).velocity(t)


def find_index(index, flags):
    n = len(flags)
    start = 0
    for ip in range(index - 1, -1, -1):
        if flags[ip] != flags[index]:
            start = ip
            break
    final = n - 2
    for ip in range(index + 1, n):
        if flags[ip] != flags[index]:
            final = ip - 1
            break
    return start, final


class SwingPlanner:
    """SwingTrajectoryPlanner: keeps latestStanceposition_ between updates."""

    def __init__(self, latest_stance=None):
        self.latest = np.zeros((4, 3)) if latest_stance is None else np.array(latest_stance, dtype=float).reshape(4, 3).copy()
        self.body_vel_cmd = np.zeros(6)
        self.current_feet = np.zeros((4, 3))

    def next_foot_pos(self, foot, current_time, stop_time, next_middle_time, next_middle_body_pos, current_body_pos, current_body_vel):
        roted_bias = rot_zyx(next_middle_body_pos[3:6]) @ np.array(FEET_BIAS[foot])
        rot = rot_zyx(current_body_pos[3:6])
        vel_cmd_linear = rot @ self.body_vel_cmd[:3]
        vel_cmd_angular = rot @ self.body_vel_cmd[3:]
        vel_linear = np.array(current_body_vel, dtype=float); vel_linear[2] = 0.0
        k = 0.03
        p_shoulder = (stop_time - current_time) * (0.5 * vel_linear + 0.5 * vel_cmd_linear) + roted_bias
        p_symmetry = (next_middle_time - stop_time) * vel_linear + k * (vel_linear - vel_cmd_linear)
        p_centrifugal = 0.5 * math.sqrt(current_body_pos[2] / 9.81) * np.cross(vel_linear, vel_cmd_angular)
        r = current_body_pos[:3] + p_shoulder + p_symmetry + p_centrifugal
        r[2] = NEXT_Z
        return r

    @staticmethod
    def swing_splines(t0, t1, a, b):
        out = []
        for ax in range(2):
            out.append(MultiCubicSpline([(t0, a[ax], 0.0),
                                         ((1 - 0.417) * t0 + 0.417 * t1, (1 - 0.650) * a[ax] + 0.650 * b[ax], 1.770 * (b[ax] - a[ax]) / (t1 - t0)),
                                         (t1, b[ax], 0.0)]))
        scaling = min(1.0, (t1 - t0) / SWING_TIME_SCALE)
        max_z = max(a[2], b[2]) + scaling * SWING_HEIGHT
        out.append(MultiCubicSpline([(t0, a[2], 0.0),
                                     ((1 - 0.251) * t0 + 0.251 * t1, 0.749 * max_z, 1.338 * (0.749 * (max_z - a[2])) / (0.251 * (t1 - t0))),
                                     ((1 - 0.630) * t0 + 0.630 * t1, 0.570 * max_z + (1 - 0.570) * b[2], 1.633 * 0.570 * (b[2] - max_z) / ((1 - 0.630) * (t1 - t0))),
                                     (t1, b[2], 0.0)]))
        return out

    def update(self, ms, target, init_time):
        self.events = ms.events
        legs = stance_legs(ms.mode_at(init_time + 0.001))
        for i in range(4):
            if legs[i]:
                self.latest[i] = self.current_feet[i]
            self.latest[i][2] = NEXT_Z
        last, nxt = self.latest.copy(), self.latest.copy()
        n = len(ms.modes)
        self.trajs = [[None] * n for _ in range(4)]
        for j in range(4):
            flags = [stance_legs(m)[j] for m in ms.modes]
            last_final = 0
            for p in range(n):
                si, fi = find_index(p, flags)
                if not flags[p]:
                    if si < 0 or fi >= n - 1:
                        raise RuntimeError("swing phase without take-off / touch-down")
                    ts, tf = ms.events[si], ms.events[fi]
                    if init_time < tf and fi > last_final:
                        last[j] = nxt[j]
                        if fi < n - 1:
                            _, fi2 = find_index(fi + 1, flags)
                            mid = 0.5 * (tf + ms.events[fi2])
                        else:
                            mid = tf
                        nxt[j] = self.next_foot_pos(j, init_time, tf, mid, target.state(mid)[6:12], target.state(init_time)[6:12], target.states[0][0:3])
                        last_final = fi
                    self.trajs[j][p] = self.swing_splines(ts, tf, last[j].copy(), nxt[j].copy())
                else:
                    ts, tf = ms.events[si], ms.events[fi]
                    if tf > ts:
                        self.trajs[j][p] = [MultiCubicSpline([(ts, nxt[j][a], 0.0), (tf, nxt[j][a], 0.0)]) for a in range(3)]
                    else:   # zero-length spline of the reference (never queried inside a solver window)
                        self.trajs[j][p] = None

    def foot(self, j, t):
        """(position(3), velocity(3)) of contact j at time t: phase index by lower_bound on the events (lookup::findIndexInTimeArray)."""
        p = bisect.bisect_left(self.events, t)
        sp = self.trajs[j][p]
        return np.array([s.position(t) for s in sp]), np.array([s.velocity(t) for s in sp])


def plan(t0, horizon, x0, cmd_vel, feet_pos, gait, gait_start, prev_event=None, time_to_target=None, latest_stance=None):
    """One instance: returns (ModeSchedule, Target, SwingPlanner) after update, mirroring SwitchedModelReferenceManager::modifyReferences."""
    prev_event = min(t0, gait_start) - 0.5 if prev_event is None else prev_event
    ttt = horizon if time_to_target is None else time_to_target
    ms = gait_schedule(gait, prev_event, gait_start, t0 + 2 * horizon)
    tg = cmd_vel_to_target(cmd_vel, t0, x0, ttt)
    sp = SwingPlanner(latest_stance)
    sp.body_vel_cmd = np.array([cmd_vel[0], cmd_vel[1], cmd_vel[2], cmd_vel[3], 0.0, 0.0])
    sp.current_feet = np.array(feet_pos, dtype=float).reshape(4, 3)
    sp.update(ms, tg, t0)
    return ms, tg, sp


def sample(ms, tg, sp, times, post_event=True):
    """Node-sampled references: x_ref (n x 22), swing (n x 24: per contact pos(3), vel(3)), mode (n)."""
    n = len(times)
    x_ref = np.zeros((n, 22)); swing = np.zeros((n, 24)); mode = np.zeros(n, dtype=np.int32)
    for k, t in enumerate(times):
        x_ref[k] = tg.state(t)
        # the mode of the interval starting at t (a node on an event takes the post-event mode, as the SQP time discretisation does)
        mode[k] = ms.modes[bisect.bisect_right(ms.events, t + 1e-9)] if post_event else ms.mode_at(t)
        for j in range(4):
            p, v = sp.foot(j, t)
            swing[k, 6 * j:6 * j + 3] = p; swing[k, 6 * j + 3:6 * j + 6] = v
    return x_ref, swing, mode


def eval_compact(ref, times):
    """Evaluate a compact hb_reference (ctypes struct) at `times` on the CPU; restates the device expansion kernel for CPU tests."""
    n = len(times)
    x_ref = np.zeros((n, 22)); swing = np.zeros((n, 24)); mode = np.zeros(n, dtype=np.int32)
    ev = [ref.event_times[i] for i in range(ref.n_events)]
    tt = [ref.target_times[i] for i in range(ref.n_targets)]
    ts = [np.array(ref.target_states[i][:]) for i in range(ref.n_targets)]
    for k, t in enumerate(times):
        mode[k] = ref.modes[bisect.bisect_right(ev, t + 1e-9)]
        if len(tt) <= 1 or t <= tt[0]:
            x_ref[k] = ts[0]
        elif t >= tt[-1]:
            x_ref[k] = ts[-1]
        else:
            s = max(0, bisect.bisect_right(tt, t) - 1); s = min(s, len(tt) - 2)
            a = (t - tt[s]) / (tt[s + 1] - tt[s])
            x_ref[k] = (1 - a) * ts[s] + a * ts[s + 1]
        for c in range(4):
            for a in range(3):
                ns = ref.n_segments[c][a]
                if ns == 0:
                    continue
                s = 0
                while s + 1 < ns and t >= ref.segments[c][a][s][1]:
                    s += 1
                sg = ref.segments[c][a][s]
                cs = CubicSpline((sg[0], sg[2], sg[3]), (sg[1], sg[4], sg[5]))
                swing[k, 6 * c + a] = cs.position(t); swing[k, 6 * c + 3 + a] = cs.velocity(t)
    return x_ref, swing, mode
