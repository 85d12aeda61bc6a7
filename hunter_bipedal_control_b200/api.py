"""ctypes binding of libhunter_b200.so (C ABI in include/hunter_b200.h) and host-side mirrors of the reference operators.

The product path has no CPU fallback: if the CUDA library is missing or no B200 is visible, every entry point raises.

Mirrors of the reference interface for this path (same names / argument meaning / error behaviour):
  * ``WeightedWbc.update(stateDesired, inputDesired, rbdStateMeasured, mode, period)``  -- legged_wbc/include/legged_wbc/WbcBase.h:43-44,
    legged_wbc/src/WeightedWbc.cpp:18-66 (returns the 38-vector [qdd, F, tau]; on solver failure prints and re-uses the last solution).
  * ``SqpMpc.advance / evaluatePolicy``  -- the calls legged_controllers/src/LeggedController.cpp:144-156,406 makes through MPC_MRT_Interface.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libhunter_b200.so")
_lib = None

NX, NU, NQ, NJ, NWBC = 22, 22, 16, 10, 38
HB_MAX_EVENTS, HB_MAX_TARGETS, HB_MAX_SEGMENTS = 32, 16, 24

EXPORTED_SYMBOLS = [
    "hb_shard_partition", "hb_shard_sort_by_schedule", "hb_shard_unique_id", "hb_shard_create", "hb_shard_destroy", "hb_shard_block", "hb_shard_gather_dev", "hb_shard_wait", "hb_shard_last_error",
    "hb_default_config", "hb_create", "hb_destroy", "hb_sync", "hb_strerror", "hb_last_cuda_error", "hb_launch_count", "hb_last_reference_upload_bytes", "hb_stream", "hb_profile_enable", "hb_profile_read",
    "hb_wbc_qp_batch_dev", "hb_wbc_qp_rows_batch_dev", "hb_wbc_assemble_batch_dev", "hb_wbc_assemble_batch", "hb_wbc_solve_batch_dev", "hb_mpc_cold_start_batch_dev", "hb_mpc_solve_batch_dev",
    "hb_policy_eval_batch_dev", "hb_control_step_batch_dev", "hb_rbd_to_centroidal_batch_dev", "hb_reference_expand_batch_dev",
    "hb_probe_flow_map_dev", "hb_contact_positions_batch_dev", "hb_contact_positions_batch", "hb_plan_references", "hb_plan_set_threads", "hb_gait_select", "hb_resident_cycle_batch_dev", "hb_resident_cycle_batch", "hb_resident_read_batch", "hb_plan_references_batch_dev",
    "hb_plan_references_gpu", "hb_resident_plan_cycle_batch", "hb_default_kf_params", "hb_kf_reset", "hb_estimator_update_batch_dev",
    "hb_estimator_update_batch", "hb_default_pd_gains", "hb_joint_command_batch_dev", "hb_joint_command_batch",
    "hb_wbc_qp_batch", "hb_wbc_solve_batch", "hb_mpc_cold_start_batch", "hb_mpc_solve_batch", "hb_control_step_batch",
    "hb_rbd_to_centroidal_batch", "hb_reference_expand_batch", "hb_probe_flow_map",
    "hb_observer_reset", "hb_contact_force_estimate_batch_dev", "hb_contact_force_estimate_batch",
    "hb_default_wbc_settings", "hb_parse_task_info", "hb_wbc_get_settings", "hb_wbc_set_settings", "hb_wbc_set_kp_kd", "hb_load_task_info",
    "hb_hoqp_solve_batch_dev", "hb_hierarchical_wbc_solve_batch_dev", "hb_hoqp_solve_batch", "hb_hierarchical_wbc_solve_batch", "hb_hierarchical_wbc_tasks_batch",
    "hb_default_sim_params", "hb_actuation_reset", "hb_actuation_batch_dev", "hb_actuation_batch", "hb_sim_step_batch_dev", "hb_sim_step_batch",
    "hb_resident_wbc_batch_dev", "hb_resident_wbc_batch",
    "hb_time_grid_batch_dev", "hb_reference_expand_grid_batch_dev", "hb_mpc_solve_grid_batch_dev", "hb_policy_eval_grid_batch_dev",
    "hb_time_grid_batch", "hb_reference_expand_grid_batch", "hb_mpc_solve_grid_batch", "hb_resident_read_grid_batch", "hb_resident_write_batch",
]


class HbConfig(C.Structure):
    _fields_ = [("horizon_N", C.c_int32), ("dt", C.c_double), ("max_batch", C.c_int32), ("wbc_rho", C.c_double),
                ("qp_max_iter", C.c_int32), ("line_search_max_trials", C.c_int32), ("time_horizon", C.c_double), ("event_nodes", C.c_int32), ("e2e_chunks", C.c_int32)]


class HbSolveInfo(C.Structure):
    _fields_ = [("alpha", C.c_double), ("merit0", C.c_double), ("merit1", C.c_double), ("viol0", C.c_double), ("viol1", C.c_double),
                ("armijo", C.c_double), ("status", C.c_int32), ("n_trials", C.c_int32)]


class HbReference(C.Structure):
    _fields_ = [("n_events", C.c_int32), ("event_times", C.c_double * HB_MAX_EVENTS), ("modes", C.c_int32 * (HB_MAX_EVENTS + 1)),
                ("n_targets", C.c_int32), ("target_times", C.c_double * HB_MAX_TARGETS), ("target_states", (C.c_double * 22) * HB_MAX_TARGETS),
                ("n_segments", (C.c_int32 * 3) * 4), ("segments", (((C.c_double * 6) * HB_MAX_SEGMENTS) * 3) * 4)]


class HbPlanInput(C.Structure):
    _fields_ = [("t0", C.c_double), ("horizon", C.c_double), ("time_to_target", C.c_double), ("gait_start", C.c_double), ("prev_event", C.c_double),
                ("x0", C.c_double * 22), ("cmd_vel", C.c_double * 4), ("feet_pos", C.c_double * 12), ("gait", C.c_int32), ("joint_ik", C.c_int32)]


class HbPdGains(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("kp_position", "kd_position", "kp_big_stance", "kp_big_swing", "kd_big", "kp_small_stance",
                                          "kp_small_swing", "kd_small", "kd_feet")]


def default_pd_gains():
    g = HbPdGains()
    _check(load_library().hb_default_pd_gains(C.byref(g)), "hb_default_pd_gains")
    return g


class HbKfState(C.Structure):
    _fields_ = [("x_hat", C.c_double * 18), ("P", C.c_double * 324), ("feet_heights", C.c_double * 4)]


class HbKfParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("foot_radius", "imu_process_noise_position", "imu_process_noise_velocity", "foot_process_noise_position",
                                          "foot_sensor_noise_position", "foot_sensor_noise_velocity", "foot_height_sensor_noise")]


class HbWbcSettings(C.Structure):
    _fields_ = [("torque_limits", C.c_double * 5)] + [(k, C.c_double) for k in (
        "friction_coefficient", "swing_kp", "swing_kd", "base_accel_kp", "base_accel_kd", "base_height_kp", "base_height_kd", "base_angular_kp",
        "base_angular_kd", "weight_swing_leg", "weight_base_accel", "weight_contact_force")]

    def as_array(self):
        return np.frombuffer(bytes(self), dtype=np.float64).copy()


class HbTaskInfo(C.Structure):
    _fields_ = [("wbc", HbWbcSettings), ("kalman", C.c_double * 7), ("contact_force_cutoff_frequency", C.c_double), ("contact_threshold", C.c_double),
                ("sqp_dt", C.c_double), ("sqp_iteration", C.c_int32), ("mpc_time_horizon", C.c_double), ("mpc_cold_start", C.c_int32), ("found", C.c_int32)]


def parse_task_info(path):
    """hb_parse_task_info: the WBC / estimator / discretisation settings of a task.info file (host only)."""
    ti = HbTaskInfo()
    _check(load_library().hb_parse_task_info(str(path).encode(), C.byref(ti)), "hb_parse_task_info")
    return ti


HB_ACT_CAPACITY = 16
HB_HOQP_MAX_LEVELS, HB_HOQP_N, HB_HOQP_MAX_EQ, HB_HOQP_MAX_IN, HB_HOQP_MAX_STACKED = 3, 38, 32, 40, 80


class HbHoqpProblem(C.Structure):
    _fields_ = [("n", C.c_int32), ("levels", C.c_int32), ("ma", C.c_int32 * HB_HOQP_MAX_LEVELS), ("md", C.c_int32 * HB_HOQP_MAX_LEVELS),
                ("a", ((C.c_double * HB_HOQP_N) * HB_HOQP_MAX_EQ) * HB_HOQP_MAX_LEVELS), ("b", (C.c_double * HB_HOQP_MAX_EQ) * HB_HOQP_MAX_LEVELS),
                ("d", ((C.c_double * HB_HOQP_N) * HB_HOQP_MAX_IN) * HB_HOQP_MAX_LEVELS), ("f", (C.c_double * HB_HOQP_MAX_IN) * HB_HOQP_MAX_LEVELS)]


def make_hoqp_problems(hierarchies):
    """hierarchies: list (one per instance) of lists of tasks (a, b, d, f) by decreasing priority -> ctypes array of HbHoqpProblem."""
    pbs = (HbHoqpProblem * len(hierarchies))()
    for pb, levels in zip(pbs, hierarchies):
        pb.levels = len(levels)
        n = 0
        for l, (a, b, d, f) in enumerate(levels):
            a = np.zeros((0, 0)) if a is None else np.atleast_2d(np.asarray(a, dtype=float)); d = np.zeros((0, 0)) if d is None else np.atleast_2d(np.asarray(d, dtype=float))
            n = max(n, a.shape[1] if a.size else 0, d.shape[1] if d.size else 0)
            pb.ma[l] = a.shape[0] if a.size else 0; pb.md[l] = d.shape[0] if d.size else 0
            for i in range(pb.ma[l]):
                pb.b[l][i] = float(b[i])
                for j in range(a.shape[1]):
                    pb.a[l][i][j] = a[i, j]
            for i in range(pb.md[l]):
                pb.f[l][i] = float(f[i])
                for j in range(d.shape[1]):
                    pb.d[l][i][j] = d[i, j]
        pb.n = n
    return pbs


def hoqp_tasks(pb):
    """HbHoqpProblem -> list of (a, b, d, f) numpy tasks."""
    out = []
    for l in range(pb.levels):
        a = np.array([[pb.a[l][i][j] for j in range(pb.n)] for i in range(pb.ma[l])]).reshape(pb.ma[l], pb.n)
        d = np.array([[pb.d[l][i][j] for j in range(pb.n)] for i in range(pb.md[l])]).reshape(pb.md[l], pb.n)
        out.append((a, np.array([pb.b[l][i] for i in range(pb.ma[l])]), d, np.array([pb.f[l][i] for i in range(pb.md[l])])))
    return out


class HbActuationState(C.Structure):
    _fields_ = [("count", C.c_int32), ("head", C.c_int32), ("stamp", C.c_double * HB_ACT_CAPACITY), ("cmd", (C.c_double * 50) * HB_ACT_CAPACITY)]


class HbSimParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("substeps", C.c_int32), ("ground_height", C.c_double), ("ground_stiffness", C.c_double), ("ground_damping", C.c_double),
                ("tangential_damping", C.c_double), ("friction_mu", C.c_double), ("joint_armature", C.c_double), ("joint_damping", C.c_double)]


def default_sim_params():
    p = HbSimParams()
    _check(load_library().hb_default_sim_params(C.byref(p)), "hb_default_sim_params")
    return p


def actuation_states(B):
    st = (HbActuationState * B)()
    _check(load_library().hb_actuation_reset(B, st), "hb_actuation_reset")
    return st


class HbObserverState(C.Structure):
    _fields_ = [("p_filtered", C.c_double * 16)]


def observer_states(B):
    """Freshly reset momentum-observer states (pSCgZinvlast_ = 0)."""
    st = (HbObserverState * B)()
    _check(load_library().hb_observer_reset(B, st), "hb_observer_reset")
    return st


def default_kf_params():
    p = HbKfParams()
    _check(load_library().hb_default_kf_params(C.byref(p)), "hb_default_kf_params")
    return p


def kf_states(B):
    """Freshly reset filter states (x_hat = 0, P = 100 I)."""
    st = (HbKfState * B)()
    _check(load_library().hb_kf_reset(B, st), "hb_kf_reset")
    return st


class HbGaitSelector(C.Structure):
    _fields_ = [("history", C.c_double * 50), ("vel_avg", C.c_double), ("head", C.c_int32), ("count", C.c_int32), ("gait_level", C.c_int32),
                ("reserved", C.c_int32)]


class GaitSelector:
    """Batch of speed-based gait selectors (SwitchedModelReferenceManager::calculateVelAbs + walkGait / trotGait)."""

    def __init__(self, B, gait_level=-1):
        self.B = B
        self.state = (HbGaitSelector * B)()
        for i in range(B):
            self.state[i].gait_level = gait_level

    def update(self, cmd_vel, target_state0, gait_type=0):
        lib = load_library()
        cmd_vel = _f64(np.broadcast_to(_f64(cmd_vel), (self.B, 4))); ts = _f64(target_state0).reshape(self.B, 22)
        gt = np.ascontiguousarray(np.broadcast_to(np.asarray(gait_type, dtype=np.int32), (self.B,)))
        level = np.zeros(self.B, dtype=np.int32); insert = np.zeros(self.B, dtype=np.int32)
        _check(lib.hb_gait_select(self.B, self.state, _ptr(gt), _ptr(cmd_vel), _ptr(ts), _ptr(level), _ptr(insert)), "hb_gait_select")
        return level, insert

    @property
    def vel_avg(self):
        return np.array([self.state[i].vel_avg for i in range(self.B)])


GAIT_IDS = {"stance": 0, "trot": 1, "standing_trot": 2, "flying_trot": 3}


def _gait_ids(gait, B):
    """Gait name (one for the batch) or any sequence / ndarray of names or ids (one per instance) -> int ids [B]."""
    if isinstance(gait, str):
        return [GAIT_IDS[gait]] * B
    if np.ndim(gait) == 0:
        return [int(gait)] * B
    ids = [GAIT_IDS[g] if isinstance(g, str) else int(g) for g in list(gait)]
    if len(ids) != B:
        raise ValueError("gait: expected one name or %d per-instance entries, got %d" % (B, len(ids)))
    return ids


def make_plan_inputs(t0, horizon, x0, cmd_vel, feet_pos, gait, gait_start, prev_event=None, time_to_target=None, joint_ik=True):
    """ctypes array of HbPlanInput for a batch (feet_pos may be None when the device computes it)."""
    x0 = _f64(x0); B = x0.shape[0]
    gids = _gait_ids(gait, B)
    cmd_vel = np.broadcast_to(_f64(cmd_vel), (B, 4))
    feet_pos = np.zeros((B, 12)) if feet_pos is None else _f64(feet_pos).reshape(B, 12)
    t0 = np.broadcast_to(_f64(t0), (B,)); gait_start = np.broadcast_to(_f64(gait_start), (B,))
    ins = (HbPlanInput * B)()
    for i in range(B):
        p = ins[i]
        p.t0 = t0[i]; p.horizon = horizon; p.time_to_target = horizon if time_to_target is None else time_to_target
        p.gait_start = gait_start[i]; p.prev_event = (min(t0[i], gait_start[i]) - 0.5) if prev_event is None else prev_event
        p.gait = gids[i]
        p.joint_ik = 1 if joint_ik else 0
        for j in range(22): p.x0[j] = x0[i, j]
        for j in range(4): p.cmd_vel[j] = cmd_vel[i, j]
        for j in range(12): p.feet_pos[j] = feet_pos[i, j]
    return ins


def plan_set_threads(n):
    """Host threads of plan_references (0 = all); hb_plan_set_threads."""
    _check(load_library().hb_plan_set_threads(int(n)), "hb_plan_set_threads")


def plan_references(t0, horizon, x0, cmd_vel, feet_pos, gait, gait_start, prev_event=None, time_to_target=None, latest_stance=None, joint_ik=True):
    """Host-side reference planner (hb_plan_references): returns (ctypes array of HbReference, latest_stance[B,12])."""
    lib = load_library()
    ins = make_plan_inputs(t0, horizon, x0, cmd_vel, feet_pos, gait, gait_start, prev_event, time_to_target, joint_ik)
    B = len(ins)
    ls = np.zeros((B, 12)) if latest_stance is None else _f64(latest_stance).copy()
    refs = (HbReference * B)()
    _check(lib.hb_plan_references(B, ins, _ptr(ls), refs), "hb_plan_references")
    return refs, ls


INFO_DTYPE = np.dtype([("alpha", "f8"), ("merit0", "f8"), ("merit1", "f8"), ("viol0", "f8"), ("viol1", "f8"), ("armijo", "f8"),
                       ("status", "i4"), ("n_trials", "i4")], align=True)
assert INFO_DTYPE.itemsize == C.sizeof(HbSolveInfo)


class HunterB200Error(RuntimeError):
    pass


def load_library():
    """Load libhunter_b200.so. Raises if the CUDA extension has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise HunterB200Error("libhunter_b200.so is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                                  "(the product path has no CPU fallback)")
        _lib = C.CDLL(_LIB_PATH)
        _lib.hb_strerror.restype = C.c_char_p
        _lib.hb_last_cuda_error.restype = C.c_char_p
        _lib.hb_launch_count.restype = C.c_int64
        _lib.hb_last_reference_upload_bytes.restype = C.c_int64
        _lib.hb_stream.restype = C.c_void_p
        _lib.hb_shard_last_error.restype = C.c_char_p
    return _lib


def _check(rc, what, ctx=None):
    if rc != 0:
        extra = ""
        if rc == -2 and ctx is not None:
            extra = ": " + load_library().hb_last_cuda_error(ctx).decode()
        raise HunterB200Error("%s failed: %s (%d)%s" % (what, load_library().hb_strerror(rc).decode(), rc, extra))


def _ptr(a):
    """Pointer of a numpy array (host) or of a torch tensor (host or cuda)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    return C.c_void_p(a.data_ptr())


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """Owner of one hb_ctx (one GPU, one stream). Single-owner: use it from one thread at a time."""

    def __init__(self, horizon_N=100, dt=0.01, max_batch=1024, device=0, wbc_rho=1e-8, qp_max_iter=40, line_search_max_trials=14, time_horizon=0.0,
                 event_nodes=False, e2e_chunks=0):
        lib = load_library()
        cfg = HbConfig()
        _check(lib.hb_default_config(C.byref(cfg)), "hb_default_config")
        cfg.horizon_N, cfg.dt, cfg.max_batch, cfg.wbc_rho = horizon_N, dt, max_batch, wbc_rho
        cfg.qp_max_iter, cfg.line_search_max_trials = qp_max_iter, line_search_max_trials
        cfg.time_horizon, cfg.event_nodes, cfg.e2e_chunks = float(time_horizon), 1 if event_nodes else 0, int(e2e_chunks)
        self.cfg = cfg
        self.N, self.dt, self.max_batch, self.device = horizon_N, dt, max_batch, device
        self._h = C.c_void_p()
        _check(lib.hb_create(C.byref(cfg), C.c_int(device), C.byref(self._h)), "hb_create")
        self._lib = lib

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.hb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _check(self._lib.hb_sync(self._h), "hb_sync")

    # ------------------------------------------------------------------ run-time WBC settings (WbcBase::loadTasksSetting / setKpKd)
    def wbc_settings(self):
        s = HbWbcSettings()
        _check(self._lib.hb_wbc_get_settings(self._h, C.byref(s)), "hb_wbc_get_settings")
        return s

    def set_wbc_settings(self, s):
        _check(self._lib.hb_wbc_set_settings(self._h, C.byref(s)), "hb_wbc_set_settings")

    def set_kp_kd(self, swing_kp, swing_kd):
        _check(self._lib.hb_wbc_set_kp_kd(self._h, C.c_double(swing_kp), C.c_double(swing_kd)), "hb_wbc_set_kp_kd")

    def load_task_info(self, path):
        _check(self._lib.hb_load_task_info(self._h, str(path).encode()), "hb_load_task_info")

    @property
    def launch_count(self):
        return int(self._lib.hb_launch_count(self._h))

    def profile_enable(self, on=True):
        _check(self._lib.hb_profile_enable(self._h, int(on)), "hb_profile_enable", self._h)

    def profile_read(self):
        ms = np.zeros(7); cnt = np.zeros(7, dtype=np.int64)
        _check(self._lib.hb_profile_read(self._h, _ptr(ms), _ptr(cnt)), "hb_profile_read", self._h)
        names = ["mpc_riccati", "mpc_forward_linesearch", "wbc_assemble", "qp_ipm", "other", "mpc_linearise", "mpc_lq_project"]
        return {n: dict(ms=float(m), launches=int(c)) for n, m, c in zip(names, ms, cnt)}

    @property
    def last_reference_upload_bytes(self):
        return int(self._lib.hb_last_reference_upload_bytes(self._h))

    @property
    def stream_handle(self):
        return int(self._lib.hb_stream(self._h) or 0)

    # ------------------------------------------------------------------ host-pointer calls (numpy in / numpy out)
    def wbc_qp(self, H, g, A, lbA, ubA):
        H, g, A, lbA, ubA = map(_f64, (H, g, A, lbA, ubA))
        B, n = g.shape
        m = lbA.shape[1]
        x = np.zeros((B, n)); st = np.zeros(B, dtype=np.int32); it = np.zeros(B, dtype=np.int32)
        _check(self._lib.hb_wbc_qp_batch(self._h, B, n, m, _ptr(H), _ptr(g), _ptr(A), _ptr(lbA), _ptr(ubA), _ptr(x), _ptr(st), _ptr(it)), "hb_wbc_qp_batch", self._h)
        return x, st, it

    def wbc_solve(self, x_des, u_des, rbd, mode, stance_mode=None):
        x_des, u_des, rbd = map(_f64, (x_des, u_des, rbd))
        B = x_des.shape[0]
        mode = np.ascontiguousarray(mode, dtype=np.int32)
        sm = None if stance_mode is None else np.ascontiguousarray(stance_mode, dtype=np.uint8)
        sol = np.zeros((B, NWBC)); st = np.zeros(B, dtype=np.int32)
        _check(self._lib.hb_wbc_solve_batch(self._h, B, _ptr(x_des), _ptr(u_des), _ptr(rbd), _ptr(mode), _ptr(sm), _ptr(sol), _ptr(st)), "hb_wbc_solve_batch", self._h)
        return sol, st

    def wbc_assemble(self, x_des, u_des, rbd, mode, stance_mode=None):
        """WeightedWbc's QP in the layout handed to qpOASES: returns (H [B,38,38], g [B,38], A [B,60,38], lbA, ubA [B,60], m_rows [B])."""
        x_des, u_des, rbd = map(_f64, (x_des, u_des, rbd))
        B = x_des.shape[0]
        mode = np.ascontiguousarray(mode, dtype=np.int32)
        sm = None if stance_mode is None else np.ascontiguousarray(stance_mode, dtype=np.uint8)
        H = np.zeros((B, NWBC, NWBC)); g = np.zeros((B, NWBC)); A = np.zeros((B, 60, NWBC)); lb = np.zeros((B, 60)); ub = np.zeros((B, 60))
        m = np.zeros(B, dtype=np.int32)
        _check(self._lib.hb_wbc_assemble_batch(self._h, B, _ptr(x_des), _ptr(u_des), _ptr(rbd), _ptr(mode), _ptr(sm), _ptr(H), _ptr(g), _ptr(A), _ptr(lb),
                                               _ptr(ub), _ptr(m)), "hb_wbc_assemble_batch", self._h)
        return H, g, A, lb, ub, m

    def hoqp_solve(self, problems):
        """legged::HoQp cascade for a batch (ctypes array of HbHoqpProblem): returns (x [B,38], stacked slack [B,80], status [B])."""
        B = len(problems)
        x = np.zeros((B, HB_HOQP_N)); sl = np.zeros((B, HB_HOQP_MAX_STACKED)); st = np.zeros(B, dtype=np.int32)
        _check(self._lib.hb_hoqp_solve_batch(self._h, B, problems, _ptr(x), _ptr(sl), _ptr(st)), "hb_hoqp_solve_batch", self._h)
        return x, sl, st

    def hierarchical_wbc_tasks(self, x_des, u_des, rbd, mode):
        x_des, u_des, rbd = map(_f64, (x_des, u_des, rbd)); B = x_des.shape[0]
        mode = np.ascontiguousarray(mode, dtype=np.int32)
        pbs = (HbHoqpProblem * B)()
        _check(self._lib.hb_hierarchical_wbc_tasks_batch(self._h, B, _ptr(x_des), _ptr(u_des), _ptr(rbd), _ptr(mode), pbs), "hb_hierarchical_wbc_tasks_batch", self._h)
        return pbs

    def hierarchical_wbc_solve(self, x_des, u_des, rbd, mode):
        """legged::HierarchicalWbc::update for a batch: returns (sol [B,38] = [qdd, F, tau], status [B])."""
        x_des, u_des, rbd = map(_f64, (x_des, u_des, rbd)); B = x_des.shape[0]
        mode = np.ascontiguousarray(mode, dtype=np.int32)
        sol = np.zeros((B, NWBC)); st = np.zeros(B, dtype=np.int32)
        _check(self._lib.hb_hierarchical_wbc_solve_batch(self._h, B, _ptr(x_des), _ptr(u_des), _ptr(rbd), _ptr(mode), _ptr(sol), _ptr(st)),
               "hb_hierarchical_wbc_solve_batch", self._h)
        return sol, st

    def mpc_cold_start(self, x0, mode):
        x0 = _f64(x0); B = x0.shape[0]
        mode = np.ascontiguousarray(mode, dtype=np.int32)
        xt = np.zeros((B, self.N + 1, NX)); ut = np.zeros((B, self.N, NU))
        _check(self._lib.hb_mpc_cold_start_batch(self._h, B, _ptr(x0), _ptr(mode), _ptr(xt), _ptr(ut)), "hb_mpc_cold_start_batch", self._h)
        return xt, ut

    def mpc_solve(self, x0, x_ref, swing, mode, xt, ut):
        x0, x_ref, swing = map(_f64, (x0, x_ref, swing))
        B = x0.shape[0]
        mode = np.ascontiguousarray(mode, dtype=np.int32)
        xt = _f64(xt).copy(); ut = _f64(ut).copy()
        info = np.zeros(B, dtype=INFO_DTYPE)
        _check(self._lib.hb_mpc_solve_batch(self._h, B, _ptr(x0), _ptr(x_ref), _ptr(swing), _ptr(mode), _ptr(xt), _ptr(ut), _ptr(info)), "hb_mpc_solve_batch", self._h)
        return xt, ut, info

    def time_grid(self, t0, refs):
        """Event-node time discretisation (row S1): returns (node_times [B, N+1], n_intervals [B], status [B])."""
        t0 = _f64(t0); B = t0.shape[0]
        tk = np.zeros((B, self.N + 1)); nn = np.zeros(B, dtype=np.int32); st = np.zeros(B, dtype=np.int32)
        _check(self._lib.hb_time_grid_batch(self._h, B, _ptr(t0), C.cast(refs, C.c_void_p), _ptr(tk), _ptr(nn), _ptr(st)), "hb_time_grid_batch", self._h)
        return tk, nn, st

    def reference_expand_grid(self, node_times, refs):
        tk = _f64(node_times); B = tk.shape[0]
        x_ref = np.zeros((B, self.N + 1, NX)); swing = np.zeros((B, self.N + 1, 24)); mode = np.zeros((B, self.N + 1), dtype=np.int32)
        _check(self._lib.hb_reference_expand_grid_batch(self._h, B, _ptr(tk), C.cast(refs, C.c_void_p), _ptr(x_ref), _ptr(swing), _ptr(mode)),
               "hb_reference_expand_grid_batch", self._h)
        return x_ref, swing, mode

    def mpc_solve_grid(self, x0, node_times, n_intervals, x_ref, swing, mode, xt, ut):
        x0, tk, x_ref, swing = map(_f64, (x0, node_times, x_ref, swing))
        B = x0.shape[0]
        nn = np.ascontiguousarray(n_intervals, dtype=np.int32); mode = np.ascontiguousarray(mode, dtype=np.int32)
        xt = _f64(xt).copy(); ut = _f64(ut).copy()
        info = np.zeros(B, dtype=INFO_DTYPE)
        _check(self._lib.hb_mpc_solve_grid_batch(self._h, B, _ptr(x0), _ptr(tk), _ptr(nn), _ptr(x_ref), _ptr(swing), _ptr(mode), _ptr(xt), _ptr(ut), _ptr(info)),
               "hb_mpc_solve_grid_batch", self._h)
        return xt, ut, info

    def resident_write(self, t0, xt, ut, mode=None, node_times=None, n_intervals=None):
        """Restore a resident-solution snapshot (hb_resident_write_batch)."""
        t0, xt, ut = _f64(t0), _f64(xt), _f64(ut); B = t0.shape[0]
        md = None if mode is None else np.ascontiguousarray(mode, dtype=np.int32)
        tk = None if node_times is None else _f64(node_times); nn = None if n_intervals is None else np.ascontiguousarray(n_intervals, dtype=np.int32)
        _check(self._lib.hb_resident_write_batch(self._h, B, _ptr(t0), _ptr(xt), _ptr(ut), _ptr(md), _ptr(tk), _ptr(nn)), "hb_resident_write_batch", self._h)

    def resident_read_grid(self, B):
        tk = np.zeros((B, self.N + 1)); nn = np.zeros(B, dtype=np.int32)
        _check(self._lib.hb_resident_read_grid_batch(self._h, B, _ptr(tk), _ptr(nn)), "hb_resident_read_grid_batch", self._h)
        return tk, nn

    def control_step(self, t_rel, x0, x_ref, swing, mode, rbd, xt, ut):
        x0, x_ref, swing, rbd = map(_f64, (x0, x_ref, swing, rbd))
        B = x0.shape[0]
        mode = np.ascontiguousarray(mode, dtype=np.int32)
        xt = _f64(xt).copy(); ut = _f64(ut).copy()
        info = np.zeros(B, dtype=INFO_DTYPE)
        sol = np.zeros((B, NWBC)); tau = np.zeros((B, NJ)); st = np.zeros(B, dtype=np.int32)
        _check(self._lib.hb_control_step_batch(self._h, B, C.c_double(t_rel), _ptr(x0), _ptr(x_ref), _ptr(swing), _ptr(mode), _ptr(rbd), _ptr(xt), _ptr(ut),
                                               _ptr(info), _ptr(sol), _ptr(tau), _ptr(st)), "hb_control_step_batch", self._h)
        return xt, ut, info, sol, tau, st

    def resident_cycle(self, cold_start, t_rel, t0, x0, refs, rbd):
        """Resident closed-loop cycle (hb_resident_cycle_batch); refs: ctypes array of HbReference. Returns (info, sol, torque, status)."""
        t0, x0, rbd = _f64(t0), _f64(x0), _f64(rbd); B = x0.shape[0]
        info = np.zeros(B, dtype=INFO_DTYPE); sol = np.zeros((B, NWBC)); tau = np.zeros((B, NJ)); st = np.zeros(B, dtype=np.int32)
        _check(self._lib.hb_resident_cycle_batch(self._h, B, 1 if cold_start else 0, C.c_double(t_rel), _ptr(t0), _ptr(x0), C.cast(refs, C.c_void_p),
                                                 _ptr(rbd), _ptr(info), _ptr(sol), _ptr(tau), _ptr(st)), "hb_resident_cycle_batch", self._h)
        return info, sol, tau, st

    def resident_read(self, B):
        t0 = np.zeros(B); xt = np.zeros((B, self.N + 1, NX)); ut = np.zeros((B, self.N, NU))
        _check(self._lib.hb_resident_read_batch(self._h, B, _ptr(t0), _ptr(xt), _ptr(ut)), "hb_resident_read_batch", self._h)
        return t0, xt, ut

    def plan_references_gpu(self, ins, latest_stance=None):
        """Device planner with host pointers: returns (refs, latest_stance, status)."""
        B = len(ins)
        ls = np.zeros((B, 12)) if latest_stance is None else _f64(latest_stance).copy()
        refs = (HbReference * B)(); st = np.zeros(B, dtype=np.int32)
        _check(self._lib.hb_plan_references_gpu(self._h, B, ins, _ptr(ls), refs, _ptr(st)), "hb_plan_references_gpu", self._h)
        return refs, ls, st

    def resident_plan_cycle(self, cold_start, t_rel, ins, rbd):
        """Whole cycle from plan inputs (hb_resident_plan_cycle_batch). Returns (info, sol, torque, wbc_status, plan_status)."""
        rbd = _f64(rbd); B = len(ins)
        info = np.zeros(B, dtype=INFO_DTYPE); sol = np.zeros((B, NWBC)); tau = np.zeros((B, NJ)); st = np.zeros(B, dtype=np.int32); ps = np.zeros(B, dtype=np.int32)
        _check(self._lib.hb_resident_plan_cycle_batch(self._h, B, 1 if cold_start else 0, C.c_double(t_rel), ins, _ptr(rbd), _ptr(info), _ptr(sol), _ptr(tau),
                                                      _ptr(st), _ptr(ps)), "hb_resident_plan_cycle_batch", self._h)
        return info, sol, tau, st, ps

    def estimator_update(self, dt, state, quat, ang_vel_local, lin_acc_local, joint_pos, joint_vel, contact_flag, params=None):
        """KalmanFilterEstimate::update for a batch; `state` (ctypes array of HbKfState) is updated in place. Returns rbd [B,32]."""
        quat, ang_vel_local, lin_acc_local, joint_pos, joint_vel = map(_f64, (quat, ang_vel_local, lin_acc_local, joint_pos, joint_vel))
        B = quat.shape[0]
        flags = np.ascontiguousarray(contact_flag, dtype=np.uint8).reshape(B, 4)
        params = params or default_kf_params()
        rbd = np.zeros((B, 32))
        _check(self._lib.hb_estimator_update_batch(self._h, B, C.byref(params), C.c_double(dt), state, _ptr(quat), _ptr(ang_vel_local), _ptr(lin_acc_local),
                                                   _ptr(joint_pos), _ptr(joint_vel), _ptr(flags), _ptr(rbd)), "hb_estimator_update_batch", self._h)
        return rbd

    def actuation(self, time, state, command, rbd, delay=0.009):
        """LeggedHWSim::writeSim: delayed hybrid joint command -> applied joint torques [B,10]; `state` (ctypes array of HbActuationState) in place."""
        command, rbd = _f64(command), _f64(rbd); B = rbd.shape[0]
        time = _f64(np.broadcast_to(_f64(time), (B,)))
        tau = np.zeros((B, NJ))
        _check(self._lib.hb_actuation_batch(self._h, B, C.c_double(delay), _ptr(time), state, _ptr(command), _ptr(rbd), _ptr(tau)), "hb_actuation_batch", self._h)
        return tau

    def sim_step(self, rbd, tau, params=None):
        """One control period of the batched rigid-body plant: returns (rbd_next [B,32], contact_force [B,12], contact_flag [B,4])."""
        rbd = _f64(rbd).copy(); tau = _f64(tau); B = rbd.shape[0]
        params = params or default_sim_params()
        cf = np.zeros((B, 12)); fl = np.zeros((B, 4), dtype=np.uint8)
        _check(self._lib.hb_sim_step_batch(self._h, B, C.byref(params), _ptr(rbd), _ptr(tau), _ptr(cf), _ptr(fl)), "hb_sim_step_batch", self._h)
        return rbd, cf, fl

    def resident_wbc(self, t_now, rbd, stance_mode=None):
        """Policy of the resident solution at absolute time t_now + WeightedWbc: returns (x_des, u_des, mode, sol, torque, status)."""
        rbd = _f64(rbd); B = rbd.shape[0]
        t_now = _f64(np.broadcast_to(_f64(t_now), (B,)))
        sm = None if stance_mode is None else np.ascontiguousarray(stance_mode, dtype=np.uint8)
        xd = np.zeros((B, NX)); ud = np.zeros((B, NU)); md = np.zeros(B, dtype=np.int32); sol = np.zeros((B, NWBC)); tau = np.zeros((B, NJ)); st = np.zeros(B, dtype=np.int32)
        _check(self._lib.hb_resident_wbc_batch(self._h, B, _ptr(t_now), _ptr(rbd), _ptr(sm), _ptr(xd), _ptr(ud), _ptr(md), _ptr(sol), _ptr(tau), _ptr(st)),
               "hb_resident_wbc_batch", self._h)
        return xd, ud, md, sol, tau, st

    def contact_force_estimate(self, dt, state, rbd, tau_cmd, cutoff_frequency=250.0):
        """StateEstimateBase::estContactForce for a batch; `state` (ctypes array of HbObserverState) is updated in place.
        Returns (est_contact_force [B,16], disturbance_torque [B,16])."""
        rbd, tau_cmd = _f64(rbd), _f64(tau_cmd); B = rbd.shape[0]
        est = np.zeros((B, 16)); dist = np.zeros((B, 16))
        _check(self._lib.hb_contact_force_estimate_batch(self._h, B, C.c_double(cutoff_frequency), C.c_double(dt), state, _ptr(rbd), _ptr(tau_cmd), _ptr(est),
                                                         _ptr(dist)), "hb_contact_force_estimate_batch", self._h)
        return est, dist

    def joint_command(self, period, x_des, u_des, wbc_sol, mode_cmd, rbd, loaded=None, estop=None, gains=None):
        """Joint command law (LeggedController.cpp:186-257): returns (command [B,10,5], output_torque [B,10], estop [B])."""
        x_des, u_des, wbc_sol, rbd = _f64(x_des), _f64(u_des), _f64(wbc_sol), _f64(rbd); B = x_des.shape[0]
        mode_cmd = np.ascontiguousarray(mode_cmd, dtype=np.int32)
        gains = gains or default_pd_gains()
        ld = None if loaded is None else np.ascontiguousarray(loaded, dtype=np.uint8)
        es = np.zeros(B, dtype=np.uint8) if estop is None else np.ascontiguousarray(estop, dtype=np.uint8).copy()
        cmd = np.zeros((B, NJ, 5)); tau = np.zeros((B, NJ))
        _check(self._lib.hb_joint_command_batch(self._h, B, C.byref(gains), C.c_double(period), _ptr(x_des), _ptr(u_des), _ptr(wbc_sol), _ptr(mode_cmd),
                                                _ptr(rbd), None if ld is None else _ptr(ld), _ptr(es), _ptr(cmd), _ptr(tau)),
               "hb_joint_command_batch", self._h)
        return cmd, tau, es

    def rbd_to_centroidal(self, rbd):
        rbd = _f64(rbd); B = rbd.shape[0]
        x = np.zeros((B, NX))
        _check(self._lib.hb_rbd_to_centroidal_batch(self._h, B, _ptr(rbd), _ptr(x)), "hb_rbd_to_centroidal_batch", self._h)
        return x

    def reference_expand(self, t0, refs):
        """refs: ctypes array of HbReference (len B)."""
        t0 = _f64(t0); B = t0.shape[0]
        x_ref = np.zeros((B, self.N + 1, NX)); swing = np.zeros((B, self.N + 1, 24)); mode = np.zeros((B, self.N + 1), dtype=np.int32)
        _check(self._lib.hb_reference_expand_batch(self._h, B, _ptr(t0), C.cast(refs, C.c_void_p), _ptr(x_ref), _ptr(swing), _ptr(mode)), "hb_reference_expand_batch", self._h)
        return x_ref, swing, mode

    def contact_positions(self, x):
        x = _f64(x); B = x.shape[0]
        pos = np.zeros((B, 12))
        _check(self._lib.hb_contact_positions_batch(self._h, B, _ptr(x), _ptr(pos)), "hb_contact_positions_batch", self._h)
        return pos

    def probe_flow_map(self, x, u):
        x, u = _f64(x), _f64(u); B = x.shape[0]
        f = np.zeros((B, NX)); A = np.zeros((B, NX, NX)); Bm = np.zeros((B, NX, NU)); ee = np.zeros((B, 24 + 36 * NX))
        _check(self._lib.hb_probe_flow_map(self._h, B, _ptr(x), _ptr(u), _ptr(f), _ptr(A), _ptr(Bm), _ptr(ee)), "hb_probe_flow_map", self._h)
        out = dict(f=f, A=A, B=Bm, epos=ee[:, :12], evel=ee[:, 12:24], dpos_dx=ee[:, 24:24 + 264].reshape(B, 12, NX),
                   dvel_dx=ee[:, 24 + 264:24 + 528].reshape(B, 12, NX), dvel_du=ee[:, 24 + 528:].reshape(B, 12, NX))
        return out

    # ------------------------------------------------------------------ device-pointer calls (torch cuda tensors, asynchronous)
    def mpc_solve_dev(self, x0, x_ref, swing, mode, xt, ut, info=None):
        B = x0.shape[0]
        _check(self._lib.hb_mpc_solve_batch_dev(self._h, B, _ptr(x0), _ptr(x_ref), _ptr(swing), _ptr(mode), _ptr(xt), _ptr(ut), _ptr(info)), "hb_mpc_solve_batch_dev", self._h)

    def mpc_cold_start_dev(self, x0, mode, xt, ut):
        _check(self._lib.hb_mpc_cold_start_batch_dev(self._h, x0.shape[0], _ptr(x0), _ptr(mode), _ptr(xt), _ptr(ut)), "hb_mpc_cold_start_batch_dev", self._h)

    def wbc_solve_dev(self, x_des, u_des, rbd, mode, stance_mode, sol, status=None):
        _check(self._lib.hb_wbc_solve_batch_dev(self._h, x_des.shape[0], _ptr(x_des), _ptr(u_des), _ptr(rbd), _ptr(mode), _ptr(stance_mode), _ptr(sol), _ptr(status)),
               "hb_wbc_solve_batch_dev", self._h)

    def wbc_qp_dev(self, n, m, H, g, A, lbA, ubA, x, status=None, iters=None):
        _check(self._lib.hb_wbc_qp_batch_dev(self._h, g.shape[0], n, m, _ptr(H), _ptr(g), _ptr(A), _ptr(lbA), _ptr(ubA), _ptr(x), _ptr(status), _ptr(iters)),
               "hb_wbc_qp_batch_dev", self._h)

    def resident_cycle_dev(self, cold_start, t_rel, t0, x0, refs_dev_ptr, rbd, info, sol, tau, status=None):
        _check(self._lib.hb_resident_cycle_batch_dev(self._h, x0.shape[0], 1 if cold_start else 0, C.c_double(t_rel), _ptr(t0), _ptr(x0), C.c_void_p(refs_dev_ptr),
                                                     _ptr(rbd), _ptr(info), _ptr(sol), _ptr(tau), _ptr(status)), "hb_resident_cycle_batch_dev", self._h)

    def control_step_dev(self, t_rel, x0, x_ref, swing, mode, rbd, xt, ut, info, sol, tau, status=None):
        _check(self._lib.hb_control_step_batch_dev(self._h, x0.shape[0], C.c_double(t_rel), _ptr(x0), _ptr(x_ref), _ptr(swing), _ptr(mode), _ptr(rbd), _ptr(xt),
                                                   _ptr(ut), _ptr(info), _ptr(sol), _ptr(tau), _ptr(status)), "hb_control_step_batch_dev", self._h)


# ---------------------------------------------------------------------------------------------------------------------
# Host-side mirrors of the reference operators (single-robot use, B = 1), for drop-in style tests.
class HierarchicalWbc:
    """Mirror of legged::HierarchicalWbc (legged_wbc/include/legged_wbc/HierarchicalWbc.h, src/HierarchicalWbc.cpp:18-31)."""

    def __init__(self, ctx=None):
        self._ctx = ctx or Context(max_batch=1)

    def update(self, stateDesired, inputDesired, rbdStateMeasured, mode, period):
        sol, st = self._ctx.hierarchical_wbc_solve(np.asarray(stateDesired)[None], np.asarray(inputDesired)[None], np.asarray(rbdStateMeasured)[None], [int(mode)])
        if st[0] != 0:
            raise HunterB200Error("[HierarchicalWbc] a level of the hierarchy did not solve (status %d)" % st[0])
        return sol[0]


class WeightedWbc:
    """Mirror of legged::WeightedWbc (legged_wbc/include/legged_wbc/WeightedWbc.h, WbcBase.h:41-76)."""

    def __init__(self, ctx=None):
        self._ctx = ctx or Context(max_batch=1)
        self._stance_mode = True      # WbcBase default until setStanceMode(false) (LeggedController.cpp:161-173)
        self._last = None

    def loadTasksSetting(self, taskFile=None, verbose=False):
        """WbcBase::loadTasksSetting + WeightedWbc::loadTasksSetting (WbcBase.cpp:352-411, WeightedWbc.cpp:96-111): torque limits, friction
        coefficient, task gains and weights from the task file; without a file the shipped values stay in force."""
        if taskFile is not None:
            self._ctx.load_task_info(taskFile)
        if verbose:
            s = self._ctx.wbc_settings()
            print(" #### WBC settings:", {k: (list(getattr(s, k)) if k == "torque_limits" else getattr(s, k)) for k, _ in s._fields_})

    def setKpKd(self, swingKp, swingKd):
        self._ctx.set_kp_kd(swingKp, swingKd)

    def setStanceMode(self, flag):
        self._stance_mode = bool(flag)

    def getContactForceSize(self):
        return 12

    def update(self, stateDesired, inputDesired, rbdStateMeasured, mode, period):
        sol, st = self._ctx.wbc_solve(np.asarray(stateDesired)[None], np.asarray(inputDesired)[None], np.asarray(rbdStateMeasured)[None],
                                      [int(mode)], [1 if self._stance_mode else 0])
        x = sol[0]
        if st[0] != 0:
            print("ERROR: WeightWBC Not Solved!!!")          # WeightedWbc.cpp:57-62
            if self._last is not None:
                x = self._last
        self._last = x.copy()
        return x


class SqpMpc:
    """Mirror of the MPC_MRT_Interface calls the controller makes (LeggedController.cpp:144-156,406) on a batch of 1."""

    def __init__(self, ctx=None, horizon_N=100, dt=0.01):
        self._ctx = ctx or Context(horizon_N=horizon_N, dt=dt, max_batch=1)
        self._xt = None
        self._ut = None
        self._mode = None

    def reset(self):
        self._xt = self._ut = None

    def advance(self, x0, x_ref, swing, mode):
        """One SQP iteration warm-started from the previous solution (mpc.coldStart false, task.info:146)."""
        x0 = np.asarray(x0, dtype=np.float64)[None]
        mode = np.asarray(mode, dtype=np.int32)[None]
        if self._xt is None:
            self._xt, self._ut = self._ctx.mpc_cold_start(x0, mode)
        xt, ut, info = self._ctx.mpc_solve(x0, np.asarray(x_ref)[None], np.asarray(swing)[None], mode, self._xt, self._ut)
        if info["status"][0] != 0:      # a failed iteration must not poison the next warm start: the previous solution stays
            raise HunterB200Error("[SqpMpc] numerical failure in the SQP iteration")   # the reference's MPC thread stops the controller (LeggedController.cpp:413-418)
        self._xt, self._ut, self._mode = xt, ut, mode
        return info[0]

    def evaluatePolicy(self, t_rel):
        s = min(max(t_rel / self._ctx.dt, 0.0), float(self._ctx.N))
        k = min(int(np.floor(s)), self._ctx.N - 1)
        al = s - k
        x = (1 - al) * self._xt[0, k] + al * self._xt[0, k + 1]
        k1 = min(k + 1, self._ctx.N - 1)
        u = (1 - al) * self._ut[0, k] + al * self._ut[0, k1]
        return x, u, int(self._mode[0, k])
