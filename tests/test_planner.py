"""Host-side reference planner (hb_plan_references) against the literal per-phase restatement in
oracle/refs.py, plus the properties the reference's planner guarantees by construction."""
import ctypes as C

import numpy as np
import pytest

import hunter_bipedal_control_b200 as hb
from hunter_bipedal_control_b200 import scenarios
from oracle import refs as R

N, DT = 40, 0.02
T = N * DT


def _cases(n, seed=3):
    rng = np.random.default_rng(seed)
    x0 = scenarios.random_initial_states(n, seed=seed)
    gaits = [["trot", "standing_trot", "flying_trot", "stance"][i % 4] for i in range(n)]
    cmd = np.stack([rng.uniform(-0.6, 0.8, n), rng.uniform(-0.2, 0.2, n), np.zeros(n), rng.uniform(-0.5, 0.5, n)], axis=1)
    cmd[0] = [0.03, 0.3, 0.0, 0.0]         # exercises the |v| < 0.06 dead band (TargetTrajectoriesPublisher.cpp:117-120)
    t0 = rng.uniform(0.0, 3.0, n)
    start = t0 + rng.uniform(-1.3, 0.3, n)  # gait started before or shortly after the solve time
    feet = np.zeros((n, 4, 3))
    for i in range(n):
        Ry = R.rot_zyx([x0[i, 9], 0, 0])
        for c in range(4):
            feet[i, c] = x0[i, 6:9] + Ry @ np.array(R.FEET_BIAS[c]) + rng.normal(0, 0.01, 3)
    latest = feet + rng.normal(0, 0.02, feet.shape)
    return x0, gaits, cmd, t0, start, feet.reshape(n, 12), latest.reshape(n, 12)


def test_struct_layout_matches_header():
    assert C.sizeof(hb.HbPlanInput) == 8 * (5 + 22 + 4 + 12) + 8


def test_planner_matches_per_phase_restatement():
    n = 24
    x0, gaits, cmd, t0, start, feet, latest = _cases(n)
    refs, ls = hb.plan_references(t0, T, x0, cmd, feet, gaits, start, latest_stance=latest)
    for i in range(n):
        ms, tg, sp = R.plan(t0[i], T, x0[i], cmd[i], feet[i], gaits[i], start[i], latest_stance=latest[i])
        np.testing.assert_allclose(ls[i], sp.latest.reshape(-1), rtol=0, atol=1e-15)
        # nodes + off-grid times, avoiding exact event instants (there the phase lookup differs by which side of the knot is used)
        times = np.concatenate([t0[i] + DT * np.arange(N + 1), t0[i] + np.random.default_rng(i).uniform(0, T, 40)])
        times = np.array([t for t in times if min([abs(t - e) for e in ms.events]) > 1e-7])
        xr, sw, md = R.sample(ms, tg, sp, times)
        xc, sc, mc = R.eval_compact(refs[i], times)
        np.testing.assert_array_equal(md, mc)
        np.testing.assert_allclose(xc, xr, rtol=0, atol=1e-9)     # joint references come out of <= 10 Newton steps per leg
        np.testing.assert_allclose(sc, sw, rtol=0, atol=1e-11)


def test_joint_references_by_ik():
    """P4: the resampled target (0.15 s steps) carries IK joint angles: same numbers as the restatement, inside the joint limits,
    and the toe of the IK configuration is closer to the planned toe position than the seed configuration was."""
    n = 16
    x0, gaits, cmd, t0, start, feet, latest = _cases(n, seed=23)
    refs, _ = hb.plan_references(t0, T, x0, cmd, feet, gaits, start, latest_stance=latest)
    refs0, _ = hb.plan_references(t0, T, x0, cmd, feet, gaits, start, latest_stance=latest, joint_ik=False)
    moved = 0
    for i in range(n):
        ms, tg, sp = R.plan(t0[i], T, x0[i], cmd[i], feet[i], gaits[i], start[i], latest_stance=latest[i])
        ns = int(np.floor(T / 0.15)) + 1
        assert refs[i].n_targets == ns == len(tg.times) and refs0[i].n_targets == 2
        for k in range(ns):
            assert abs(refs[i].target_times[k] - tg.times[k]) < 1e-15
            got = np.array(refs[i].target_states[k][:])
            np.testing.assert_allclose(got, tg.states[k], rtol=0, atol=1e-9)
            assert np.all(got[12:22] >= R.JOINT_LOWER - 1e-15) and np.all(got[12:22] <= R.JOINT_UPPER + 1e-15)
            seed = R.DEFAULT_JOINTS if k == 0 else np.array(refs[i].target_states[k - 1][12:22])
            for leg in range(2):
                des = sp.foot(leg, tg.times[k])[0]
                Rd = R.rot_zyx(x0[i, 9:12])
                q0 = seed[5 * leg:5 * leg + 5]
                q1 = R.translation_ik(got[6:12], q0, leg, des)
                q2 = R.rotation_ik(got[6:12], q1, leg, Rd)
                np.testing.assert_allclose(q2, got[12 + 5 * leg:17 + 5 * leg], rtol=0, atol=1e-9)
                e0 = np.linalg.norm(R.leg_frame(leg, got[6:12], q0)[0] - des)
                e1 = np.linalg.norm(R.leg_frame(leg, got[6:12], q1)[0] - des)
                r1 = np.linalg.norm(R.log3(Rd.T @ R.leg_frame(leg, got[6:12], q1)[1]))
                r2 = np.linalg.norm(R.log3(Rd.T @ R.leg_frame(leg, got[6:12], q2)[1]))
                assert e1 <= e0 + 1e-15 and r2 <= r1 + 1e-15
                if e0 >= 0.01 and e1 < e0:      # below err_tol the reference does not move the leg for position
                    moved += 1
    assert moved > 20


def test_ik_building_blocks():
    """Pivoted-QR basic solution and complete-pivoting kernel: defining properties on random 3x5 Jacobians."""
    rng = np.random.default_rng(2)
    for _ in range(50):
        J = rng.normal(size=(3, 5)); b = rng.normal(size=3)
        x = R.colpiv_qr_solve(J, b)
        assert np.sum(x != 0) == 3 and np.allclose(J @ x, b, atol=1e-10)
        N = R.fullpiv_lu_kernel(J)
        assert N.shape == (5, 2) and np.abs(J @ N).max() < 1e-12 and np.linalg.matrix_rank(N) == 2
    w = np.array([0.3, -0.2, 0.5])
    np.testing.assert_allclose(R.log3(R._rodrigues(w / np.linalg.norm(w), np.linalg.norm(w))), w, atol=1e-14)


def test_planner_continuity_and_footholds():
    n = 12
    x0, gaits, cmd, t0, start, feet, latest = _cases(n, seed=11)
    refs, ls = hb.plan_references(t0, T, x0, cmd, feet, gaits, start, latest_stance=latest)
    for i in range(n):
        times = t0[i] + np.linspace(0, T, 801)
        _, sw, md = R.eval_compact(refs[i], times)
        sw = sw.reshape(-1, 4, 6)
        for c in range(4):
            stance = np.array([R.stance_legs(m)[c] for m in md])
            # planted feet: constant position, zero velocity, on the ground plane next_position_z
            assert np.all(np.abs(sw[stance, c, 3:6]) < 1e-12)
            assert np.all(np.abs(sw[stance, c, 2] - R.NEXT_Z) < 1e-12)
            # positions are continuous (C1 Hermite pieces, stance = end of previous swing)
            assert np.max(np.abs(np.diff(sw[:, c, 0:3], axis=0))) < 0.05
            # swing apex below ground + swingHeight and never under the ground plane by more than the spline undershoot
            assert sw[:, c, 2].max() <= R.NEXT_Z + R.SWING_HEIGHT + 1e-9
            assert sw[:, c, 2].min() >= R.NEXT_Z - 0.01


def test_planner_latest_stance_update():
    """Feet in contact at t0 take the measured position (z forced to next_position_z); swinging feet keep the stored one."""
    x0, gaits, cmd, t0, start, feet, latest = _cases(8, seed=5)
    gaits = ["trot"] * 8
    start = t0 - 0.45           # 0.45 s into a 0.6 s trot period: mode R (1): left contacts swing
    _, ls = hb.plan_references(t0, T, x0, cmd, feet, gaits, start, latest_stance=latest)
    ls = ls.reshape(8, 4, 3); f = feet.reshape(8, 4, 3); l0 = latest.reshape(8, 4, 3)
    for c in (1, 3):
        np.testing.assert_array_equal(ls[:, c, :2], f[:, c, :2])
    for c in (0, 2):
        np.testing.assert_array_equal(ls[:, c, :2], l0[:, c, :2])
    assert np.all(ls[:, :, 2] == R.NEXT_Z)


def test_gait_selector_matches_restatement():
    """P2: 50-sample moving average + thresholds, driven by a speed profile that crosses every threshold in both directions."""
    B, steps = 6, 260
    rng = np.random.default_rng(4)
    sel = hb.GaitSelector(B)
    ref = [R.GaitSelectorRef() for _ in range(B)]
    yaw = rng.uniform(-1, 1, B)
    seen = set()
    for k in range(steps):
        speed = 0.5 * (1 - np.cos(2 * np.pi * k / 130.0)) * np.linspace(0.1, 0.9, B)      # 0 -> peak -> 0 -> peak
        cmd = np.stack([speed, 0.1 * speed, np.zeros(B), 0.3 * speed], axis=1)
        tgt = np.zeros((B, 22)); tgt[:, 9] = yaw
        for i in range(B):
            tgt[i, 0:3] = R.rot_zyx([yaw[i], 0, 0]) @ cmd[i, :3]
        gt = np.array([0, 0, 0, 0, 2, 0], dtype=np.int32)
        level, insert = sel.update(cmd, tgt, gt)
        for i in range(B):
            l, ins = ref[i].update(cmd[i], tgt[i], int(gt[i]))
            assert (level[i], insert[i]) == (l, ins)
            assert abs(sel.vel_avg[i] - ref[i].avg) < 1e-14
            seen.add(int(level[i]))
    assert seen >= {0, 1, 3}


def test_planner_threads_give_identical_bytes():
    n = 200
    x0, gaits, cmd, t0, start, feet, latest = _cases(n, seed=31)
    hb.plan_set_threads(1)
    r1, l1 = hb.plan_references(t0, T, x0, cmd, feet, gaits, start, latest_stance=latest)
    hb.plan_set_threads(3)
    r3, l3 = hb.plan_references(t0, T, x0, cmd, feet, gaits, start, latest_stance=latest)
    hb.plan_set_threads(0)
    assert bytes(r1) == bytes(r3) and np.array_equal(l1, l3)


def test_long_running_gait_is_trimmed_not_truncated():
    """A gait that started minutes ago: whole periods older than t0 - T are skipped (GaitSchedule::getModeSchedule drops them too);
    the window the solver sees is the same as in the untrimmed restatement."""
    n = 6
    x0, gaits, cmd, t0, start, feet, latest = _cases(n, seed=37)
    gaits = ["trot", "flying_trot", "standing_trot", "trot", "flying_trot", "standing_trot"]
    t0 = t0 + 200.0
    start = t0 - np.array([50.3, 77.77, 120.0, 31.0, 64.2, 199.0])
    refs, ls = hb.plan_references(t0, T, x0, cmd, feet, gaits, start, latest_stance=latest, joint_ik=False)
    for i in range(n):
        ms, tg, sp = R.plan(t0[i], T, x0[i], cmd[i], feet[i], gaits[i], start[i], latest_stance=latest[i], joint_ik=False)
        times = t0[i] + np.linspace(0, T, 57)
        times = np.array([t for t in times if min([abs(t - e) for e in ms.events]) > 1e-6])
        xr, sw, md = R.sample(ms, tg, sp, times)
        xc, sc, mc = R.eval_compact(refs[i], times)
        np.testing.assert_array_equal(md, mc)
        np.testing.assert_allclose(sc, sw, rtol=0, atol=1e-9)      # event times accumulate differently over hundreds of periods


def test_planner_rejects_bad_input():
    x0, gaits, cmd, t0, start, feet, latest = _cases(2)
    with pytest.raises(RuntimeError):
        hb.plan_references(t0, -1.0, x0, cmd, feet, gaits, start)
    with pytest.raises(RuntimeError):
        hb.plan_references(t0, T, x0, cmd, feet, gaits, start, prev_event=1e9)


@pytest.mark.gpu
def test_plan_then_expand_on_device():
    """plan (host) -> hb_reference_expand_batch (device) equals the per-phase restatement sampled on the node grid, and
    hb_contact_positions_batch equals the oracle's forward kinematics."""
    from oracle import hbo
    ctx = hb.Context(horizon_N=N, dt=DT, max_batch=32, device=0)
    n = 24
    x0, gaits, cmd, t0, start, _, latest = _cases(n, seed=17)
    feet = ctx.contact_positions(x0)
    for i in range(n):
        ee = hbo.ee_kinematics(x0[i], np.zeros(22))
        np.testing.assert_allclose(feet[i], np.asarray(ee[0]).reshape(-1), rtol=0, atol=1e-12)
    refs, _ = hb.plan_references(t0, T, x0, cmd, feet, gaits, start, latest_stance=latest)
    xr, sw, md = ctx.reference_expand(t0, refs)
    for i in range(n):
        ms, tg, sp = R.plan(t0[i], T, x0[i], cmd[i], feet[i], gaits[i], start[i], latest_stance=latest[i])
        times = t0[i] + DT * np.arange(N + 1)
        ok = np.array([min([abs(t - e) for e in ms.events]) > 1e-7 for t in times])
        xo, so, mo = R.sample(ms, tg, sp, times)
        np.testing.assert_array_equal(md[i][ok], mo[ok])
        np.testing.assert_allclose(xr[i], xo, rtol=0, atol=1e-13)
        np.testing.assert_allclose(sw[i][ok], so[ok], rtol=0, atol=1e-11)
    ctx.close()


@pytest.mark.gpu
def test_joint_command_law():
    """W6: command (posDes, velDes, kp, kd, ff) and output torque per joint, limit protection and pre-load hold."""
    ctx = hb.Context(horizon_N=N, dt=DT, max_batch=64, device=0)
    B = 48
    rng = np.random.default_rng(8)
    x_des = rng.normal(size=(B, 22)); u_des = rng.normal(size=(B, 22)); sol = rng.normal(size=(B, 38)) * 5
    rbd = rng.normal(size=(B, 32)) * 0.3
    rbd[:, 6:16] = np.clip(rbd[:, 6:16] + R.DEFAULT_JOINTS, R.JOINT_LOWER, R.JOINT_UPPER)
    rbd[5, 6 + 3] = R.JOINT_UPPER[3] + 0.05          # knee beyond the limit: emergency stop from joint 3 on
    rbd[6, 6 + 0] = R.JOINT_LOWER[0] - 0.019         # inside the 0.02 margin: no stop
    mode = rng.integers(0, 4, B).astype(np.int32)
    loaded = np.ones(B, dtype=np.uint8); loaded[10:14] = 0
    estop = np.zeros(B, dtype=np.uint8); estop[20] = 1
    cmd, tau, es = ctx.joint_command(0.002, x_des, u_des, sol, mode, rbd, loaded=loaded, estop=estop)
    for i in range(B):
        c, t, e = R.joint_command(0.002, x_des[i], u_des[i], sol[i], int(mode[i]), rbd[i], loaded=bool(loaded[i]), estop=bool(estop[i]))
        np.testing.assert_allclose(cmd[i], c, rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(tau[i], t, rtol=1e-12, atol=1e-12)
        assert bool(es[i]) == e
    assert es[5] == 1 and es[6] == 0 and es[20] == 1
    assert np.all(cmd[5, :3, 2] > 0) and np.all(cmd[5, 3:, 3] == 1.0)
    ctx.close()


def _ref_fields(r):
    ne, nt = r.n_events, r.n_targets
    segs = [[np.array([list(r.segments[c][a][k][:]) for k in range(r.n_segments[c][a])]).reshape(-1, 6) for a in range(3)] for c in range(4)]
    return (ne, nt, np.array(r.event_times[:ne]), np.array(r.modes[:ne + 1]), np.array(r.target_times[:nt]),
            np.array([list(r.target_states[k][:]) for k in range(nt)]), segs)


@pytest.mark.gpu
def test_device_planner_matches_host_planner():
    """Row N1: the planner source compiled for the device (one thread per instance) gives the host planner's plan."""
    ctx = hb.Context(horizon_N=N, dt=DT, max_batch=256, device=0)
    n = 200
    x0, gaits, cmd, t0, start, feet, latest = _cases(n, seed=53)
    ins = hb.make_plan_inputs(t0, T, x0, cmd, feet, gaits, start)
    rd, lsd, st = ctx.plan_references_gpu(ins, latest)
    rh, lsh = hb.plan_references(t0, T, x0, cmd, feet, gaits, start, latest_stance=latest)
    assert (st == 0).all()
    np.testing.assert_allclose(lsd, lsh, rtol=0, atol=1e-14)
    for i in range(n):
        a, b = _ref_fields(rd[i]), _ref_fields(rh[i])
        assert a[0] == b[0] and a[1] == b[1]
        np.testing.assert_allclose(a[2], b[2], rtol=0, atol=1e-12)
        np.testing.assert_array_equal(a[3], b[3])
        np.testing.assert_allclose(a[4], b[4], rtol=0, atol=1e-12)
        np.testing.assert_allclose(a[5], b[5], rtol=0, atol=1e-8)       # IK joint angles: fused multiply-adds on the device
        for c in range(4):
            for ax in range(3):
                assert a[6][c][ax].shape == b[6][c][ax].shape
                np.testing.assert_allclose(a[6][c][ax], b[6][c][ax], rtol=0, atol=1e-11)
    # invalid input is reported per instance and replaced by a safe all-stance reference
    ins[3].horizon = -1.0
    rd, _, st = ctx.plan_references_gpu(ins, latest)
    assert st[3] == -1 and (np.delete(st, 3) == 0).all() and rd[3].n_events == 0 and rd[3].modes[0] == 3
    ctx.close()


@pytest.mark.gpu
def test_plan_cycle_on_device_equals_host_plan_plus_cycle():
    ctx = hb.Context(horizon_N=N, dt=DT, max_batch=512, device=0)
    from hunter_bipedal_control_b200 import scenarios as sc
    for n in (40, 300):
        x0, gaits, cmd, t0, start, _, _ = _cases(n, seed=59)
        start = t0 + np.random.default_rng(n).uniform(0.05, 0.3, n)     # standing at t0 (as at controller start): every foot has a stance history
        cmd = 0.5 * cmd
        rbd = sc.consistent_rbd(x0)
        ins = hb.make_plan_inputs(t0, T, x0, cmd, None, gaits, start)
        info, sol, tau, st, ps = ctx.resident_plan_cycle(True, 0.002, ins, rbd)
        assert (ps == 0).all()
        feet = ctx.contact_positions(x0)
        refs, _ = hb.plan_references(t0, T, x0, cmd, feet, gaits, start, latest_stance=np.zeros((n, 12)))
        info2, sol2, tau2, st2 = ctx.resident_cycle(True, 0.002, t0, x0, refs, rbd)
        assert (st == 0).all() and (st2 == 0).all()
        assert np.abs(tau - tau2).max() < 1e-6 * np.abs(tau2).max()
        assert np.array_equal(info["alpha"], info2["alpha"])
    # second cycle: warm start + planner state (latest stance) carried on the device
    t1 = t0 + 0.02
    ins1 = hb.make_plan_inputs(t1, T, x0, cmd, None, gaits, start)
    info, sol, tau, st, ps = ctx.resident_plan_cycle(False, 0.002, ins1, rbd)
    assert (ps == 0).all() and np.isfinite(tau).all()
    ctx.close()


def test_planner_fuzz_capacity_and_finiteness():
    """Random solve times, gait phases, commands and horizons up to 2 s: the planner never overflows hb_reference for the shipped gaits,
    every number it writes is finite, and horizons beyond the target capacity are refused with -5 instead of truncated."""
    rng = np.random.default_rng(77)
    n = 400
    x0 = scenarios.random_initial_states(n, seed=77)
    gaits = [["trot", "standing_trot", "flying_trot", "stance"][i % 4] for i in range(n)]
    cmd = np.stack([rng.uniform(-0.8, 0.8, n), rng.uniform(-0.3, 0.3, n), np.zeros(n), rng.uniform(-0.8, 0.8, n)], axis=1)
    t0 = rng.uniform(0.0, 500.0, n)
    start = t0 - rng.uniform(-0.3, 300.0, n)
    feet = np.zeros((n, 12))
    for i in range(n):
        Ry = R.rot_zyx([x0[i, 9], 0, 0])
        feet[i] = np.concatenate([x0[i, 6:9] + Ry @ np.array(b) for b in R.FEET_BIAS])
    for horizon in (0.3, 1.0, 2.0):
        refs, ls = hb.plan_references(t0, horizon, x0, cmd, feet, gaits, start, latest_stance=feet)
        assert np.isfinite(ls).all()
        for i in range(0, n, 7):
            r = refs[i]
            assert 0 <= r.n_events <= 32 and 2 <= r.n_targets <= 16
            for k in range(r.n_targets):
                assert np.isfinite(np.array(r.target_states[k][:])).all()
            for c in range(4):
                for a in range(3):
                    ns = r.n_segments[c][a]
                    assert 1 <= ns <= 24
                    seg = np.array([list(r.segments[c][a][s][:]) for s in range(ns)])
                    assert np.isfinite(seg).all() and np.all(seg[:, 1] > seg[:, 0])
                    assert seg[0, 0] <= t0[i] + 1e-9 and seg[-1, 1] >= t0[i] + horizon - 1e-9       # the window is covered
    with pytest.raises(RuntimeError):
        hb.plan_references(t0[:4], 3.0, x0[:4], cmd[:4], feet[:4], gaits[:4], start[:4])          # 21 samples > HB_MAX_TARGETS
