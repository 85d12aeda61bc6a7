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
        np.testing.assert_allclose(xc, xr, rtol=0, atol=1e-13)
        np.testing.assert_allclose(sc, sw, rtol=0, atol=1e-11)


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
