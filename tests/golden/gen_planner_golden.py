"""Golden vectors of the reference preprocessing (planner P1/P3/P4/P5) and of the Kalman-filter estimator, generated with the
per-phase / dense restatements in oracle/refs.py (which follow the reference's own object structure; the reference itself is C++
against OCS2 / ROS and cannot be run here). Run from the repo root:  python tests/golden/gen_planner_golden.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import refs as R, hbo
from hunter_bipedal_control_b200 import scenarios

N, DT = 40, 0.02
T = N * DT
n = 8
rng = np.random.default_rng(2024)
x0 = scenarios.random_initial_states(n, seed=2024)
gaits = ["trot", "standing_trot", "flying_trot", "stance", "trot", "flying_trot", "standing_trot", "trot"]
cmd = np.stack([rng.uniform(-0.4, 0.6, n), rng.uniform(-0.1, 0.1, n), np.zeros(n), rng.uniform(-0.4, 0.4, n)], axis=1)
t0 = rng.uniform(0.0, 2.0, n)
start = t0 + rng.uniform(-1.0, 0.25, n)
feet = np.zeros((n, 12)); latest = np.zeros((n, 12))
for i in range(n):
    ee = hbo.ee_kinematics(x0[i], np.zeros(22))[0]
    feet[i] = ee
    latest[i] = ee + rng.normal(0, 0.01, 12)
xr = np.zeros((n, N + 1, 22)); sw = np.zeros((n, N + 1, 24)); md = np.zeros((n, N + 1), dtype=np.int32)
on_event = np.zeros((n, N + 1), dtype=bool); ls_out = np.zeros((n, 12)); tg_t = []; tg_x = []
for i in range(n):
    ms, tg, sp = R.plan(t0[i], T, x0[i], cmd[i], feet[i], gaits[i], start[i], latest_stance=latest[i])
    times = t0[i] + DT * np.arange(N + 1)
    xr[i], sw[i], md[i] = R.sample(ms, tg, sp, times)
    on_event[i] = [min(abs(t - e) for e in ms.events) <= 1e-7 for t in times]
    ls_out[i] = sp.latest.reshape(-1)
# estimator sequence
B, steps, dt = 4, 6, 0.002
kin = lambda q, v: (hbo.rbd(q, v)["cpos"], hbo.rbd(q, v)["J"] @ v)
kf = [R.KalmanFilterRef() for _ in range(B)]
quat = np.zeros((steps, B, 4)); wl = rng.normal(0, 0.4, (steps, B, 3)); al = rng.normal(0, 0.8, (steps, B, 3)) + np.array([0, 0, 9.81])
jpos = np.clip(R.DEFAULT_JOINTS + rng.normal(0, 0.1, (steps, B, 10)), R.JOINT_LOWER, R.JOINT_UPPER); jvel = rng.normal(0, 0.4, (steps, B, 10))
flags = (rng.uniform(size=(steps, B, 4)) > 0.3).astype(np.uint8)
rbd = np.zeros((steps, B, 32)); xh = np.zeros((steps, B, 18)); Pm = np.zeros((steps, B, 18, 18))
for k in range(steps):
    for i in range(B):
        v = rng.normal(0, 0.05, 3)
        quat[k, i] = [v[0], v[1], v[2], np.sqrt(1 - v @ v)]
        rbd[k, i] = kf[i].update(dt, quat[k, i], wl[k, i], al[k, i], jpos[k, i], jvel[k, i], flags[k, i], kin)
        xh[k, i] = kf[i].x; Pm[k, i] = kf[i].P
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "planner_golden.npz"), N=N, DT=DT, x0=x0, gaits=np.array(gaits), cmd=cmd, t0=t0, start=start,
                    feet=feet, latest=latest, x_ref=xr, swing=sw, mode=md, on_event=on_event, latest_out=ls_out,
                    kf_dt=dt, kf_quat=quat, kf_wl=wl, kf_al=al, kf_jpos=jpos, kf_jvel=jvel, kf_flags=flags, kf_rbd=rbd, kf_x=xh, kf_P=Pm)
print("written", xr.shape, rbd.shape)
