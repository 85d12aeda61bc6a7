"""Kinematic closed-loop rollout at benchmark size (1024 instances, N=100): statuses and wall time per warm cycle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hunter_bipedal_control_b200 as hb
from hunter_bipedal_control_b200 import scenarios
B, N, DT, cycles = 1024, 100, 0.01, int(sys.argv[1]) if len(sys.argv) > 1 else 30
ctx = hb.Context(horizon_N=N, dt=DT, max_batch=B)
x = scenarios.random_initial_states(B, seed=5); x[:, 0:6] = 0.0
rng = np.random.default_rng(1)
cmd = np.stack([rng.uniform(-0.3, 0.5, B), np.zeros(B), np.zeros(B), rng.uniform(-0.3, 0.3, B)], axis=1)
gaits = [["trot", "standing_trot"][i % 2] for i in range(B)]
t = 0.0; bad = 0; times = []
for c in range(cycles):
    rbd = scenarios.consistent_rbd(x)
    ins = hb.make_plan_inputs(np.full(B, t), N * DT, x, cmd, None, gaits, 0.2)
    t0 = time.perf_counter()
    info, sol, tau, st, ps = ctx.resident_plan_cycle(c == 0, 0.002, ins, rbd)
    times.append(time.perf_counter() - t0)
    nb = int((ps != 0).sum() + (info["status"] != 0).sum() + (st != 0).sum())
    bad += nb
    _, xt, ut = ctx.resident_read(B)
    x = xt[:, 1].copy(); t += DT
    if c % 5 == 0 or nb:
        print("cycle", c, "bad", nb, "alpha<1:", int((info["alpha"] < 1).sum()), "rejected:", int((info["alpha"] == 0).sum()), "max|tau|", float(np.abs(tau).max()), "ms", round(times[-1] * 1e3, 2))
print("total bad", bad, "median cycle ms", round(float(np.median(times[1:])) * 1e3, 3), "height range", float(x[:, 8].min()), float(x[:, 8].max()))
