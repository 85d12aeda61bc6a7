import numpy as np, sys
sys.path.insert(0, '.')
import hunter_bipedal_control_b200 as hb
from hunter_bipedal_control_b200 import scenarios as sc
sys.path.insert(0, 'tests')
from test_planner import _cases, N, DT, T
ctx = hb.Context(horizon_N=N, dt=DT, max_batch=512, device=0)
n = 300
x0, gaits, cmd, t0, start, _, _ = _cases(n, seed=59)
rbd = sc.consistent_rbd(x0)
ins = hb.make_plan_inputs(t0, T, x0, cmd, None, gaits, start)
info, sol, tau, st, ps = ctx.resident_plan_cycle(True, 0.002, ins, rbd)
bad = np.nonzero(st)[0]
print("fused: bad", bad, "gaits", [gaits[i] for i in bad], "info status", info["status"][bad], "alpha", info["alpha"][bad])
feet = ctx.contact_positions(x0)
ins2 = hb.make_plan_inputs(t0, T, x0, cmd, feet, gaits, start)
rd, lsd, pst = ctx.plan_references_gpu(ins2, np.zeros((n, 12)))
i2, s2, tau2, st2 = ctx.resident_cycle(True, 0.002, t0, x0, rd, rbd)
print("gpu-plan + cycle: bad", np.nonzero(st2)[0], "info", i2["status"][np.nonzero(st2)[0]])
rh, _ = hb.plan_references(t0, T, x0, cmd, feet, gaits, start, latest_stance=np.zeros((n, 12)))
i3, s3, tau3, st3 = ctx.resident_cycle(True, 0.002, t0, x0, rh, rbd)
print("host-plan + cycle: bad", np.nonzero(st3)[0])
for i in bad[:3]:
    xr, sw, md = ctx.reference_expand(t0[i:i+1], (hb.HbReference * 1)(rd[i]))
    print(i, "expand finite", np.isfinite(xr).all(), np.isfinite(sw).all(), "n_seg", [[rd[i].n_segments[c][a] for a in range(3)] for c in range(4)])
    print("  tau fused", tau[i], "tau host", tau3[i])
# repeat fused to see determinism
info_b, sol_b, tau_b, st_b, ps_b = ctx.resident_plan_cycle(True, 0.002, ins, rbd)
print("fused again: bad", np.nonzero(st_b)[0])
