#!/usr/bin/env python3
"""Generate tests/golden/path_golden.npz with the CPU oracle: small seeded input/output vectors of the hot path
(one SQP iteration on N=12 horizons for four gaits, WeightedWbc solutions in all modes, control-step torques).
The reference itself cannot run here (OCS2/Pinocchio/qpOASES absent), so these goldens are ORACLE-generated: they pin the
oracle against regressions and give the GPU tests fixed vectors; they are not outputs of the upstream binaries."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.normpath(os.path.join(HERE, "..", "..")))
from oracle import hbo
from hunter_bipedal_control_b200 import scenarios as S

N, dt = 12, 0.025
gaits = ["stance", "trot", "standing_trot", "flying_trot", "trot", "trot"]
B = len(gaits)
x0 = S.random_initial_states(B, seed=777)
x_ref = np.zeros((B, N + 1, 22)); swing = np.zeros((B, N + 1, 24)); mode = np.zeros((B, N + 1), dtype=np.int32)
for i, g in enumerate(gaits):
    x_ref[i], swing[i], mode[i], _ = S.make_reference(x0[i], (0.2, 0.05 * i, 0, 0.1 * i), g, N, dt, phase=0.07 * i)
xt0 = np.zeros((B, N + 1, 22)); ut0 = np.zeros((B, N, 22))
xt1 = np.zeros_like(xt0); ut1 = np.zeros_like(ut0); xt2 = np.zeros_like(xt0); ut2 = np.zeros_like(ut0)
alpha = np.zeros((B, 2)); merit = np.zeros((B, 2)); viol = np.zeros((B, 2))
for i in range(B):
    xt0[i], ut0[i] = hbo.mpc_cold_start(N, dt, x0[i], mode[i])
    xt1[i], ut1[i], i1 = hbo.mpc_iteration(N, dt, x0[i], x_ref[i], swing[i], mode[i], xt0[i], ut0[i])
    xt2[i], ut2[i], i2 = hbo.mpc_iteration(N, dt, x0[i], x_ref[i], swing[i], mode[i], xt1[i], ut1[i])
    alpha[i] = [i1["alpha"], i2["alpha"]]; merit[i] = [i1["merit1"], i2["merit1"]]; viol[i] = [i1["viol1"], i2["viol1"]]
# WBC cases
rng = np.random.default_rng(778)
W = 12
wmode = np.array([3, 3, 2, 1, 0, 2, 1, 3, 2, 1, 3, 0], dtype=np.int32)
wstance = np.array([1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0], dtype=np.uint8)
wx = np.tile(S.INITIAL_STATE, (W, 1)) + rng.uniform(-.05, .05, (W, 22))
wu = np.zeros((W, 22))
for i in range(W):
    fl = S.mode_flags(int(wmode[i]))
    for c in range(4):
        if fl[c]:
            wu[i, 3 * c + 2] = S.TOTAL_MASS * 9.81 / sum(fl)
    wu[i, 12:] = rng.uniform(-.5, .5, 10)
wrbd = S.consistent_rbd(wx, rng, 0.02)
wsol = np.zeros((W, 38))
for i in range(W):
    wsol[i], st = hbo.wbc_solve(wx[i], wu[i], wrbd[i], int(wmode[i]), bool(wstance[i]), 1e-8)
    assert st == 0
np.savez_compressed(os.path.join(HERE, "path_golden.npz"), N=N, dt=dt, x0=x0, x_ref=x_ref, swing=swing, mode=mode, xt0=xt0, ut0=ut0, xt1=xt1, ut1=ut1,
                    xt2=xt2, ut2=ut2, alpha=alpha, merit=merit, viol=viol, wmode=wmode, wstance=wstance, wx=wx, wu=wu, wrbd=wrbd, wsol=wsol)
print("wrote path_golden.npz", alpha.tolist())
