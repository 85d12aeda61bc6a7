"""Quick GPU-vs-oracle parity report (development aid; the parity tests proper live in tests/)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import hunter_bipedal_control_b200 as hb
from hunter_bipedal_control_b200 import scenarios as S
from oracle import hbo

N, dt, B = 100, 0.01, 8
ctx = hb.Context(horizon_N=N, dt=dt, max_batch=64)
rng = np.random.default_rng(1)
x0, x_ref, swing, mode = S.make_batch(B, N, dt, gait="trot")
u0 = np.zeros((B, 22)); u0[:, [2, 5, 8, 11]] = S.TOTAL_MASS * 9.81 / 4; u0 += rng.uniform(-1, 1, (B, 22))
pr = ctx.probe_flow_map(x0, u0)
for i in range(2):
    f, A, Bm = hbo.flow_map(x0[i], u0[i])
    pos, vel, dp, dvx, dvu = hbo.ee_kinematics(x0[i], u0[i])
    print("flow map", i, "f %.1e A %.1e B %.1e pos %.1e vel %.1e dp %.1e dvx %.1e dvu %.1e" % (
        np.abs(f - pr["f"][i]).max(), np.abs(A - pr["A"][i]).max(), np.abs(Bm - pr["B"][i]).max(), np.abs(pos - pr["epos"][i]).max(),
        np.abs(vel - pr["evel"][i]).max(), np.abs(dp - pr["dpos_dx"][i]).max(), np.abs(dvx - pr["dvel_dx"][i]).max(), np.abs(dvu - pr["dvel_du"][i]).max()))
# WBC
rbd = S.consistent_rbd(x0, rng, 0.02)
modes = np.array([3, 2, 1, 3, 2, 1, 2, 1], dtype=np.int32)
ud = u0.copy()
sol, st = ctx.wbc_solve(x0, ud, rbd, modes, np.zeros(B, dtype=np.uint8))
for i in range(B):
    so, sto = hbo.wbc_solve(x0[i], ud[i], rbd[i], int(modes[i]), False, 1e-8)
    print("wbc", i, "mode", modes[i], "status", st[i], sto, "tau rel err %.2e  x rel err %.2e" % (
        np.abs(so[28:] - sol[i, 28:]).max() / max(1, np.abs(so[28:]).max()), np.abs(so - sol[i]).max() / np.abs(so).max()))
# MPC
xt, ut = ctx.mpc_cold_start(x0, mode)
xo_c, uo_c = hbo.mpc_cold_start(N, dt, x0[0], mode[0])
print("cold start", np.abs(xo_c - xt[0]).max(), np.abs(uo_c - ut[0]).max())
for it in range(3):
    t = time.time()
    xt1, ut1, info = ctx.mpc_solve(x0, x_ref, swing, mode, xt, ut)
    tg = time.time() - t
    for i in range(3):
        xo, uo, io = hbo.mpc_iteration(N, dt, x0[i], x_ref[i], swing[i], mode[i], xt[i], ut[i])
        print("mpc it", it, "inst", i, "gpu alpha %.4f merit %.6f->%.6f viol %.3e->%.3e | oracle alpha %.4f merit %.6f->%.6f viol %.3e->%.3e | dx %.2e du %.2e armijo %.3e/%.3e" % (
            info["alpha"][i], info["merit0"][i], info["merit1"][i], info["viol0"][i], info["viol1"][i], io["alpha"], io["merit0"], io["merit1"], io["viol0"], io["viol1"],
            np.abs(xo - xt1[i]).max(), np.abs(uo - ut1[i]).max(), info["armijo"][i], io["armijo"]))
    print("  gpu call %.3fs" % tg)
    xt, ut = xt1, ut1
print("launches", ctx.launch_count)
