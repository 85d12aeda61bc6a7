"""dev aid: closed-loop standing roll-out with diagnostics (run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import hunter_bipedal_control_b200 as hb
from hunter_bipedal_control_b200 import scenarios as sc
np.set_printoptions(precision=3, suppress=True, linewidth=200)
B, N, dt = 2, 50, 0.02
mode_ff_only = len(sys.argv) > 1 and sys.argv[1] == "ff"
ctx = hb.Context(horizon_N=N, dt=dt, max_batch=B, device=0)
x0 = np.tile(sc.INITIAL_STATE, (B, 1))
rbd = sc.consistent_rbd(x0)
foot_z = ctx.contact_positions(x0).reshape(B, 4, 3)[:, :, 2].min(axis=1)
ground = 0.02
rbd[:, 5] -= foot_z - (ground - 0.001)
z0 = rbd[:, 5].copy()
compacts = []
for i in range(B):
    xi = ctx.rbd_to_centroidal(rbd[i:i + 1])[0]
    c = sc.make_reference(xi, (0.0, 0.0, 0.0, 0.0), "stance", N, dt)[3]
    c["target_times"] = np.array([0.0, 10.0]); c["target_states"][:, 8] = z0[i]
    compacts.append(c)
refs = sc.pack_references(compacts, 3.0)
act = hb.actuation_states(B)
prm = hb.default_sim_params(); prm.ground_height = ground
estop = np.zeros(B, dtype=np.uint8)
period = 0.002
tau_lim = np.tile([28, 60, 60, 60, 28], 2)
for tick in range(200):
    t = tick * period
    if tick % 5 == 0:
        x_meas = ctx.rbd_to_centroidal(rbd)
        info, _, _, st = ctx.resident_cycle(tick == 0, 0.0, np.full(B, t), x_meas, refs, rbd)
    xd, ud, md, sol, tau_ff, st = ctx.resident_wbc(t, rbd)
    cmd, tau_cmd, estop = ctx.joint_command(period, xd, ud, sol, md, rbd, estop=estop)
    tau = ctx.actuation(t, act, cmd, rbd, 0.009)
    tau = np.clip(tau, -tau_lim, tau_lim)
    rbd, cf, fl = ctx.sim_step(rbd, tau, prm)
    if tick % 10 == 0 or estop.any():
        print("tick", tick, "z", rbd[0, 5], "zyx", rbd[0, 0:3], "alpha", info["alpha"][0], "st", st[0], "estop", estop[0])
        print("   qj", rbd[0, 6:16]); print("   xd_j", xd[0, 12:]); print("   tau", tau[0]); print("   tau_ff", tau_ff[0]); print("   Fz", cf[0, 2::3], "ud Fz", ud[0, 2:12:3], "qdd", sol[0, :6])
    if estop.any():
        break
