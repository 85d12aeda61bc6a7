"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): golden-size MPC + WBC through the C ABI."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hunter_bipedal_control_b200 as hb
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "path_golden.npz"))
N, dt = int(g["N"]), float(g["dt"])
ctx = hb.Context(horizon_N=N, dt=dt, max_batch=16)
xt, ut = ctx.mpc_cold_start(g["x0"], g["mode"])
xt1, ut1, info = ctx.mpc_solve(g["x0"], g["x_ref"], g["swing"], g["mode"], xt, ut)
print("alpha", info["alpha"], "max dx", np.abs(xt1 - g["xt1"]).max())
sol, st = ctx.wbc_solve(g["wx"], g["wu"], g["wrbd"], g["wmode"], g["wstance"])
print("wbc status", st, "max err", np.abs(sol - g["wsol"]).max())
