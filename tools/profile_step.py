"""One warm-up + a few device-resident control steps of the bench workload (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hunter_bipedal_control_b200 as hb
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
x0, x_ref, swing, mode, rbd = bench.workload(B)
ctx = hb.Context(horizon_N=100, dt=0.01, max_batch=B)
dev = torch.device("cuda", 0)
to = lambda a, dt_=torch.float64: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dtype=dt_)
d_x0, d_xref, d_swing, d_rbd, d_mode = to(x0), to(x_ref), to(swing), to(rbd), to(mode, torch.int32)
d_xt0 = torch.zeros((B, 101, 22), dtype=torch.float64, device=dev); d_ut0 = torch.zeros((B, 100, 22), dtype=torch.float64, device=dev)
ctx.mpc_cold_start_dev(d_x0, d_mode, d_xt0, d_ut0); ctx.sync()
d_info = torch.zeros((B, 7), dtype=torch.float64, device=dev); d_sol = torch.zeros((B, 38), dtype=torch.float64, device=dev)
d_tau = torch.zeros((B, 10), dtype=torch.float64, device=dev); d_st = torch.zeros(B, dtype=torch.int32, device=dev)
for _ in range(steps):
    d_xt = d_xt0.clone(); d_ut = d_ut0.clone(); torch.cuda.synchronize()
    ctx.control_step_dev(0.002, d_x0, d_xref, d_swing, d_mode, d_rbd, d_xt, d_ut, d_info, d_sol, d_tau, d_st)
    ctx.sync()
print("done", int((d_st == 0).sum()), "converged; alpha mean", float(d_info[:, 0].mean()))
