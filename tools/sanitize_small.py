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
# planner on the device + resident closed-loop cycle (cold, then warm) + joint command law
from hunter_bipedal_control_b200 import scenarios as sc
B = 6
x0 = sc.random_initial_states(B, seed=3)
rbd = sc.consistent_rbd(x0)
gaits = ["trot", "standing_trot", "flying_trot", "stance", "trot", "flying_trot"]
ins = hb.make_plan_inputs(np.zeros(B), N * dt, x0, (0.3, 0.0, 0.0, 0.1), None, gaits, 0.1)
info, sol, tau, st, ps = ctx.resident_plan_cycle(True, 0.002, ins, rbd)
ins1 = hb.make_plan_inputs(np.full(B, 0.02), N * dt, x0, (0.3, 0.0, 0.0, 0.1), None, gaits, 0.1)
info, sol, tau, st, ps = ctx.resident_plan_cycle(False, 0.002, ins1, rbd)
print("plan cycle: plan status", ps, "wbc status", st, "finite", np.isfinite(tau).all())
feet = ctx.contact_positions(x0)
refs, _ = hb.plan_references(np.zeros(B), N * dt, x0, (0.3, 0.0, 0.0, 0.1), feet, gaits, 0.1)
info, sol, tau, st = ctx.resident_cycle(True, 0.002, np.zeros(B), x0, refs, rbd)
cmd, out_tau, es = ctx.joint_command(0.002, x0, np.zeros((B, 22)), sol, np.full(B, 3, dtype=np.int32), rbd)
print("resident cycle wbc status", st, "command finite", np.isfinite(cmd).all())
# state estimator
st = hb.kf_states(B)
quat = np.tile([0.0, 0.0, 0.0, 1.0], (B, 1))
rb = ctx.estimator_update(0.002, st, quat, np.zeros((B, 3)), np.tile([0, 0, 9.81], (B, 1)), x0[:, 12:22], np.zeros((B, 10)), np.ones((B, 4), dtype=np.uint8))
print("estimator rbd finite", np.isfinite(rb).all(), "height", rb[0, 5])
# round-2 kernels: event-node grid cycle, momentum observer, plant step, actuation model, 500 Hz tick, HoQP / hierarchical WBC, WBC assembly
ev = hb.Context(horizon_N=40, dt=0.015, max_batch=8, time_horizon=0.45, event_nodes=True)
comps = [sc.make_reference(x0[i], (0.3, 0.0, 0.0, 0.1), gaits[i], 30, 0.015, phase=0.011 * (i + 1))[3] for i in range(B)]
erefs = sc.pack_references(comps, 2.0)
for cyc, t_now in enumerate((0.0, 0.01)):
    info, sol, tau, st = ev.resident_cycle(cyc == 0, 0.002, np.full(B, t_now), x0, erefs, rbd)
xd, ud, md, sol2, tau2, st2 = ev.resident_wbc(np.full(B, 0.014), rbd)
tk, nn = ev.resident_read_grid(B)
print("event-node cycle: status", info["status"], st, "intervals", nn, "tick status", st2)
obs = hb.observer_states(B)
est, dist = ctx.contact_force_estimate(0.002, obs, rbd, tau)
act = hb.actuation_states(B)
cmd, out_tau, es = ctx.joint_command(0.002, xd, ud, sol2, md, rbd)
tau_a = ctx.actuation(0.002, act, cmd, rbd)
nxt, cf, fl = ctx.sim_step(rbd, tau_a)
print("observer finite", np.isfinite(est).all(), "plant finite", np.isfinite(nxt).all(), "contacts", fl.sum(axis=1))
hs, hst = ctx.hierarchical_wbc_solve(xd, ud, rbd, md)
H, gq, A, lb, ub, m = ctx.wbc_assemble(xd, ud, rbd, md)
print("hierarchical WBC status", hst, "rows", m)
# pinned reference array: zero-copy gather + validation on the device; shard gather (world 1: pack + compaction kernels).
# Buffers come straight from the CUDA runtime and are released again (torch's caching allocators would show up as leaks in the memcheck log).
import ctypes as C
from cuda.bindings import runtime as cudart
from hunter_bipedal_control_b200 import sharding
nbytes = C.sizeof(refs)
err, hptr = cudart.cudaHostAlloc(nbytes, 0); assert int(err) == 0
C.memmove(hptr, C.addressof(refs), nbytes)
info, sol, tau, st = ctx.resident_cycle(True, 0.002, np.zeros(B), x0, C.cast(C.c_void_p(hptr), C.POINTER(hb.HbReference)), rbd)
shard = sharding.Shard(ctx, None, 1, 0, B, max_row_doubles=10)
err, dptr = cudart.cudaMalloc(tau.nbytes); assert int(err) == 0
err, = cudart.cudaMemcpy(dptr, tau.ctypes.data, tau.nbytes, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice); assert int(err) == 0
out = C.c_void_p()
hb.api._check(ctx._lib.hb_shard_gather_dev(shard._h, 10, C.c_void_p(dptr), None, C.byref(out)), "hb_shard_gather_dev", ctx._h)
shard.wait(block_host=True)
print("pinned cycle status", st, "gather equal", np.array_equal(shard.to_host(out.value, 10), tau))
shard.close()
cudart.cudaFree(dptr); cudart.cudaFreeHost(hptr)
ev.close(); ctx.close()
print("contexts closed")
