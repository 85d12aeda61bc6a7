"""dev aid: device WBC at rho = 1e-8 / 1e-9 / 1e-10 against the exact two-stage least-norm optimum (numpy)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hunter_bipedal_control_b200 as hb
from oracle import hbo
from test_oracle_solvers import wbc_case, exact_least_norm_optimum
cases = []
for mode in (3, 2, 1, 0):
    for seed in range(12):
        rng = np.random.default_rng(90 + mode + 10 * seed)
        cases.append((mode,) + wbc_case(hbo, rng, mode))
mode = np.array([c[0] for c in cases], dtype=np.int32); xd = np.array([c[1] for c in cases]); ud = np.array([c[2] for c in cases]); rbd = np.array([c[3] for c in cases])
exact = []
for m, x, u, r in cases:
    H, g, A, lb, ub = hbo.wbc_assemble(x, u, r, m, False)
    xi, st, _ = hbo.qp_solve(H, g, A, lb, ub, 1e-9)
    exact.append(exact_least_norm_optimum(H, g, A, lb, ub, xi)[0])
exact = np.array(exact)
for rho in (1e-8, 3e-9, 1e-9, 1e-10):
    ctx = hb.Context(horizon_N=1, max_batch=1024, wbc_rho=rho)
    sol, st = ctx.wbc_solve(xd, ud, rbd, mode, np.zeros(len(cases), dtype=np.uint8))
    rel = np.abs(sol[:, 28:] - exact[:, 28:]).max(axis=1) / np.maximum(1.0, np.abs(exact[:, 28:]).max(axis=1))
    # timing on 1024 problems
    reps = 1024 // len(cases) + 1
    X = np.tile(xd, (reps, 1))[:1024]; U = np.tile(ud, (reps, 1))[:1024]; R = np.tile(rbd, (reps, 1))[:1024]; M = np.tile(mode, reps)[:1024]
    ctx.wbc_solve(X, U, R, M); ctx.profile_enable(True)
    for _ in range(5):
        s2, st2 = ctx.wbc_solve(X, U, R, M)
    pr = ctx.profile_read()
    print("rho", rho, "status ok", int((st == 0).sum()), "/", len(cases), "worst rel vs exact", rel.max(), "median", np.median(rel), "kernel ms/1024", pr["qp_ipm"]["ms"] / 5, "solved 1024:", int((st2 == 0).sum()))
    ctx.close()
