"""dev aid: level-0 lifted QP of the hierarchical WBC through the generic device QP vs the oracle's IPM."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import hunter_bipedal_control_b200 as hb
from oracle import hbo
from oracle.hoqp import hierarchical_wbc
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from test_gpu_hoqp import _wbc_cases
np.set_printoptions(precision=3, suppress=False, linewidth=200)
ctx = hb.Context(horizon_N=4, max_batch=16, wbc_rho=1e-10, qp_max_iter=80)
x, u, rbd, mode = _wbc_cases(8, 4)
for i in range(8):
    so, levels, tasks = hierarchical_wbc(x[i], u[i], rbd[i], int(mode[i]))
    t0 = tasks[0]
    nx, nv = 38, t0.d.shape[0]
    H = np.zeros((nx + nv, nx + nv)); H[:nx, :nx] = t0.a.T @ t0.a + 1e-12 * np.eye(nx); H[nx:, nx:] = np.eye(nv)
    c = np.concatenate([-t0.a.T @ t0.b, np.zeros(nv)])
    D = np.zeros((2 * nv, nx + nv)); D[:nv, nx:] = -np.eye(nv); D[nv:, :nx] = t0.d; D[nv:, nx:] = -np.eye(nv)
    f = np.concatenate([np.zeros(nv), t0.f])
    xo, sto, ito = hbo.qp_solve(H, c, D, np.full(f.size, -1e20), f, 1e-10)
    xs, st, it = ctx.wbc_qp(H[None], c[None], D[None], np.full((1, f.size), -1e20), f[None])
    ro = np.abs(t0.a @ xo[:nx] - t0.b).max(); rd = np.abs(t0.a @ xs[0, :nx] - t0.b).max()
    print(i, "mode", mode[i], "oracle st/it", sto, ito, "res", ro, "| dev st/it", st[0], it[0], "res", rd, "slack max", np.abs(xs[0, nx:]).max(), np.abs(xo[nx:]).max(),
          "viol", (t0.d @ xs[0, :nx] - t0.f).max(), "dx", np.abs(xs[0] - xo).max())
