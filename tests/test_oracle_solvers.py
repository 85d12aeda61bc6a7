"""Pins the oracle's solver layer: the dense QP by an optimality certificate (KKT conditions on the recovered active set),
the WBC assembly by physical facts, and the projected Riccati step by one dense KKT solve of the same equality-constrained
LQ problem."""
import numpy as np
import pytest

X0 = np.array([0, 0, 0, 0, 0, 0, 0, 0, 0.63, 0, 0, 0, .1, 0, .4, .93, .53, -.1, 0, -.4, .93, -.53])
M = 12.586944


def wbc_case(oracle, rng, mode):
    xd = X0 + rng.uniform(-.05, .05, 22)
    ud = np.zeros(22)
    fl = [mode in (2, 3), mode in (1, 3), mode in (2, 3), mode in (1, 3)]
    for c in range(4):
        if fl[c]:
            ud[3 * c + 2] = M * 9.81 / sum(fl)
    ud[12:] = rng.uniform(-.5, .5, 10)
    q = xd[6:] + rng.uniform(-.02, .02, 16)
    rbd = np.r_[q[3:6], q[0:3], q[6:], rng.uniform(-.3, .3, 16)]
    return xd, ud, rbd


def kkt_certificate(H, g, A, lb, ub, rho, x, tol=1e-6):
    n = len(g)
    Hr = H + rho * np.eye(n)
    Ax = A @ x
    assert (Ax <= ub + 1e-7 * (1 + np.abs(ub))).all() and (Ax >= lb - 1e-7 * (1 + np.abs(lb))).all()
    eq = np.isclose(lb, ub) & (np.abs(A).max(axis=1) > 0)
    act_u = (~eq) & (ub < 1e19) & (np.abs(Ax - ub) < 1e-6) & (np.abs(A).max(axis=1) > 0)
    act_l = (~eq) & (lb > -1e19) & (np.abs(Ax - lb) < 1e-6) & (np.abs(A).max(axis=1) > 0)
    rows = np.vstack([A[eq], A[act_u], -A[act_l]])
    grad = Hr @ x + g
    lam, *_ = np.linalg.lstsq(rows.T, -grad, rcond=None)
    res = np.abs(rows.T @ lam + grad).max()
    assert res < tol * (1 + np.abs(g).max()), res
    lam_in = lam[eq.sum():]
    assert (lam_in > -1e-6 * (1 + np.abs(lam).max())).all()


@pytest.mark.parametrize("mode,stance", [(3, True), (3, False), (2, False), (1, False), (0, False)])
def test_wbc_qp_optimality(oracle, mode, stance):
    rng = np.random.default_rng(10 + mode)
    for _ in range(3):
        xd, ud, rbd = wbc_case(oracle, rng, mode)
        H, g, A, lb, ub = oracle.wbc_assemble(xd, ud, rbd, mode, stance)
        nsw = 4 - sum([mode in (2, 3), mode in (1, 3)]) * 2
        assert A.shape == (16 + 3 * nsw + 20 + 5 * (4 - nsw) + 3 * nsw, 38)      # 56 / 58 / 60 rows
        assert np.linalg.matrix_rank(H) <= min(16, 6 + 3 * nsw if not stance else 6)
        for rho in (1e-6, 1e-8):
            x, st, it = oracle.qp_solve(H, g, A, lb, ub, rho)
            assert st == 0 and it < 40
            kkt_certificate(H, g, A, lb, ub, rho, x)
            # EoM rows hold to 1e-9
            assert np.abs(A[:16] @ x - ub[:16]).max() < 1e-8
            # friction pyramid and torque limits
            for c in range(4):
                F = x[16 + 3 * c:19 + 3 * c]
                assert F[2] > -1e-7 and abs(F[0]) <= 0.7 * F[2] + 1e-6 and abs(F[1]) <= 0.7 * F[2] + 1e-6
            assert (np.abs(x[28:]) <= np.tile([28, 60, 60, 60, 28], 2) + 1e-6).all()


def test_wbc_stance_mode_at_rest(oracle):
    """Before /set_walk the task is qdd_base = 0 (WeightedWbc.cpp:83-94): at rest the base acceleration vanishes and
    the contact forces carry the weight."""
    u = np.zeros(22); u[[2, 5, 8, 11]] = M * 9.81 / 4
    rbd = np.r_[X0[9:12], X0[6:9], X0[12:], np.zeros(16)]
    x, st = oracle.wbc_solve(X0, u, rbd, 3, True, 1e-8)
    assert st == 0
    assert np.abs(x[:6]).max() < 1e-4
    assert abs(x[[18, 21, 24, 27]].sum() - M * 9.81) < 0.05   # joint accelerations of the least-norm point carry a little momentum


def test_least_norm_limit(oracle):
    """The Tikhonov weight selects (to O(rho)) the least-norm point of the optimal face: ||x|| is non-increasing in rho and the
    torques converge as rho -> 0 (qpOASES setToMPC regularisation)."""
    rng = np.random.default_rng(5)
    xd, ud, rbd = wbc_case(oracle, rng, 2)
    H, g, A, lb, ub = oracle.wbc_assemble(xd, ud, rbd, 2, False)
    xs = [oracle.qp_solve(H, g, A, lb, ub, r)[0] for r in (1e-7, 1e-8, 1e-9)]
    assert np.abs(xs[1][28:] - xs[2][28:]).max() < 1e-4 * np.abs(xs[2][28:]).max()
    assert np.abs(xs[0][28:] - xs[1][28:]).max() < 1e-3 * np.abs(xs[2][28:]).max()


def test_random_qp_against_enumeration(oracle):
    """Small strictly convex QPs: the IPM optimum equals the best vertex of the active-set enumeration."""
    import itertools
    rng = np.random.default_rng(7)
    for _ in range(5):
        n, m = 4, 5
        Mx = rng.normal(size=(n, n)); H = Mx @ Mx.T + 0.1 * np.eye(n); g = rng.normal(size=n)
        A = rng.normal(size=(m, n)); ub = rng.uniform(0.1, 1, m); lb = -1e20 * np.ones(m)
        x, st, _ = oracle.qp_solve(H, g, A, lb, ub, 0.0)
        assert st == 0
        best = None
        for k in range(m + 1):
            for S in itertools.combinations(range(m), k):
                S = list(S)
                K = np.block([[H, A[S].T], [A[S], np.zeros((k, k))]])
                try:
                    sol = np.linalg.solve(K, np.r_[-g, ub[S]])
                except np.linalg.LinAlgError:
                    continue
                xx, lam = sol[:n], sol[n:]
                if (A @ xx <= ub + 1e-9).all() and (lam >= -1e-9).all():
                    f = 0.5 * xx @ H @ xx + g @ xx
                    if best is None or f < best[0]:
                        best = (f, xx)
        assert best is not None and np.abs(best[1] - x).max() < 1e-6


def test_riccati_against_dense_kkt(oracle):
    """One SQP iteration (projection + Riccati + forward pass) equals the solution of the same LQ subproblem assembled
    as one dense KKT system, on a short horizon with a mode switch."""
    N, dt = 6, 0.01
    rng = np.random.default_rng(11)
    mode = np.array([3, 3, 2, 2, 1, 1, 1], dtype=np.int32)
    x0 = X0 + rng.uniform(-.02, .02, 22)
    xref = np.tile(X0, (N + 1, 1)); swing = np.zeros((N + 1, 4, 6)); swing[:, :, 2] = 0.02
    pos, *_ = oracle.ee_kinematics(x0, np.zeros(22))
    swing[:, :, 0:2] = pos.reshape(4, 3)[None, :, 0:2]
    swing = swing.reshape(N + 1, 24)
    xt, ut = oracle.mpc_cold_start(N, dt, x0, mode)
    xt = xt + rng.uniform(-.01, .01, xt.shape); xt[0] = x0
    ut = ut + rng.uniform(-.5, .5, ut.shape)
    # dense problem: variables [dx_1..dx_N, du_0..du_{N-1}], dx_0 = 0
    nx, nu = 22, 22
    nv = N * nx + N * nu
    Hh = np.zeros((nv, nv)); gg = np.zeros(nv)
    rows = []; rhs = []
    ix = lambda k: slice((k - 1) * nx, k * nx)
    iu = lambda k: slice(N * nx + k * nu, N * nx + (k + 1) * nu)
    for k in range(N):
        lq = oracle.node_lq(dt, xt[k], ut[k], xt[k + 1], xref[k], swing[k], int(mode[k]))
        m = lq["m"]
        Hh[iu(k), iu(k)] += dt * lq["R"]; gg[iu(k)] += dt * lq["r"]
        if k > 0:
            Hh[ix(k), ix(k)] += dt * lq["Q"]; gg[ix(k)] += dt * lq["q"]
            Hh[iu(k), ix(k)] += dt * lq["P"]; Hh[ix(k), iu(k)] += dt * lq["P"].T
        row = np.zeros((nx, nv)); row[:, ix(k + 1)] = -np.eye(nx); row[:, iu(k)] = lq["Bd"]
        if k > 0:
            row[:, ix(k)] = lq["Ad"]
        rows.append(row); rhs.append(-lq["b"])
        D, Cm, e = lq["D"][:m], lq["C"][:m], lq["e"][:m]
        row = np.zeros((nu, nv)); row[:, iu(k)] = D.T @ D
        if k > 0:
            row[:, ix(k)] = D.T @ Cm
        rows.append(row); rhs.append(-D.T @ e)
    Aeq = np.vstack(rows); beq = np.concatenate(rhs)
    # null-space method on the (rank-deficient but consistent) constraints
    U, s, Vt = np.linalg.svd(Aeq)
    r = (s > 1e-9 * s[0]).sum()
    xp = Vt[:r].T @ ((U[:, :r].T @ beq) / s[:r])
    Z = Vt[r:].T
    z = np.linalg.solve(Z.T @ Hh @ Z, -Z.T @ (gg + Hh @ xp))
    sol = xp + Z @ z
    xt1, ut1, info = oracle.mpc_iteration(N, dt, x0, xref, swing, mode, xt, ut)
    assert info["alpha"] > 0
    dx = (xt1 - xt) / info["alpha"]; du = (ut1 - ut) / info["alpha"]
    dx[0] = 0
    for k in range(1, N + 1):
        assert np.abs(dx[k] - sol[ix(k)]).max() < 1e-6 * max(1, np.abs(sol).max()), k
    for k in range(N):
        assert np.abs(du[k] - sol[iu(k)]).max() < 1e-6 * max(1, np.abs(sol).max()), k


def test_sqp_converges_on_stance(oracle):
    N, dt = 40, 0.01
    mode = np.full(N + 1, 3, dtype=np.int32)
    xref = np.tile(X0, (N + 1, 1)); swing = np.zeros((N + 1, 24))
    xt, ut = oracle.mpc_cold_start(N, dt, X0, mode)
    viol = []
    for _ in range(4):
        xt, ut, info = oracle.mpc_iteration(N, dt, X0, xref, swing, mode, xt, ut)
        assert info["status"] == 0 and info["alpha"] == 1.0
        viol.append(info["viol1"])
    assert viol[-1] < 1e-6 and viol[1] < 0.1 * viol[0]


@pytest.mark.parametrize("mode", [3, 1])
def test_wbc_qp_against_independent_solver(oracle, mode):
    """The oracle's interior-point solution of the WeightedWbc QP against a solver that shares no code with it (SciPy trust-constr
    with the exact Hessian on the same regularised problem). qpOASES itself is not available offline; this pins the QP layer to a
    third-party implementation."""
    from scipy.optimize import minimize, LinearConstraint
    rng = np.random.default_rng(40 + mode)
    xd, ud, rbd = wbc_case(oracle, rng, mode)
    H, g, A, lb, ub = oracle.wbc_assemble(xd, ud, rbd, mode, False)
    rho = 1e-6
    x_ip, st, _ = oracle.qp_solve(H, g, A, lb, ub, rho)
    assert st == 0
    Hr = H + rho * np.eye(38)
    keep = np.abs(A).sum(axis=1) > 0
    A, lb, ub = A[keep], np.where(lb[keep] < -1e19, -np.inf, lb[keep]), np.where(ub[keep] > 1e19, np.inf, ub[keep])
    res = minimize(lambda x: 0.5 * x @ Hr @ x + g @ x, np.zeros(38), jac=lambda x: Hr @ x + g, hess=lambda x: Hr,
                   constraints=[LinearConstraint(A, lb, ub)], method="trust-constr", options={"maxiter": 3000, "gtol": 1e-10, "xtol": 1e-14})
    f_ip = 0.5 * x_ip @ Hr @ x_ip + g @ x_ip
    viol = max(np.maximum(A @ res.x - ub, 0).max(), np.maximum(lb - A @ res.x, 0).max())
    assert viol < 1e-7
    assert abs(res.fun - f_ip) <= 1e-6 * max(1.0, abs(f_ip))                 # same optimal value
    # the optimal face is nearly flat (rank-deficient H), so a solver that stops at a relative objective accuracy of 1e-6 cannot pin the
    # argmin; test_least_norm_optimum_exact_two_stage below does that exactly


def exact_least_norm_optimum(H, g, A, lb, ub, x_seed):
    """argmin ||x||^2 over argmin of the QP, exactly (no regularisation), by a dense active-set method in numpy.
    Stage 1: an exact optimum of the unregularised QP from the KKT system on the active set (H is singular: x1 is one point of the optimal
    face, H x1 is the face's unique value). Stage 2: min ||x||^2 over the face {rows with a positive multiplier active, H x = H x1, all
    other inequalities kept as inequalities}: a strictly convex QP solved by working-set iteration (add the most violated row, drop rows with
    a negative multiplier). x_seed (any near-optimal point) only seeds the active set; the caller certifies the result on its own."""
    from scipy.optimize import nnls
    keep = np.abs(A).sum(axis=1) > 0
    A, lb, ub = A[keep], lb[keep], ub[keep]
    eq = lb == ub
    Ax = A @ x_seed
    act_u = (~eq) & (ub < 1e19) & (np.abs(Ax - ub) < 1e-7 * (1 + np.abs(ub)))
    act_l = (~eq) & (lb > -1e19) & (np.abs(Ax - lb) < 1e-7 * (1 + np.abs(lb)))
    Aeq, Au, Al = A[eq], A[act_u], A[act_l]

    def multipliers(x):
        au = (~eq) & (ub < 1e19) & (np.abs(A @ x - ub) < 1e-8 * (1 + np.abs(ub))); al = (~eq) & (lb > -1e19) & (np.abs(A @ x - lb) < 1e-8 * (1 + np.abs(lb)))
        Mx = np.hstack([Aeq.T, -Aeq.T, A[au].T, -A[al].T])
        lam, res = nnls(Mx, -(H @ x + g), maxiter=20000)
        return res
    Mx = np.hstack([Aeq.T, -Aeq.T, Au.T, -Al.T])
    lam, _ = nnls(Mx, -(H @ x_seed + g), maxiter=20000)
    ne = Aeq.shape[0]
    lu, ll = lam[2 * ne:2 * ne + Au.shape[0]], lam[2 * ne + Au.shape[0]:]
    su, sl = lu > 1e-8 * max(1.0, np.abs(lam).max()), ll > 1e-8 * max(1.0, np.abs(lam).max())
    Ca = np.vstack([Aeq, Au[su], Al[sl]]); da = np.concatenate([ub[eq], ub[act_u][su], lb[act_l][sl]])
    n, mc = H.shape[0], Ca.shape[0]
    K = np.block([[H, Ca.T], [Ca, np.zeros((mc, mc))]])
    x1 = np.linalg.lstsq(K, np.concatenate([-g, da]), rcond=1e-13)[0][:n]
    C = np.vstack([Ca, H]); d = np.concatenate([da, H @ x1])
    # stage 2: remaining inequalities G x <= h
    G = np.vstack([A[(~eq) & (ub < 1e19)], -A[(~eq) & (lb > -1e19)]]); h = np.concatenate([ub[(~eq) & (ub < 1e19)], -lb[(~eq) & (lb > -1e19)]])
    W = []
    for _ in range(200):
        Cw = np.vstack([C, G[W]]) if W else C
        dw = np.concatenate([d, h[W]]) if W else d
        x = np.linalg.lstsq(Cw, dw, rcond=1e-12)[0]                 # least-norm point of {Cw x = dw}
        viol = G @ x - h
        viol[W] = -np.inf
        j = int(np.argmax(viol))
        if viol[j] > 1e-10 * (1 + abs(h[j])):
            W.append(j)
            continue
        if W:       # multipliers of the working rows in x = Cw' mu (x = -Cw' nu / 2): a row may leave when its multiplier has the wrong sign
            mu = np.linalg.lstsq(Cw.T, x, rcond=1e-12)[0][C.shape[0]:]
            k = int(np.argmax(mu))
            if mu[k] > 1e-10 * max(1.0, np.abs(mu).max()):
                W.pop(k)
                continue
        break
    Cf = np.vstack([C, G[W]]) if W else C
    return x, dict(A=A, lb=lb, ub=ub, C=Cf, stationarity=multipliers, working=W)


@pytest.mark.parametrize("mode", [3, 2, 1, 0])
def test_least_norm_optimum_exact_two_stage(oracle, mode):
    """The definition the kernels are held to -- x* = least-norm point of the optimal face (qpOASES setToMPC) -- computed
    exactly in two stages with numpy (dense active set) and certified on its own. The interior point with the Tikhonov weight rho converges to
    it at O(rho): torques within 5e-4 relative at rho = 1e-8 (the product setting; worst case measured 1.3e-4, median 4e-6) and 5e-5 at
    rho = 1e-9. Replaces the former 5 % cross-check against a loosely converged third-party solver."""
    e8, e9 = [], []
    for seed in range(4):
        rng = np.random.default_rng(90 + mode + 10 * seed)
        xd, ud, rbd = wbc_case(oracle, rng, mode)
        H, g, A, lb, ub = oracle.wbc_assemble(xd, ud, rbd, mode, False)
        x9, st9, _ = oracle.qp_solve(H, g, A, lb, ub, 1e-9)
        x8, st8, _ = oracle.qp_solve(H, g, A, lb, ub, 1e-8)
        assert st9 == 0 and st8 == 0
        x, c = exact_least_norm_optimum(H, g, A, lb, ub, x9)
        Ak, lbk, ubk = c["A"], c["lb"], c["ub"]
        sc_ = max(1.0, np.abs(x).max())
        # certificate 1: feasible
        assert np.maximum(Ak @ x - ubk, 0).max() < 1e-9 * sc_ and np.maximum(lbk - Ak @ x, 0).max() < 1e-9 * sc_
        # certificate 2: optimal for the UNREGULARISED QP (stationarity with non-negative multipliers on the active rows)
        res = c["stationarity"](x)
        assert res < 1e-8 * max(1.0, np.abs(g).max()), res
        # certificate 3: least norm on the face -- x lies in the row space of the face's equations and working inequality rows
        proj = c["C"].T @ np.linalg.lstsq(c["C"].T, x, rcond=1e-12)[0]
        assert np.abs(proj - x).max() < 1e-8 * sc_
        f = lambda z: 0.5 * z @ H @ z + g @ z
        assert f(x) <= f(x9) + 1e-12 * max(1.0, abs(f(x))) and abs(f(x) - f(x9)) < 1e-9 * max(1.0, abs(f(x)))      # x9 pays O(rho) of objective for a smaller norm
        tn = max(1.0, np.abs(x[28:]).max())
        e9.append(np.abs(x[28:] - x9[28:]).max() / tn); e8.append(np.abs(x[28:] - x8[28:]).max() / tn)
    assert max(e9) < 5e-5 and max(e8) < 5e-4, (e8, e9)
    if mode != 0:       # flight: no contact forces, the optimum does not depend on rho at all
        assert np.mean(e9) < 0.3 * np.mean(e8)        # O(rho) convergence towards the exact optimum


def test_nonuniform_grid_iteration(oracle):
    """Per-interval dt_k in the oracle's SQP iteration (groundwork for SURVEY 8a row S1: OCS2 re-anchors the grid at mode switches).
    A uniform array reproduces the scalar path bit for bit; on a grid whose intervals are split in two the iteration stays close to
    the coarse one at the shared nodes; an interval shortened to land on an event keeps the iteration well defined."""
    from hunter_bipedal_control_b200 import scenarios as sc
    N, dt = 20, 0.02
    x0 = sc.random_initial_states(1, seed=4)[0]
    xr, sw, md, _ = sc.make_reference(x0, (0.2, 0, 0, 0), "trot", N, dt)
    xt, ut = oracle.mpc_cold_start(N, dt, x0, md)
    a = oracle.mpc_iteration(N, dt, x0, xr, sw, md, xt, ut)
    b = oracle.mpc_iteration(N, np.full(N, dt), x0, xr, sw, md, xt, ut)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2]["alpha"] == b[2]["alpha"]
    # refined grid: 2N intervals of dt/2, references sampled on it
    xr2, sw2, md2, _ = sc.make_reference(x0, (0.2, 0, 0, 0), "trot", 2 * N, dt / 2)
    xt2, ut2 = oracle.mpc_cold_start(2 * N, dt / 2, x0, md2)
    c = oracle.mpc_iteration(2 * N, np.full(2 * N, dt / 2), x0, xr2, sw2, md2, xt2, ut2)
    assert c[2]["status"] == 0 and c[2]["alpha"] > 0
    assert np.abs(c[0][::2] - a[0]).max() < 0.05 * max(1.0, np.abs(a[0]).max())     # same problem, finer discretisation
    # non-uniform: the first interval is cut short (a node dropped onto an event at 0.3 dt), the rest re-anchored there
    dts = np.full(N, dt); dts[0] = 0.3 * dt
    times = np.concatenate([[0.0], np.cumsum(dts)])
    xr3 = np.stack([xr[0] + (xr[-1] - xr[0]) * (t / (N * dt)) for t in times])
    d = oracle.mpc_iteration(N, dts, x0, xr3, sw, md, xt, ut)
    assert d[2]["status"] == 0 and np.isfinite(d[0]).all() and d[2]["alpha"] > 0
    # the short interval moves the state less than a full one
    assert np.abs(d[0][1] - x0).max() < np.abs(a[0][1] - x0).max() + 1e-12
