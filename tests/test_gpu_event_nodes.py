"""SURVEY 8a row S1 on the device: time discretisation with event nodes at the SHIPPED solver settings (sqp.dt 0.015, mpc.timeHorizon 0.8,
legged_controllers/config/hunter/task.info:82,144) with the gait events off the 15 ms grid. The CUDA path (node times, expansion on the
grid, SQP iteration with per-interval dt_k, warm-start shift between two non-uniform grids, policy evaluation) against the oracle's
per-interval iteration and plain numpy restatements."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
DT, T, CAP = 0.015, 0.8, 64
TAU_RTOL = 1e-4


def rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


@pytest.fixture(scope="module")
def ev_ctx():
    import hunter_bipedal_control_b200 as hb
    ctx = hb.Context(horizon_N=CAP, dt=DT, max_batch=32, device=0, time_horizon=T, event_nodes=True)
    yield ctx
    ctx.close()


def _cases(B, seed, t0=0.0):
    from hunter_bipedal_control_b200 import scenarios as sc
    rng = np.random.default_rng(seed)
    gaits = ["trot", "standing_trot", "flying_trot", "trot"]
    x0 = sc.random_initial_states(B, seed=100 + seed)
    compacts = []
    for i in range(B):
        # the schedule is anchored at time 0 whatever t0 is; phase puts the switches off the 15 ms grid
        _, _, _, c = sc.make_reference(x0[i], (0.3, 0.0, 0.0, 0.1), gaits[i % 4], 54, DT, t0=0.0, phase=float(rng.uniform(0.011, 0.29)))
        compacts.append(c)
    return x0, compacts, sc.pack_references(compacts, t0 + 2.0)


def test_time_grid_and_expansion_on_the_grid(ev_ctx):
    from hunter_bipedal_control_b200 import scenarios as sc
    B = 12
    x0, compacts, refs = _cases(B, 1)
    t0 = np.linspace(0.0, 0.11, B)          # some instances start after their first switch
    tk, nn, st = ev_ctx.time_grid(t0, refs)
    assert (st == 0).all()
    xr, sw, md = ev_ctx.reference_expand_grid(tk, refs)
    n_event_nodes = 0
    for i in range(B):
        g = sc.event_time_grid(t0[i], T, DT, compacts[i]["events"], CAP)
        assert nn[i] == len(g) - 1
        assert np.abs(tk[i, :nn[i] + 1] - g).max() < 1e-12
        assert abs(tk[i, nn[i]] - (t0[i] + T)) < 1e-12 and (np.diff(tk[i, :nn[i] + 1]) > 1e-9).all() and (np.diff(tk[i, :nn[i] + 1]) < DT + 1e-12).all()
        inside = [e for e in compacts[i]["events"] if t0[i] + 1e-9 < e < t0[i] + T - 1e-9]
        for e in inside:                     # every switch inside the horizon is a node
            assert np.abs(tk[i, :nn[i] + 1] - e).min() < 1e-12
        n_event_nodes += len(inside)
        xr_o, sw_o, md_o = sc.sample_reference(compacts[i], g)
        assert np.array_equal(md[i, :nn[i] + 1], md_o)
        assert np.abs(xr[i, :nn[i] + 1] - xr_o).max() < 1e-12 and np.abs(sw[i, :nn[i] + 1] - sw_o).max() < 1e-10
    assert n_event_nodes >= 2 * B            # the test does exercise off-grid switches
    assert len(set(nn.tolist())) > 1         # and instances with different interval counts in one batch


def test_sqp_iteration_on_event_grids_vs_oracle(ev_ctx, oracle):
    """One and two SQP iterations on per-instance non-uniform grids: alpha, trial counts and trajectories equal the oracle's (per-interval dt_k)."""
    B = 8
    x0, compacts, refs = _cases(B, 2)
    t0 = np.full(B, 0.004)
    tk, nn, st = ev_ctx.time_grid(t0, refs)
    xr, sw, md = ev_ctx.reference_expand_grid(tk, refs)
    xt, ut = ev_ctx.mpc_cold_start(x0, md)
    a1 = ev_ctx.mpc_solve_grid(x0, tk, nn, xr, sw, md, xt, ut)
    a2 = ev_ctx.mpc_solve_grid(x0, tk, nn, xr, sw, md, a1[0], a1[1])
    assert (a1[2]["status"] == 0).all() and (a2[2]["status"] == 0).all()
    for i in range(B):
        n = int(nn[i])
        dts = np.diff(tk[i, :n + 1])
        xo, uo = xt[i, :n + 1], ut[i, :n]
        for it, dev in enumerate((a1, a2)):
            xo, uo, io = oracle.mpc_iteration(n, dts, x0[i], xr[i, :n + 1], sw[i, :n + 1], md[i, :n + 1], xo, uo)
            assert io["alpha"] == dev[2]["alpha"][i] and io["n_trials"] == dev[2]["n_trials"][i], (i, it, io, dev[2][i])
            assert abs(io["merit0"] - dev[2]["merit0"][i]) < 1e-8 * max(1.0, abs(io["merit0"]))
            assert np.abs(xo - dev[0][i, :n + 1]).max() < 1e-7 * max(1.0, np.abs(xo).max()), (i, it)
            assert np.abs(uo - dev[1][i, :n]).max() < 1e-6 * max(1.0, np.abs(uo).max()), (i, it)


def _shift_numpy(tk_prev, n_prev, xp, up, tk_new, n_new, x0, mode_new, mass=12.586944):
    """SqpSolver::initializeStateInputTrajectories between two arbitrary grids (restated): previous solution interpolated where the new
    node lies inside the previous horizon, the initializer (state kept, weight-compensating input) afterwards."""
    from hunter_bipedal_control_b200 import scenarios as sc
    tp = tk_prev[:n_prev + 1]
    t_end = tp[-1]

    def locate(t):
        k = int(np.clip(np.searchsorted(tp, t, side="right") - 1, 0, n_prev - 1))
        al = float(np.clip((t - tp[k]) / (tp[k + 1] - tp[k]), 0.0, 1.0))
        return k, al
    istar = n_new
    for i in range(n_new):
        if tk_new[i + 1] > t_end + 1e-9:
            istar = i
            break
    x = np.zeros((n_new + 1, 22)); u = np.zeros((n_new, 22))
    for k in range(n_new + 1):
        ks = min(k, istar)
        if ks == 0:
            x[k] = x0
        else:
            j, al = locate(tk_new[ks])
            x[k] = (1 - al) * xp[j] + al * xp[j + 1]
    for k in range(n_new):
        if k < istar:
            j, al = locate(tk_new[k])
            j1 = min(j + 1, n_prev - 1)
            u[k] = (1 - al) * up[j] + al * up[j1]
        else:
            fl = sc.mode_flags(int(mode_new[k]))
            for c in range(4):
                if fl[c]:
                    u[k, 3 * c + 2] = mass * 9.81 / sum(fl)
    return x, u


def test_resident_cycles_with_event_nodes_vs_oracle(ev_ctx, oracle):
    """Closed loop at the shipped rates: cold cycle at t0, warm cycles 10 ms apart (MPC 100 Hz); every cycle re-discretises the horizon, so the
    warm start interpolates between two different non-uniform grids. Resident trajectories and torques against the restated pipeline."""
    from hunter_bipedal_control_b200 import scenarios as sc
    B = 6
    x0, compacts, refs = _cases(B, 3, t0=0.05)
    rbd = sc.consistent_rbd(x0, np.random.default_rng(7), 0.0)
    prev = None
    for cyc, t_now in enumerate((0.0, 0.01, 0.02, 0.03)):
        t0 = np.full(B, t_now)
        info, sol, tau, st = ev_ctx.resident_cycle(cyc == 0, 0.002, t0, x0, refs, rbd)
        assert (info["status"] == 0).all() and (st == 0).all()
        _, xt_d, ut_d = ev_ctx.resident_read(B)
        tk_d, nn_d = ev_ctx.resident_read_grid(B)
        cur = []
        for i in range(B):
            g = sc.event_time_grid(t_now, T, DT, compacts[i]["events"], CAP)
            n = len(g) - 1
            assert nn_d[i] == n and np.abs(tk_d[i, :n + 1] - g).max() < 1e-12
            xr, sw, md = sc.sample_reference(compacts[i], g)
            if cyc == 0:
                xs, us = oracle.mpc_cold_start(n, DT, x0[i], md)
            else:
                gp, xp, up = prev[i]
                xs, us = _shift_numpy(gp, len(gp) - 1, xp, up, g, n, x0[i], md)
            xo, uo, io = oracle.mpc_iteration(n, np.diff(g), x0[i], xr, sw, md, xs, us)
            assert io["alpha"] == info["alpha"][i], (cyc, i, io, info[i])
            assert np.abs(xo - xt_d[i, :n + 1]).max() < 1e-6 * max(1.0, np.abs(xo).max()), (cyc, i)
            assert np.abs(uo - ut_d[i, :n]).max() < 1e-5 * max(1.0, np.abs(uo).max()), (cyc, i)
            # policy at t0 + 2 ms on the grid, then the WBC
            k = int(np.clip(np.searchsorted(g, t_now + 0.002, side="right") - 1, 0, n - 1))
            al = (t_now + 0.002 - g[k]) / (g[k + 1] - g[k])
            xd = (1 - al) * xo[k] + al * xo[k + 1]; ud = (1 - al) * uo[k] + al * uo[min(k + 1, n - 1)]
            so, sto = oracle.wbc_solve(xd, ud, rbd[i], int(md[k]), False, 1e-8)
            assert sto == 0 and rel(tau[i], so[28:]) < TAU_RTOL, (cyc, i)
            cur.append((g, xo, uo))
        prev = cur
