"""BASELINE.json configs 3-5 at (near) full size on the GPU: size-independent properties plus oracle parity on samples."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TAU_RTOL = 1e-4


def rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max())


@pytest.fixture(scope="module")
def big_ctx():
    import hunter_bipedal_control_b200 as hb
    ctx = hb.Context(horizon_N=100, dt=0.01, max_batch=2048, device=0)
    yield ctx
    ctx.close()


def test_config3_cmd_vel_grid(big_ctx, oracle):
    """Config 3 shape: walking mode (WBC non-stance tasks), trot schedule, cmd_vel grid vx x wz in [-0.5, 0.5]^2 (32 x 32) x 2 pose seeds."""
    from hunter_bipedal_control_b200 import scenarios as sc
    N, dt = 100, 0.01
    vx = np.linspace(-0.5, 0.5, 32); wz = np.linspace(-0.5, 0.5, 32)
    cmds = [(a, 0.0, 0.0, b) for a in vx for b in wz] * 2
    B = len(cmds)
    x0 = np.concatenate([sc.random_initial_states(1024, seed=0), sc.random_initial_states(1024, seed=1)])
    x_ref = np.zeros((B, N + 1, 22)); swing = np.zeros((B, N + 1, 24)); mode = np.zeros((B, N + 1), dtype=np.int32)
    for i in range(B):
        x_ref[i], swing[i], mode[i], _ = sc.make_reference(x0[i], cmds[i], "trot", N, dt)
    rbd = sc.consistent_rbd(x0, np.random.default_rng(5), 0.0)
    xt, ut = big_ctx.mpc_cold_start(x0, mode)
    xt1, ut1, info, sol, tau, st = big_ctx.control_step(0.002, x0, x_ref, swing, mode, rbd, xt, ut)
    assert (info["status"] == 0).all() and (st == 0).all()
    assert (info["alpha"] > 0).mean() > 0.99
    assert np.isfinite(tau).all() and (np.abs(tau) <= np.tile([28, 60, 60, 60, 28], 2) + 1e-6).all()
    # friction pyramid and unilateral contact on every instance
    F = sol[:, 16:28].reshape(B, 4, 3)
    assert (F[:, :, 2] > -1e-6).all()
    assert (np.abs(F[:, :, 0]) <= 0.7 * F[:, :, 2] + 1e-5).all() and (np.abs(F[:, :, 1]) <= 0.7 * F[:, :, 2] + 1e-5).all()
    # defects shrink after the step wherever the filter accepted it on violation
    assert np.median(info["viol1"]) < np.median(info["viol0"])
    for i in (0, 517, 1023, 1024, 2047):
        xo, uo, io = oracle.mpc_iteration(N, dt, x0[i], x_ref[i], swing[i], mode[i], xt[i], ut[i])
        assert io["alpha"] == info["alpha"][i]
        assert np.abs(xo - xt1[i]).max() < 1e-7 * max(1, np.abs(xo).max())
        al = 0.002 / dt
        so, sto = oracle.wbc_solve((1 - al) * xo[0] + al * xo[1], (1 - al) * uo[0] + al * uo[1], rbd[i], int(mode[i][0]), False, 1e-8)
        assert sto == 0 and rel(tau[i], so[28:]) < TAU_RTOL


def test_config4_mixed_schedules_sorted_permutation(big_ctx):
    """Config 4 shape on one GPU: schedule of instance i = i mod 4, random phase; sorting instances by schedule is a pure permutation."""
    from hunter_bipedal_control_b200 import scenarios as sc, sharding
    N, dt, B = 100, 0.01, 512
    gaits = [["stance", "trot", "standing_trot", "flying_trot"][i % 4] for i in range(B)]
    rng = np.random.default_rng(4)
    x0 = sc.random_initial_states(B, seed=44)
    x_ref = np.zeros((B, N + 1, 22)); swing = np.zeros((B, N + 1, 24)); mode = np.zeros((B, N + 1), dtype=np.int32)
    for i in range(B):
        x_ref[i], swing[i], mode[i], _ = sc.make_reference(x0[i], (0.2, 0, 0, 0), gaits[i], N, dt, phase=rng.uniform(0, 0.3))
    perm, inv = sharding.sort_by_schedule(mode)
    xt, ut = big_ctx.mpc_cold_start(x0, mode)
    a = big_ctx.mpc_solve(x0, x_ref, swing, mode, xt, ut)
    b = big_ctx.mpc_solve(x0[perm], x_ref[perm], swing[perm], mode[perm], xt[perm], ut[perm])
    assert np.array_equal(b[0][inv], a[0]) and np.array_equal(b[1][inv], a[1])
    assert (a[2]["status"] == 0).all()
    assert set(np.unique(mode)) == {0, 1, 2, 3}


def test_config5_raw_qp_sweep(big_ctx, oracle):
    """Config 5: raw WeightedWbc QPs in qpOASES layout, modes {STANCE 50 %, L 25 %, R 25 %}; B = 2048 on the device, oracle parity on a sample."""
    from hunter_bipedal_control_b200 import scenarios as sc
    rng = np.random.default_rng(8)
    nb = 64
    H0 = np.zeros((nb, 38, 38)); g0 = np.zeros((nb, 38)); A0 = np.zeros((nb, 60, 38)); lb0 = np.full((nb, 60), -1e20); ub0 = np.full((nb, 60), 1e20)
    for i in range(nb):
        md = int(rng.choice([3, 3, 2, 1]))
        x = sc.INITIAL_STATE + rng.uniform(-.05, .05, 22)
        u = np.zeros(22)
        fl = sc.mode_flags(md)
        for c in range(4):
            if fl[c]:
                u[3 * c + 2] = sc.TOTAL_MASS * 9.81 / sum(fl)
        u[12:] = rng.uniform(-.5, .5, 10)
        rbd = sc.consistent_rbd(x[None], rng, 0.02)[0]
        Hi, gi, Ai, lbi, ubi = oracle.wbc_assemble(x, u, rbd, md, False)
        m = Ai.shape[0]
        H0[i] = Hi; g0[i] = gi; A0[i, :m] = Ai; lb0[i, :m] = lbi; ub0[i, :m] = ubi
    reps = 2048 // nb
    H = np.tile(H0, (reps, 1, 1)); g = np.tile(g0, (reps, 1)); A = np.tile(A0, (reps, 1, 1)); lb = np.tile(lb0, (reps, 1)); ub = np.tile(ub0, (reps, 1))
    g = g * (1.0 + 0.01 * rng.uniform(-1, 1, g.shape))       # distinct problems
    x, st, it = big_ctx.wbc_qp(H, g, A, lb, ub)
    assert (st == 0).all(), (np.unique(st, return_counts=True), it.max())
    assert it.max() < 40
    Ax = np.einsum("bij,bj->bi", A, x)
    assert (Ax <= ub + 1e-6 * (1 + np.abs(ub))).all() and (Ax >= lb - 1e-6 * (1 + np.abs(lb))).all()
    xo, sto = oracle.wbc_qp_batch(H[:16], g[:16], A[:16], lb[:16], ub[:16], 1e-8, threads=4)
    for i in range(16):
        assert sto[i] == 0 and rel(x[i, 28:], xo[i, 28:]) < TAU_RTOL


def test_fixed_point_idempotence(big_ctx):
    """At an SQP fixed point a further iteration leaves the trajectories unchanged (to solver tolerance)."""
    from hunter_bipedal_control_b200 import scenarios as sc
    N, dt, B = 100, 0.01, 64
    x0, x_ref, swing, mode = sc.make_batch(B, N, dt, gait="stance", seed=71)
    xt, ut = big_ctx.mpc_cold_start(x0, mode)
    for _ in range(25):
        xt, ut, info = big_ctx.mpc_solve(x0, x_ref, swing, mode, xt, ut)
    xt2, ut2, info2 = big_ctx.mpc_solve(x0, x_ref, swing, mode, xt, ut)
    assert (info2["status"] == 0).all()
    dxm = np.abs(xt2 - xt).reshape(B, -1).max(axis=1)
    # Gauss-Newton converges linearly on some instances: most are at their fixed point, none moves far, rejected steps change nothing
    assert np.median(dxm) < 1e-6 and dxm.max() < 1e-2
    rej = info2["alpha"] == 0
    assert rej.sum() > 0 and (dxm[rej] == 0).all()
    # the violation floor is the least-squares residual of the toe/heel contact rows, it must not grow
    assert (info2["viol1"] <= info2["viol0"] * (1 + 1e-9) + 1e-12).all() and np.median(info2["viol1"]) < 1e-2
