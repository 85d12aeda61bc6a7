"""Property tests of the oracle over random states (hypothesis): structural facts that hold for every configuration."""
import numpy as np
from hypothesis import given, settings, strategies as st

X0 = np.array([0, 0, 0, 0, 0, 0, 0, 0, 0.63, 0, 0, 0, .1, 0, .4, .93, .53, -.1, 0, -.4, .93, -.53])
LO = np.array([-0.2, -0.5, -0.8, 0.0, -1.1, -0.5, -1.0, -1.2, 0.0, -1.1])
HI = np.array([0.5, 1.0, 1.2, 1.5, 1.1, 0.2, 0.5, 0.8, 1.5, 1.1])


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 10**6))
def test_rigid_body_structure(oracle, seed):
    rng = np.random.default_rng(seed)
    q = np.r_[rng.uniform(-1, 1, 3), rng.uniform(-np.pi, np.pi), rng.uniform(-0.6, 0.6, 2), rng.uniform(LO, HI)]
    v = rng.uniform(-2, 2, 16)
    r = oracle.rbd(q, v)
    M = r["M"]
    assert np.abs(M - M.T).max() < 1e-11 and np.linalg.eigvalsh(M).min() > 1e-6
    # translation block of M is m I; linear momentum rows of A equal the translation rows of M
    assert np.allclose(M[:3, :3], 12.586944 * np.eye(3), atol=1e-10)
    assert np.allclose(r["A"][:3], M[:3], atol=1e-10)
    assert np.abs(r["A"] @ v - r["h"]).max() < 1e-11
    # translating the base moves contacts and CoM rigidly and leaves M, A unchanged
    d = rng.uniform(-1, 1, 3)
    q2 = q.copy(); q2[:3] += d
    r2 = oracle.rbd(q2, v)
    assert np.abs(r2["M"] - M).max() < 1e-11 and np.abs(r2["A"] - r["A"]).max() < 1e-11
    assert np.abs(r2["cpos"].reshape(4, 3) - r["cpos"].reshape(4, 3) - d).max() < 1e-12
    # kinetic energy identity 1/2 v'Mv >= 1/2 |h_lin|^2 / m (Koenig)
    assert 0.5 * v @ M @ v >= 0.5 * (r["h"][:3] @ r["h"][:3]) / 12.586944 - 1e-9


@settings(max_examples=10, deadline=None)
@given(st.integers(0, 10**6), st.sampled_from([0, 1, 2, 3]))
def test_wbc_solution_satisfies_reference_constraints(oracle, seed, mode):
    rng = np.random.default_rng(seed)
    x = X0 + rng.uniform(-.05, .05, 22)
    u = np.zeros(22)
    fl = [mode in (2, 3), mode in (1, 3), mode in (2, 3), mode in (1, 3)]
    for c in range(4):
        if fl[c]:
            u[3 * c + 2] = 12.586944 * 9.81 / sum(fl)
    q = x[6:] + rng.uniform(-.02, .02, 16)
    rbd = np.r_[q[3:6], q[0:3], q[6:], rng.uniform(-.3, .3, 16)]
    H, g, A, lb, ub = oracle.wbc_assemble(x, u, rbd, mode, False)
    sol, stt, _ = oracle.qp_solve(H, g, A, lb, ub, 1e-8)
    assert stt == 0
    Ax = A @ sol
    assert (Ax <= ub + 1e-7 * (1 + np.abs(ub))).all() and (Ax >= lb - 1e-7 * (1 + np.abs(lb))).all()
    for c in range(4):
        if not fl[c]:
            assert np.abs(sol[16 + 3 * c:19 + 3 * c]).max() < 1e-8
