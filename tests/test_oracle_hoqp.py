"""HoQP restatement (oracle/hoqp.py, groundwork for SURVEY 8f row N4) against the reference's own unit test
legged_wbc/test/HoQp_test.cpp:19-60 (TEST(HoQP, twoTask)), restated with the same random matrices: Eigen's Random() draws
-1 + 2 rand()/RAND_MAX in column-major order after srand(0)."""
import ctypes

import numpy as np

from oracle.hoqp import HoQp, Task


def eigen_random(libc, rows, cols):
    m = np.zeros((rows, cols))
    for c in range(cols):
        for r in range(rows):
            m[r, c] = -1.0 + 2.0 * libc.rand() / 2147483647.0
    return m


def test_reference_two_task_unit_test():
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(0)
    a0 = eigen_random(libc, 2, 4); d0 = eigen_random(libc, 2, 4)
    task0 = Task(a0, np.ones(2), d0, np.ones(2))
    task1 = Task(np.ones((2, 4)), np.ones(2), d0, np.ones(2))
    h0 = HoQp(task0); h1 = HoQp(task1, h0)
    x0, x1, s0, s1 = h0.solution(), h1.solution(), h0.stacked_slack, h1.stacked_slack
    prec = 1e-6
    if np.allclose(s0, 0, atol=1e-9):
        assert np.allclose(task0.a @ x0, task0.b, rtol=prec, atol=prec)
    if np.allclose(s1, 0, atol=1e-9):
        assert np.allclose(task0.a @ x1, task0.b, rtol=prec, atol=prec)        # the higher priority is untouched ...
        # ... and the lower priority is met as well as the remaining freedom allows (A1 = ones has rank 1, the test only requires this
        # when it is attainable: isApprox on a rank-deficient target is checked in the least-squares sense here)
        Z = h0.stacked_z
        r_best = np.linalg.lstsq(task1.a @ Z, task1.b - task1.a @ x0, rcond=None)[0]
        best = task1.a @ (x0 + Z @ r_best)
        assert np.allclose(task1.a @ x1, best, atol=1e-6)
    assert np.all(task0.d @ x0 <= task0.f + s0 + 1e-8)
    assert np.all(task1.d @ x1 <= task1.f + h1.slack + 1e-8)
    assert np.all(s0 >= -1e-9) and np.all(s1 >= -1e-9)


def test_strict_priority_and_slack_carry_over():
    """Three levels on random data: equalities of a higher level keep their residual, inequalities keep their slack."""
    rng = np.random.default_rng(3)
    n = 8
    t0 = Task(rng.normal(size=(3, n)), rng.normal(size=3), rng.normal(size=(2, n)), rng.normal(size=2) + 1.0)
    t1 = Task(rng.normal(size=(3, n)), rng.normal(size=3), rng.normal(size=(3, n)), rng.normal(size=3))
    t2 = Task(rng.normal(size=(4, n)), rng.normal(size=4))
    h0 = HoQp(t0); h1 = HoQp(t1, h0); h2 = HoQp(t2, h1)
    for h in (h1, h2):
        assert np.allclose(t0.a @ h.solution(), t0.a @ h0.solution(), atol=1e-6)
        assert np.all(t0.d @ h.solution() <= t0.f + h0.slack + 1e-5)
    assert np.allclose(t1.a @ h2.solution(), t1.a @ h1.solution(), atol=1e-6)
    assert np.all(t1.d @ h2.solution() <= t1.f + h1.slack + 1e-5)
    assert h1.stacked_z.shape[1] == n - 6 and h2.stacked_z.shape[1] == 0 and h2.stacked_slack.size == 5
    # the last level uses the two remaining degrees of freedom to reduce its own residual
    assert np.linalg.norm(t2.a @ h2.solution() - t2.b) < np.linalg.norm(t2.a @ h1.solution() - t2.b)


def test_hierarchical_wbc_priorities(oracle):
    """HierarchicalWbc on the oracle's WBC terms: the physics level (EoM, limits, friction, planted stance feet) holds exactly, the base
    task is met as well as the physics allows, and the lowest level cannot change either."""
    from oracle.hoqp import hierarchical_wbc
    from hunter_bipedal_control_b200 import scenarios
    x = scenarios.random_initial_states(3, seed=9)
    for i, mode in enumerate((3, 2, 1)):
        u = np.zeros(22)
        st = [mode in (2, 3), mode in (1, 3), mode in (2, 3), mode in (1, 3)]
        for c in range(4):
            if st[c]:
                u[3 * c + 2] = 12.586944 * 9.81 / sum(st)
        rbd = scenarios.consistent_rbd(x[i:i + 1])[0]
        sol, levels, tasks = hierarchical_wbc(x[i], u, rbd, mode)
        t0, t1, t2 = tasks
        assert np.abs(t0.a @ sol - t0.b).max() < 1e-6                       # EoM, zero swing forces, no contact motion
        assert np.all(t0.d @ sol <= t0.f + 1e-5)                            # torque limits, friction pyramid
        assert np.allclose(t1.a @ sol, t1.a @ levels[1].solution(), atol=1e-5)
        # the weighted WBC solves the same physics with soft priorities: its residual on the base task cannot be smaller
        wsol, wst = oracle.wbc_solve(x[i], u, rbd, mode, False, 1e-8)
        assert wst == 0
        assert np.linalg.norm(t1.a @ sol - t1.b) <= np.linalg.norm(t1.a @ wsol - t1.b) + 1e-4
        assert np.isfinite(sol).all() and np.abs(sol[28:]).max() <= 60 + 1e-6
