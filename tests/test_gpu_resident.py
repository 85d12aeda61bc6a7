"""Resident closed-loop cycle (hb_resident_cycle_batch): the primal solution stays on the device between solves like
ocs2::SqpSolver's primalSolution_; only t0 / x0 / compact references / rbd go in and torques come out."""
import numpy as np
import pytest

import hunter_bipedal_control_b200 as hb
from hunter_bipedal_control_b200 import scenarios
from oracle import refs as R

pytestmark = pytest.mark.gpu

N, DT = 40, 0.02
MASS = 12.586944


def _setup(B, seed=41):
    x0 = scenarios.random_initial_states(B, seed=seed)
    gaits = [["trot", "standing_trot", "flying_trot", "stance"][i % 4] for i in range(B)]
    compacts = []
    for i in range(B):
        _, _, _, c = scenarios.make_reference(x0[i], (0.3, 0.0, 0.0, 0.1), gaits[i], N, DT, phase=0.03 * i)
        compacts.append(c)
    refs = scenarios.pack_references(compacts, 2 * N * DT)
    rbd = scenarios.consistent_rbd(x0)
    return x0, refs, rbd


def test_cold_cycle_equals_control_step():
    ctx = hb.Context(horizon_N=N, dt=DT, max_batch=512, device=0)
    for B in (7, 300):                      # single chunk and two pipelined half-batches
        x0, refs, rbd = _setup(B)
        t0 = np.zeros(B)
        info, sol, tau, st = ctx.resident_cycle(True, 0.002, t0, x0, refs, rbd)
        xr, sw, md = ctx.reference_expand(t0, refs)
        xt, ut = ctx.mpc_cold_start(x0, md)
        xt2, ut2, info2, sol2, tau2, st2 = ctx.control_step(0.002, x0, xr, sw, md, rbd, xt, ut)
        tr, xres, ures = ctx.resident_read(B)
        assert np.array_equal(tr, t0)
        assert np.array_equal(xres, xt2) and np.array_equal(ures, ut2)          # same kernels, same inputs: bit-identical
        assert np.array_equal(sol, sol2) and np.array_equal(tau, tau2) and np.array_equal(st, st2)
        assert info.tobytes() == info2.tobytes()
    ctx.close()


def test_warm_cycle_shifts_previous_solution():
    ctx = hb.Context(horizon_N=N, dt=DT, max_batch=64, device=0)
    B = 12
    x0, refs, rbd = _setup(B, seed=43)
    t0 = np.zeros(B)
    ctx.resident_cycle(True, 0.002, t0, x0, refs, rbd)
    _, xprev, uprev = ctx.resident_read(B)
    # next solve: grid-aligned shift for even instances, off-grid for odd ones; the measured state moved a little
    t1 = np.where(np.arange(B) % 2 == 0, 2 * DT, 0.013)
    x1 = xprev[:, 0] + 0.3 * (xprev[:, 1] - xprev[:, 0]) + 1e-3
    rbd1 = scenarios.consistent_rbd(x1)
    info, sol, tau, st = ctx.resident_cycle(False, 0.002, t1, x1, refs, rbd1)
    _, xnew, unew = ctx.resident_read(B)
    # expected: warm start by the restated rule, then the ordinary control step on it
    xr, sw, md = ctx.reference_expand(t1, refs)
    xw = np.zeros_like(xprev); uw = np.zeros_like(uprev)
    for i in range(B):
        xw[i], uw[i] = R.warm_start_shift(0.0, t1[i], DT, xprev[i], uprev[i], x1[i], md[i], MASS)
    # the tail of the shifted horizon must come from the initializer, the head from the previous solution
    assert np.all(uw[0, -2:, 12:] == 0.0) and np.allclose(uw[0, 0], uprev[0, 2])
    xt2, ut2, info2, sol2, tau2, st2 = ctx.control_step(0.002, x1, xr, sw, md, rbd1, xw, uw)
    scale = np.abs(tau2).max()
    assert np.abs(xnew - xt2).max() < 1e-9 and np.abs(unew - ut2).max() < 1e-7 * max(1.0, np.abs(ut2).max())
    assert np.abs(tau - tau2).max() < 1e-6 * scale
    assert np.array_equal(st, st2)
    ctx.close()


def test_warm_cycle_requires_previous_solution():
    ctx = hb.Context(horizon_N=N, dt=DT, max_batch=16, device=0)
    x0, refs, rbd = _setup(4)
    with pytest.raises(RuntimeError):
        ctx.resident_cycle(False, 0.002, np.zeros(4), x0, refs, rbd)
    ctx.close()
