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
    ctx = hb.Context(horizon_N=N, dt=DT, max_batch=512, device=0, e2e_chunks=2)
    for B in (7, 300):                      # single chunk and two pipelined half-batches (packed reference upload per chunk)
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


def test_kinematic_closed_loop_rollout():
    """Closed loop without a simulator: the measured state of the next cycle is the planned state one node ahead (perfect tracking).
    Exercises planner state, device planner, warm-start shift and the solve over many consecutive cycles (towards SURVEY 8f row N2)."""
    ctx = hb.Context(horizon_N=N, dt=DT, max_batch=16, device=0)
    B, cycles = 8, 40
    x = scenarios.random_initial_states(B, seed=71)
    x[:, 0:6] = 0.0
    cmd = np.tile([0.3, 0.0, 0.0, 0.0], (B, 1)); cmd[B // 2:, 0] = -0.2
    gaits = ["trot"] * (B // 2) + ["standing_trot"] * (B - B // 2)
    x_start = x.copy()
    t = 0.0
    worst_tau = 0.0
    for c in range(cycles):
        rbd = scenarios.consistent_rbd(x)
        ins = hb.make_plan_inputs(np.full(B, t), N * DT, x, cmd, None, gaits, 0.2)
        info, sol, tau, st, ps = ctx.resident_plan_cycle(c == 0, 0.002, ins, rbd)
        assert (ps == 0).all(), (c, ps)
        assert (info["status"] == 0).all(), (c, info["status"])
        assert (st == 0).all(), (c, st)
        worst_tau = max(worst_tau, np.abs(tau).max())
        _, xt, ut = ctx.resident_read(B)
        x = xt[:, 1].copy()                      # perfect tracking of the plan over one node
        t += DT
    T_total = cycles * DT
    assert np.isfinite(x).all() and worst_tau <= 60.0 + 1e-6
    # the robots kept their height and moved in the commanded direction (the gait starts after 0.2 s of stance)
    assert np.all(np.abs(x[:, 8] - 0.63) < 0.05)
    yaw = x_start[:, 9]
    fwd = (x[:, 6] - x_start[:, 6]) * np.cos(yaw) + (x[:, 7] - x_start[:, 7]) * np.sin(yaw)
    assert np.all(fwd[:B // 2] > 0.05) and np.all(fwd[B // 2:] < -0.03), fwd
    assert np.all(np.abs(fwd) < 1.5 * np.abs(cmd[:, 0]) * T_total + 0.05)
    ctx.close()


def test_wbc_fallback_returns_previous_solution():
    """W5 (WeightedWbc.cpp:57-64): an instance whose QP does not solve keeps the solution of its previous cycle."""
    ctx = hb.Context(horizon_N=N, dt=DT, max_batch=16, device=0)
    B = 6
    x0, refs, rbd = _setup(B, seed=47)
    t0 = np.zeros(B)
    info, sol0, tau0, st0 = ctx.resident_cycle(True, 0.002, t0, x0, refs, rbd)
    assert (st0 == 0).all()
    bad = rbd.copy()
    bad[2, 22:32] = 1e7          # absurd joint velocities: the torque-limit rows cannot be met
    bad[4, 6:16] = np.nan        # NaN measurement
    info, sol1, tau1, st1 = ctx.resident_cycle(False, 0.002, t0 + DT, x0, refs, bad)
    assert st1[2] != 0 and st1[4] != 0 and (np.delete(st1, [2, 4]) == 0).all()
    for i in (2, 4):
        assert np.array_equal(sol1[i], sol0[i]) and np.array_equal(tau1[i], tau0[i])
    for i in (0, 1, 3, 5):
        assert not np.array_equal(sol1[i], sol0[i])
    ctx.close()


def test_snapshot_write_back_resumes_bit_identically():
    """Checkpoint / resume: a resident solution read out and written into a FRESH context continues exactly like the original one --
    uniform grid and event-node grid."""
    for event_nodes in (False, True):
        kw = dict(horizon_N=N, dt=DT, max_batch=32, device=0)
        if event_nodes:
            kw.update(time_horizon=0.7, event_nodes=True)
        a = hb.Context(**kw); b = hb.Context(**kw)
        B = 10
        x0, refs, rbd = _setup(B, seed=47)
        a.resident_cycle(True, 0.002, np.zeros(B), x0, refs, rbd)
        a.resident_cycle(False, 0.002, np.full(B, 0.01), x0 + 1e-3, refs, rbd)
        t_s, x_s, u_s = a.resident_read(B)
        grid = a.resident_read_grid(B) if event_nodes else (None, None)
        xr, sw, md = (a.reference_expand_grid(grid[0], refs) if event_nodes else a.reference_expand(t_s, refs))
        ra = a.resident_cycle(False, 0.002, np.full(B, 0.02), x0 + 2e-3, refs, rbd)
        b.resident_write(t_s, x_s, u_s, md, *grid)
        rb = b.resident_cycle(False, 0.002, np.full(B, 0.02), x0 + 2e-3, refs, rbd)
        assert ra[0].tobytes() == rb[0].tobytes() and np.array_equal(ra[1], rb[1]) and np.array_equal(ra[2], rb[2])
        xa, xb = a.resident_read(B), b.resident_read(B)
        assert np.array_equal(xa[1], xb[1]) and np.array_equal(xa[2], xb[2])
        # the 500 Hz tick works on the restored snapshot too (needs the node modes)
        c = hb.Context(**kw)
        c.resident_write(t_s, x_s, u_s, md, *grid)
        ta = a.resident_wbc(np.full(B, 0.024), rbd); tb = b.resident_wbc(np.full(B, 0.024), rbd)
        assert np.array_equal(ta[3], tb[3])
        with pytest.raises(hb.HunterB200Error):
            hb.Context(**kw).resident_cycle(False, 0.002, np.zeros(B), x0, refs, rbd)      # nothing to shift in a fresh context
        for cx in (a, b, c):
            cx.close()


def test_pinned_reference_array_is_gathered_by_the_device():
    """A page-locked hb_reference array is read by the device directly (zero-copy gather of the used entries, no host packing pass);
    a pageable one goes through the packed staging upload. Same results bit for bit, one chunk and two chunks."""
    import ctypes as C
    import torch
    ctx = hb.Context(horizon_N=N, dt=DT, max_batch=512, device=0, e2e_chunks=2)
    for B in (9, 300):
        x0, refs, rbd = _setup(B, seed=77)
        t0 = np.full(B, 0.004)
        a = ctx.resident_cycle(True, 0.002, t0, x0, refs, rbd)                        # ctypes array: pageable
        up_pageable = ctx.last_reference_upload_bytes
        _, xa, ua = ctx.resident_read(B)
        nbytes = C.sizeof(refs)
        pinned = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        C.memmove(pinned.data_ptr(), C.addressof(refs), nbytes)
        b = ctx.resident_cycle(True, 0.002, t0, x0, C.cast(C.c_void_p(pinned.data_ptr()), C.POINTER(hb.HbReference)), rbd)
        up_pinned = ctx.last_reference_upload_bytes
        _, xb, ub = ctx.resident_read(B)
        assert np.array_equal(xa, xb) and np.array_equal(ua, ub)
        for p, q in zip(a, b):
            assert np.array_equal(p, q)
        assert 0 < up_pinned <= up_pageable < nbytes                                   # only used entries cross PCIe either way
    ctx.close()


def test_malformed_references_are_rejected_on_both_upload_paths():
    """hb_reference structs with counts beyond their capacity, unordered times or an empty target list are refused (-1) whether the host
    validates them (pageable array) or the device does while it gathers them (pinned array); nothing is indexed out of bounds."""
    import ctypes as C
    import torch
    B = 12
    ctx = hb.Context(horizon_N=N, dt=DT, max_batch=64, device=0)
    x0, refs, rbd = _setup(B, seed=5)
    t0 = np.zeros(B)

    def corrupt(kind):
        bad = (hb.HbReference * B)()
        C.memmove(C.addressof(bad), C.addressof(refs), C.sizeof(refs))
        if kind == "events":
            bad[3].n_events = 4000
        elif kind == "targets":
            bad[7].n_targets = 0
        elif kind == "segments":
            bad[5].n_segments[1][2] = -3
        elif kind == "order":
            bad[2].target_times[1] = bad[2].target_times[0] - 1.0
        elif kind == "mode":
            bad[9].modes[0] = 7
        return bad

    for kind in ("events", "targets", "segments", "order", "mode"):
        bad = corrupt(kind)
        with pytest.raises(hb.HunterB200Error, match="invalid argument"):
            ctx.resident_cycle(True, 0.002, t0, x0, bad, rbd)
        pinned = torch.empty(C.sizeof(bad), dtype=torch.uint8).pin_memory()
        C.memmove(pinned.data_ptr(), C.addressof(bad), C.sizeof(bad))
        with pytest.raises(hb.HunterB200Error, match="invalid argument"):
            ctx.resident_cycle(True, 0.002, t0, x0, C.cast(C.c_void_p(pinned.data_ptr()), C.POINTER(hb.HbReference)), rbd)
    # the context is still usable afterwards
    info, sol, tau, st = ctx.resident_cycle(True, 0.002, t0, x0, refs, rbd)
    assert (info["status"] == 0).all() and np.isfinite(tau).all()
    ctx.close()
