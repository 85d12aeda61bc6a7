"""Committed golden vectors (tests/golden/planner_golden.npz, generator gen_planner_golden.py): the host planner and the
restatement on the CPU, the device planner + expansion and the Kalman filter on the GPU."""
import os

import numpy as np
import pytest

import hunter_bipedal_control_b200 as hb
from oracle import refs as R

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "planner_golden.npz"))
N, DT = int(G["N"]), float(G["DT"])
T = N * DT


def test_host_planner_against_golden():
    n = G["x0"].shape[0]
    gaits = [str(g) for g in G["gaits"]]
    refs, ls = hb.plan_references(G["t0"], T, G["x0"], G["cmd"], G["feet"], gaits, G["start"], latest_stance=G["latest"])
    np.testing.assert_allclose(ls, G["latest_out"], rtol=0, atol=1e-14)
    for i in range(n):
        times = G["t0"][i] + DT * np.arange(N + 1)
        ok = ~G["on_event"][i]
        xc, sc, mc = R.eval_compact(refs[i], times)
        np.testing.assert_array_equal(mc[ok], G["mode"][i][ok])
        np.testing.assert_allclose(xc, G["x_ref"][i], rtol=0, atol=1e-9)
        np.testing.assert_allclose(sc[ok], G["swing"][i][ok], rtol=0, atol=1e-11)


def test_restatement_against_golden():
    """The restatement regenerates its own golden bit for bit (guards the fixture against silent edits of oracle/refs.py)."""
    i = 2
    ms, tg, sp = R.plan(G["t0"][i], T, G["x0"][i], G["cmd"][i], G["feet"][i], str(G["gaits"][i]), G["start"][i], latest_stance=G["latest"][i])
    xr, sw, md = R.sample(ms, tg, sp, G["t0"][i] + DT * np.arange(N + 1))
    assert np.array_equal(xr, G["x_ref"][i]) and np.array_equal(sw, G["swing"][i]) and np.array_equal(md, G["mode"][i])


@pytest.mark.gpu
def test_device_planner_and_expansion_against_golden():
    ctx = hb.Context(horizon_N=N, dt=DT, max_batch=16, device=0)
    n = G["x0"].shape[0]
    gaits = [str(g) for g in G["gaits"]]
    ins = hb.make_plan_inputs(G["t0"], T, G["x0"], G["cmd"], G["feet"], gaits, G["start"])
    refs, ls, st = ctx.plan_references_gpu(ins, G["latest"])
    assert (st == 0).all()
    xr, sw, md = ctx.reference_expand(G["t0"], refs)
    for i in range(n):
        ok = ~G["on_event"][i]
        np.testing.assert_array_equal(md[i][ok], G["mode"][i][ok])
        np.testing.assert_allclose(xr[i], G["x_ref"][i], rtol=0, atol=1e-8)
        np.testing.assert_allclose(sw[i][ok], G["swing"][i][ok], rtol=0, atol=1e-10)
    ctx.close()


@pytest.mark.gpu
def test_kalman_filter_against_golden():
    ctx = hb.Context(horizon_N=N, dt=DT, max_batch=16, device=0)
    steps, B = G["kf_quat"].shape[:2]
    st = hb.kf_states(B)
    for k in range(steps):
        rbd = ctx.estimator_update(float(G["kf_dt"]), st, G["kf_quat"][k], G["kf_wl"][k], G["kf_al"][k], G["kf_jpos"][k], G["kf_jvel"][k], G["kf_flags"][k])
        np.testing.assert_allclose(rbd, G["kf_rbd"][k], rtol=0, atol=1e-9)
        for i in range(B):
            np.testing.assert_allclose(np.array(st[i].x_hat[:]), G["kf_x"][k, i], rtol=0, atol=1e-9)
            P = np.array(st[i].P[:]).reshape(18, 18)
            assert np.abs(P - G["kf_P"][k, i]).max() < 1e-9 * max(1.0, np.abs(G["kf_P"][k, i]).max())
    ctx.close()
