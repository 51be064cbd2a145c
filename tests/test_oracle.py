"""CPU tests of the oracle itself: golden vector, internal consistency, gradient checks.

The oracle (numpy fp32 forward, plain-C forward+adjoint, torch-autograd gradient) is the
checker for the CUDA path; these tests pin it to the only numbers the reference recorded and
cross-check its three restatements against each other.
"""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from doper_b200 import synthetic as syn
from oracle import ball_sim_np as onp
from oracle import ball_sim_torch as ot
from oracle.notebook_cell7 import run_notebook_sim
from tests.helpers import assert_bitexact, grad_rel_err, same_f32

GOLDEN = Path(__file__).parent / "golden"


# ---- golden: reference notebook cell 7 ---------------------------------------------------------
def _golden():
    return json.loads((GOLDEN / "notebook_cell7.json").read_text())


def test_notebook_cell7_bitexact():
    """All 126 float32 values the reference recorded are reproduced exactly."""
    g = _golden()
    x, v, traj = run_notebook_sim(g["physics"])
    assert np.array_equal(x, np.asarray(g["final_coordinate"], np.float32))
    assert np.array_equal(v, np.asarray(g["final_velocity"], np.float32))
    assert np.array_equal(traj, np.asarray(g["trajectory"], np.float32))


def test_notebook_cell7_needs_contracted_position_update(monkeypatch):
    """Documents WHY get_new_state's coordinate update is a single-rounding FMA in the oracle:
    evaluated 'as written' (two roundings) 36 recorded values are off by exactly one ulp."""
    g = _golden()
    monkeypatch.setattr(onp, "CONTRACT_POSITION_UPDATE", False)
    x, v, traj = run_notebook_sim(g["physics"])
    gold = [np.asarray(g[k], np.float32) for k in ("final_coordinate", "final_velocity", "trajectory")]
    ulps = [np.abs(a.view(np.int32) - b.view(np.int32)) for a, b in zip((x, v, traj), gold)]
    assert max(int(u.max()) for u in ulps) == 1
    assert sum(int((u != 0).sum()) for u in ulps) == 36


def test_fma_f32_against_exact_rational():
    from fractions import Fraction
    rng = np.random.default_rng(0)
    a = rng.normal(size=2000).astype(np.float32)
    b = rng.normal(size=2000).astype(np.float32)
    c = (-(a.astype(np.float64) * b.astype(np.float64)) * (1 + rng.normal(size=2000) * 1e-7)).astype(np.float32)
    got = onp.fma_f32(a, b, c)
    for i in range(2000):
        exact = Fraction(float(a[i])) * Fraction(float(b[i])) + Fraction(float(c[i]))
        lo = np.float32(float(exact))  # python rounds Fraction->float64 correctly; then to f32
        # correct rounding of the exact rational to binary32 via integer arithmetic
        cand = [np.nextafter(lo, np.float32(-np.inf)), lo, np.nextafter(lo, np.float32(np.inf))]
        best = min(cand, key=lambda z: (abs(Fraction(float(z)) - exact), int(np.float32(z).view(np.uint32)) & 1))
        assert got[i] == best


# ---- three restatements agree ------------------------------------------------------------------
@pytest.fixture(scope="module")
def small():
    w = syn.make_workload("c2", n_balls=192)
    w["c"] = onp.make_constants(w["constants"])
    return w


def test_numpy_vs_c_forward_bitexact(small, cref):
    w, n = small, 300
    r = onp.run_sim(n / 1000, n, w["segments"], w["x0"], w["v0"], w["attractor"], w["c"])
    rc = cref.rollout_fwd(n, np.float32((n / 1000) / n), w["c"], w["attractor"], w["segments"], w["x0"], w["v0"],
                          want_traj=True)
    assert r["hit"].sum() > 20, "test input must exercise collisions"
    assert_bitexact(r["x_T"], rc["x_T"], "x_T")
    assert_bitexact(r["v_T"], rc["v_T"], "v_T")
    assert np.array_equal(r["idx"], rc["idx"]) and np.array_equal(r["hit"], rc["hit"])
    assert_bitexact(r["trajectory"].coordinate, rc["traj_x"], "traj x")
    assert_bitexact(r["trajectory"].velocity, rc["traj_v"], "traj v")
    assert_bitexact(r["trajectory"].acceleration, rc["traj_a"], "traj a")


def test_torch_forward_bitexact_and_c_adjoint_matches_autograd(small, cref):
    w, n = small, 300
    (res, aux) = ot.value_and_grad(n / 1000, n, w["segments"], w["x0"], w["v0"], w["target"], w["attractor"],
                                   w["c"], grad_wrt_x0=True)
    (loss, (xT, vT)), gv = res
    g = cref.value_and_grad(n, np.float32((n / 1000) / n), w["c"], w["attractor"], w["target"], w["segments"],
                            w["x0"], w["v0"])
    assert_bitexact(xT, g["x_T"], "x_T")
    assert_bitexact(vT, g["v_T"], "v_T")
    assert np.array_equal(aux["idx"], g["idx"]) and np.array_equal(aux["hit"], g["hit"])
    assert np.isfinite(gv).all() and np.isfinite(g["grad_v0"]).all()
    # hand adjoint (branch form, its own op order) vs autograd (select form): fp32 rounding only
    assert grad_rel_err(g["grad_v0"], gv).max() < 1e-4
    assert grad_rel_err(g["grad_x0"], aux["grad_x0"]).max() < 1e-4
    assert abs(g["loss_per_ball"].sum(dtype=np.float64) - loss) < 1e-4 * abs(loss)


def test_fp32_adjoints_vs_float64_truth(small, cref):
    """Calibrates the gradient tolerance: the reference-order fp32 adjoint (C) and torch autograd
    are both within rounding noise of the float64 adjoint evaluated on the same fp32 trajectory —
    median ~1e-7, but a few 1e-4 on the worst ball (bounces amplify rounding)."""
    w, n = small, 300
    dt = np.float32((n / 1000) / n)
    truth, _ = cref.grad_truth64(n, dt, w["c"], w["attractor"], w["target"], w["segments"], w["x0"], w["v0"])
    seq = cref.value_and_grad(n, dt, w["c"], w["attractor"], w["target"], w["segments"], w["x0"], w["v0"])["grad_v0"]
    (res, _) = ot.value_and_grad(n / 1000, n, w["segments"], w["x0"], w["v0"], w["target"], w["attractor"], w["c"])
    e_seq, e_torch = grad_rel_err(seq, truth), grad_rel_err(res[1], truth)
    assert np.median(e_seq) < 1e-6 and np.median(e_torch) < 1e-6
    assert e_seq.max() < 2e-3 and e_torch.max() < 2e-3
    assert np.quantile(e_seq, 0.95) < 1e-4 and np.quantile(e_torch, 0.95) < 1e-4


def test_per_env_scenes(cref):
    E, n = 24, 200
    segs = syn.make_env_maps(7, E, 20, domain=(-4, 4, 0, 6))
    x0, v0 = syn.make_init_states(8, E, segs)
    c = onp.make_constants(syn.reference_constants())
    A = np.array(syn.REFERENCE_ATTRACTOR, np.float32)
    r = onp.run_sim(n / 1000, n, segs, x0, v0, A, c)
    rc = cref.rollout_fwd(n, np.float32((n / 1000) / n), c, A, segs, x0, v0)
    assert r["hit"].sum() > 0
    assert_bitexact(r["x_T"], rc["x_T"])
    assert np.array_equal(r["idx"], rc["idx"])


def test_fp64_autograd_vs_finite_differences():
    """The adjoint differentiates the right function: fp64 autograd vs central differences on a
    collision-rich but smooth-in-v0 neighbourhood (same discrete history)."""
    w = syn.make_workload("c2", n_balls=48)
    n = 200
    c = onp.make_constants(w["constants"])
    args = (n / 1000, n, w["segments"], w["x0"])
    (res, aux) = ot.value_and_grad(*args, w["v0"], w["target"], w["attractor"], c, dtype=torch.float64)
    gv = res[1]
    idx, hit = aux["idx"], aux["hit"]
    eps = 1e-6
    checked = 0
    for comp in range(2):
        vp = w["v0"].astype(np.float64).copy(); vm = vp.copy()
        vp[:, comp] += eps; vm[:, comp] -= eps

        lp = _loss64(args, vp, w, c, idx, hit)
        lm = _loss64(args, vm, w, c, idx, hit)
        fd = (lp - lm) / (2 * eps)
        ok = np.abs(fd - gv[:, comp]) <= 1e-4 * np.maximum(1.0, np.abs(gv[:, comp]))
        checked += int(ok.sum())
        assert ok.mean() > 0.97, (fd[~ok], gv[~ok, comp])
    assert checked > 80


def _loss64(args, v64, w, c, idx, hit):
    """Per-ball L1 loss of an fp64 rollout from a float64 velocity, replaying (idx, hit)."""
    sim_time, n, segs, x0 = args
    out = ot.rollout(sim_time, n, segs, x0, v64, w["attractor"], c, dtype=torch.float64,
                     forced_idx=idx, forced_hit=hit)
    tgt = torch.tensor(w["target"], dtype=torch.float64).reshape(1, 2)
    return torch.abs(out["x_T"].detach() - tgt).sum(dim=1).numpy()


# ---- semantics the oracle must keep -------------------------------------------------------------
def test_argmin_first_index_on_exact_ties():
    # two segments sharing the vertex (1,0): the point (2,0) is equidistant -> first index wins
    segs = np.array([[[0, 0], [1, 0]], [[1, 0], [1, 1]], [[1, 0], [0, -1]]], np.float32)
    idx, dist, seg = onp.closest_segment(np.array([[2.0, 0.0]], np.float32), segs)
    assert idx[0] == 0 and dist[0] == 1.0
    idx, _, _ = onp.closest_segment(np.array([[2.0, 0.0]], np.float32), segs[::-1].copy())
    assert idx[0] == 0


def test_degenerate_segment_is_nan_and_wins(cref):
    segs = np.array([[[0, 0], [1, 0]], [[5, 5], [5, 5]], [[0, 1], [1, 1]]], np.float32)
    pts = np.array([[0.5, 0.1]], np.float32)
    idx, dist, _ = onp.closest_segment(pts, segs)
    assert idx[0] == 1 and np.isnan(dist[0])  # 0/0 -> NaN counts as minimal (argmin semantics)
    ci, cd = cref.closest_segment(pts, segs)
    assert ci[0] == 1 and np.isnan(cd[0])


def test_inner_check_matches_geometry(cref):
    segs = syn.make_map(3, 40)
    rng = np.random.default_rng(1)
    pts = rng.uniform([-13, -1], [13, 21], size=(4000, 2)).astype(np.float32)
    a = onp.points_inside_any_polygon(pts, segs)
    b = cref.points_inside(pts, segs)
    assert np.array_equal(a, b)
    _, inside64 = syn._dist_and_inside(pts.astype(np.float64), segs)
    assert (a != inside64).mean() < 1e-3 and 0.02 < a.mean() < 0.6


def test_closest_segment_c_vs_numpy(cref):
    segs = syn.make_map(5, 67)
    rng = np.random.default_rng(2)
    pts = rng.uniform([-13, -1], [13, 21], size=(3000, 2)).astype(np.float32)
    # put a third of the points exactly on shared vertices' bisectors: ties galore
    pts[:1000] = segs[rng.integers(0, len(segs), 1000), 0] + np.float32(0.25)
    i1, d1, _ = onp.closest_segment(pts, segs)
    i2, d2 = cref.closest_segment(pts, segs)
    assert np.array_equal(i1, i2)
    assert_bitexact(d1, d2)
