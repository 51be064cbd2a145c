"""CPU emulation (numpy fp32, exact FMA) of the broad phase's two geometric shortcuts (csrc/sweep.cuh,
DESIGN.md §5.2b), checked against the oracle's reference distances:

* hit list: built at a centre c with radius rho, it must contain every segment whose REFERENCE
  distance to a query point p is <= r, for every p the displacement guard admits;
* safe disc: a point the safe-disc test admits must have every reference distance > r.

The inputs aim at the boundaries (centres a hair outside the ball radius from a wall, query points at
the edge of the admitted displacement) at several coordinate magnitudes, because the slacks are multiples
of u * M."""
import numpy as np
import pytest

from doper_b200 import synthetic as syn
from oracle.ball_sim_np import distances_to_segments, fma_f32

f32 = np.float32
U = f32(2.0 ** -24)
UP = f32(1.0 + 2.0 ** -20)          # kCullUp
DOWN = f32(1.0 - 2.0 ** -20)
SAFE_PER_M = f32(128.0) * U          # kCullSafePerM
RHO_CHECK = f32(0.98)                # kCullRhoCheck


def approx_d2(p, segs):
    """sweep.cuh:approx_d2 for points [N,2] x segments [S,2,2] -> [N,S] (same operations, fp32)."""
    px, py = p[:, 0:1], p[:, 1:2]
    s0x, s0y = segs[:, 0, 0][None], segs[:, 0, 1][None]
    svx, svy = segs[:, 1, 0][None] - s0x, segs[:, 1, 1][None] - s0y
    rl = f32(1.0) / (svx * svx + svy * svy)
    pvx, pvy = px - s0x, py - s0y
    dot = fma_f32(pvx, svx, pvy * svy)
    t = np.clip(dot * rl, f32(0), f32(1))
    ex, ey = fma_f32(-t, svx, pvx), fma_f32(-t, svy, pvy)
    return fma_f32(ex, ex, ey * ey)


def refresh(c, rho, r, segs, scale):
    """glist_refresh (hit list) for centres c [N,2], radii rho [N] -> membership [N,S], safe2 [N], rho2 [N]."""
    M = np.maximum(scale, np.maximum(np.abs(c[:, 0]), np.abs(c[:, 1])) + rho).astype(f32)
    reach = (fma_f32(np.full_like(M, SAFE_PER_M), M, (rho + r).astype(f32)) * UP).astype(f32)
    T = ((reach * reach) * UP).astype(f32)
    a = approx_d2(c, segs)
    member = a <= T[:, None]
    rho2 = ((RHO_CHECK * rho) * (RHO_CHECK * rho)).astype(f32)
    sigma = (np.sqrt(a.min(1)).astype(f32) - (fma_f32(np.full_like(M, SAFE_PER_M), M, np.full_like(M, r)) * UP).astype(f32)).astype(f32)
    safe2 = np.where(sigma > 0, np.minimum((sigma * sigma) * DOWN, rho2), f32(-1)).astype(f32)
    return member, safe2, rho2


@pytest.mark.parametrize("scale,offset", [(1.0, 0.0), (1.0, 3.0e4), (1.0e3, 0.0), (1.0e-2, 0.0)])
def test_hit_list_and_safe_disc_are_sound(scale, offset):
    rng = np.random.default_rng(11)
    segs = (syn.make_map(1002, 67).astype(np.float64) * scale + offset).astype(f32)   # [201,2,2]
    r = f32(0.1 * scale)
    sc_scale = f32(np.abs(segs).max())
    S = segs.shape[0]
    n = 6000
    # centres: a third random, two thirds at r * (1 + tiny .. 3) from a random point of a random segment
    k = rng.integers(0, S, n)
    t = rng.uniform(0, 1, (n, 1))
    on = segs[k, 0].astype(np.float64) + t * (segs[k, 1].astype(np.float64) - segs[k, 0].astype(np.float64))
    d = segs[k, 1].astype(np.float64) - segs[k, 0].astype(np.float64)
    nrm = np.stack([d[:, 1], -d[:, 0]], 1) / np.linalg.norm(d, axis=1, keepdims=True)
    gap = float(r) * (1.0 + 10.0 ** rng.uniform(-7, 0.5, (n, 1)))
    c = on + nrm * gap * rng.choice([-1.0, 1.0], (n, 1))
    lo, hi = segs.min(), segs.max()
    c[: n // 3] = rng.uniform(lo, hi, (n // 3, 2))
    c = c.astype(f32)
    rho = (float(r) * 10.0 ** rng.uniform(-1.5, 1.5, n)).astype(f32)
    member, safe2, rho2 = refresh(c, rho, r, segs, sc_scale)

    for trial in range(4):
        # query points: towards the nearest wall and in random directions, at the edge of what is admitted
        ang = rng.uniform(0, 2 * np.pi, n)
        dirs = np.stack([np.cos(ang), np.sin(ang)], 1)
        if trial % 2 == 0:
            dirs = -nrm * np.sign((c.astype(np.float64) - on) @ np.ones(2) * 0 + ((c.astype(np.float64) - on) * nrm).sum(1))[:, None]
        frac = rng.choice([1.0, 0.999999, 0.99, 0.5], n)
        radius = np.sqrt(np.where((trial >= 2) & (safe2 > 0), safe2, rho2).astype(np.float64)) * frac
        p = (c.astype(np.float64) + dirs * radius[:, None]).astype(f32)
        dx, dy = p[:, 0] - c[:, 0], p[:, 1] - c[:, 1]
        dd2 = (dx * dx + dy * dy).astype(f32)
        d_ref = distances_to_segments(p, segs)                       # [n,S] reference arithmetic
        touching = d_ref <= r
        admitted = dd2 <= rho2
        # (1) every touching segment of an admitted point is a list member
        missed = admitted[:, None] & touching & ~member
        assert not missed.any(), f"hit list misses a touching segment ({int(missed.sum())} cases, trial {trial})"
        # (2) a point inside the safe disc touches nothing
        safe = dd2 <= safe2
        bad = safe & touching.any(1)
        assert not bad.any(), f"safe disc admits a touching point ({int(bad.sum())} cases, trial {trial})"
    # the shortcuts must actually fire on this data, or the test proves nothing
    assert (safe2 > 0).mean() > 0.2 and member.any(1).mean() > 0.2
