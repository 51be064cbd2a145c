"""NumPy float32, op-for-op restatement of the reference rollout (forward truth).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED by the
reference for everything except the Euler/force arithmetic (notebook cell 7).

Follows, line by line:
  * ``doper/sim/ball_sim.py``  (reference)  :15-25 potential, :28-45 friction,
    :48-61 acceleration, :64-79 Euler, :82-132 resolve_collision, :135-168
    collide, :171-209 sim_step, :212-253 _run_sim, :256-286 _compute_loss,
    :289-324 reduce_loss.
  * ``doper/sim/jax_geometry.py`` (reference) :12-24 normal, :41-53 sign,
    :97-116 inner check, :119-138 projection, :141-156 distance, :159-177
    closest segment.

Arithmetic rules: every value is IEEE binary32; python
float constants are rounded to f32 once (JAX weak typing); a 2-vector ``dot`` is
``a0*b0 + a1*b1`` with three roundings and NO fused multiply-add; ``norm`` is
``sqrt(a0*a0 + a1*a1)``; ``clip(x, lo, hi) = min(max(x, lo), hi)``.  All
functions are vectorised over a leading ball axis ``B``; the scene is either
shared (``segments [S,2,2]``) or per-environment (``segments [B,S,2,2]``, an
extension the reference cannot express because it broadcasts the scene,
``ball_sim.py:332``).
"""
from __future__ import annotations

from collections import namedtuple
from typing import Dict, Optional

import numpy as np

f32 = np.float32

BallState = namedtuple("BallState", ["coordinate", "velocity", "acceleration"])  # ball_sim.py:12
Consts = namedtuple("Consts", ["radius", "mass", "friction", "elasticity", "k_s", "g"])

G_FREE_FALL = 9.8  # ball_sim.py:29 default argument


def make_constants(constants: Dict[str, float]) -> Consts:
    """f32 runtime scalars, as the reference's dict leaves become under jit.

    ``mass`` may be absent: ``ball_trainer.py:22-23`` derives it in host double as
    ``4*pi*r**3/3 * density`` and only then does it get rounded to f32.
    """
    c = dict(constants)
    if "mass" not in c:
        volume = 4 * np.pi * (float(c["radius"]) ** 3) / 3  # ball_trainer.py:22
        c["mass"] = volume * float(c["density"])  # ball_trainer.py:23
    return Consts(
        radius=f32(c["radius"]),
        mass=f32(c["mass"]),
        friction=f32(c["rolling_friction_coefficient"]),
        elasticity=f32(c["walls_elasticity"]),
        k_s=f32(c["attractor_strength"]),
        g=f32(G_FREE_FALL),
    )


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


# --------------------------------------------------------------------------
# O(1) physics
# --------------------------------------------------------------------------
def potential_force(x: np.ndarray, attractor: np.ndarray) -> np.ndarray:
    """``-grad(sum((x-A)**2))`` = ``-(2*(x-A))`` (ball_sim.py:15-25,192,124)."""
    return -(f32(2.0) * (x - attractor))


def friction_force(v: np.ndarray, c: Consts) -> np.ndarray:
    """``-clip(v,-1,1) * m * g * f / r`` evaluated left to right (ball_sim.py:43-45)."""
    return ((((-np.clip(v, f32(-1.0), f32(1.0))) * c.mass) * c.g) * c.friction) / c.radius


def acceleration(pot: np.ndarray, fr: np.ndarray, c: Consts) -> np.ndarray:
    """ball_sim.py:61."""
    return (pot + fr) / c.mass


# The one place where the reference's only recorded output (notebook cell 7) shows that the
# authors' XLA:CPU did NOT evaluate the expression "as written": the coordinate update
# ``x + v'*dt`` of get_new_state / get_new_cv was contracted into a single-rounding fused
# multiply-add, while the velocity update ``v + a*dt`` kept two roundings.  With this switch
# on, all 126 printed float32 values of cell 7 are reproduced bit-exactly; with it off, 33 of
# them are off by exactly 1 ulp (tests/test_oracle.py::test_notebook_cell7_*).  Every other
# mul+add of the path is unpinned by the reference and stays "as written" (SURVEY App. A).
CONTRACT_POSITION_UPDATE = True


def fma_f32(a, b, c) -> np.ndarray:
    """Correctly rounded float32 ``a*b + c`` (single rounding), vectorised.

    The product of two binary32 values is exact in binary64; the binary64 sum is then
    repaired for the (rare) double-rounding case with an error-free TwoSum before the
    final rounding to binary32.
    """
    a, b, c = np.broadcast_arrays(np.asarray(a, np.float32), np.asarray(b, np.float32),
                                  np.asarray(c, np.float32))
    with np.errstate(invalid="ignore", over="ignore"):
        p = a.astype(np.float64) * b.astype(np.float64)
        c64 = c.astype(np.float64)
        s = p + c64
        bb = s - p  # TwoSum (Knuth): err = exact(p + c) - s
        err = (p - (s - bb)) + (c64 - bb)
        bits = s.view(np.uint64) if s.flags.c_contiguous else np.ascontiguousarray(s).view(np.uint64)
        tie = ((bits & np.uint64(0x1FFFFFFF)) == np.uint64(0x10000000)) & (err != 0) & np.isfinite(s)
        if np.any(tie):
            s = np.where(tie, np.nextafter(s, np.where(err > 0, np.inf, -np.inf)), s)
        return s.astype(np.float32)


def euler(x: np.ndarray, v: np.ndarray, a: np.ndarray, dt) -> tuple:
    """Semi-implicit Euler, ``get_new_state`` (ball_sim.py:75-76). ``dt``: f32 scalar or [B,1].

    ``v' = v + a*dt`` (two roundings); ``x' = fma(v', dt, x)`` (one rounding) — see
    ``CONTRACT_POSITION_UPDATE`` above."""
    v1 = v + a * dt
    x1 = fma_f32(v1, dt, x) if CONTRACT_POSITION_UPDATE else x + v1 * dt
    return x1, v1


# --------------------------------------------------------------------------
# geometry
# --------------------------------------------------------------------------
def segment_projection(px, py, s0x, s0y, s1x, s1y):
    """jax_geometry.py:132-136.  Returns (proj_x, proj_y, ratio). Broadcasts."""
    svx = s1x - s0x
    svy = s1y - s0y
    seg_len_sq = svx * svx + svy * svy
    pvx = px - s0x
    pvy = py - s0y
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = (pvx * svx + pvy * svy) / seg_len_sq
    ratio = np.minimum(np.maximum(ratio, f32(0.0)), f32(1.0))  # .clip(0, 1); NaN propagates
    return s0x + ratio * svx, s0y + ratio * svy, ratio


def distances_to_segments(points: np.ndarray, segments: np.ndarray) -> np.ndarray:
    """[B,2] x ([S,2,2] | [B,S,2,2]) -> [B,S] (jax_geometry.py:141-156)."""
    points = _f32(points)
    segments = _f32(segments)
    px = points[:, 0:1]
    py = points[:, 1:2]
    s0x, s0y = segments[..., 0, 0], segments[..., 0, 1]
    s1x, s1y = segments[..., 1, 0], segments[..., 1, 1]
    prx, pry, _ = segment_projection(px, py, s0x, s0y, s1x, s1y)
    dx = px - prx
    dy = py - pry
    with np.errstate(invalid="ignore"):
        return np.sqrt(dx * dx + dy * dy)


def closest_segment(points: np.ndarray, segments: np.ndarray):
    """find_closest_segment_to_points_batch (jax_geometry.py:159-177).

    Returns (idx int32 [B], dist f32 [B], seg f32 [B,2,2]).  ``argmin`` returns
    the first minimal index and treats NaN as minimal (numpy semantics; SURVEY
    Appendix D).
    """
    segments = _f32(segments)
    d = distances_to_segments(points, segments)
    idx = np.argmin(d, axis=1).astype(np.int32)
    rows = np.arange(d.shape[0])
    dist = d[rows, idx]
    seg = segments[idx] if segments.ndim == 3 else segments[rows, idx]
    return idx, dist, seg


def points_inside_any_polygon(points: np.ndarray, segments: np.ndarray) -> np.ndarray:
    """if_points_inside_any_polygon (jax_geometry.py:97-116) -> bool [N].

    ``segments`` is polygon-major with exactly 3 segments per polygon (:113).
    """
    points = _f32(points)
    if points.ndim == 1:
        points = points.reshape(1, -1)  # :108-109
    segments = _f32(segments)
    n_poly = segments.shape[0] // 3
    s0x, s0y = segments[:, 0, 0][None], segments[:, 0, 1][None]
    svx = segments[:, 1, 0][None] - s0x  # :21
    svy = segments[:, 1, 1][None] - s0y
    nx, ny = svy, -svx  # :22
    with np.errstate(divide="ignore", invalid="ignore"):
        nrm = np.sqrt(nx * nx + ny * ny)
        nhx, nhy = nx / nrm, ny / nrm  # :23
        pvx = points[:, 0:1] - s0x  # :52
        pvy = points[:, 1:2] - s0y
        sign = np.sign(pvx * nhx + pvy * nhy)  # :53
        neg = sign < 0  # :112 (NaN -> False)
    neg = neg.reshape(points.shape[0], n_poly, 3)  # :113
    return np.any(np.all(neg, axis=-1), axis=-1)  # :114-115


# --------------------------------------------------------------------------
# collision
# --------------------------------------------------------------------------
def resolve_collision(x, v, x1, v1, a, seg, attractor, c: Consts, dt):
    """ball_sim.py:97-132, vectorised over balls; ``seg`` is [B,2,2]."""
    s0x, s0y, s1x, s1y = seg[:, 0, 0], seg[:, 0, 1], seg[:, 1, 0], seg[:, 1, 1]
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        c1x, c1y, _ = segment_projection(x1[:, 0], x1[:, 1], s0x, s0y, s1x, s1y)  # :99
        ox = c1x - x[:, 0]  # :100 (OLD coordinate)
        oy = c1y - x[:, 1]
        od = np.sqrt(ox * ox + oy * oy)  # :101
        nx, ny = ox / od, oy / od  # :102
        toi = np.abs(od - c.radius) / np.abs(nx * v1[:, 0] + ny * v1[:, 1])  # :104-106
        toi = np.minimum(np.maximum(toi, f32(0.0)), dt)  # :108
        xs_, vs_ = euler(x, v, a, toi[:, None])  # :110 -> get_new_state :75-76
        vsx, vsy = vs_[:, 0], vs_[:, 1]
        xsx, xsy = xs_[:, 0], xs_[:, 1]
        c2x, c2y, _ = segment_projection(xsx, xsy, s0x, s0y, s1x, s1y)  # :113
        o2x = c2x - xsx  # :114
        o2y = c2y - xsy
        o2d = np.sqrt(o2x * o2x + o2y * o2y)  # :115
        n2x, n2y = o2x / o2d, o2y / o2d  # :116
        k = n2x * vsx + n2y * vsy
        vnx, vny = n2x * k, n2y * k  # :117
        vpx, vpy = vsx - vnx, vsy - vny  # :118
        vax = vpx - c.elasticity * vnx  # :119
        vay = vpy - c.elasticity * vny
        xs = np.stack([xsx, xsy], axis=1)
        va = np.stack([vax, vay], axis=1)
        pot = potential_force(xs, attractor)  # :124 (no attractor_strength here)
        fr = friction_force(va, c)  # :125-130
        ab = acceleration(pot, fr, c)  # :131
        tau = (dt - toi)[:, None]  # :132
        xb, vb = euler(xs, va, ab, tau)
    return xb, vb, ab


def sim_step(x, v, segments, attractor, c: Consts, dt):
    """One ``sim_step`` (ball_sim.py:171-209) for a batch of balls.

    Returns (x', v', a', idx, hit, dist).  Batched ``lax.cond`` is a select
    (ball_sim.py:163-168 under vmap :332); forward values equal a true branch.
    """
    pot = potential_force(x, attractor)  # :192
    fr = friction_force(v, c)  # :193-198
    a = acceleration(c.k_s * pot, fr, c)  # :199-201
    x1, v1 = euler(x, v, a, dt)  # :202
    idx, dist, seg = closest_segment(x1, segments)  # :203-205
    with np.errstate(invalid="ignore"):
        hit = dist <= c.radius  # :164 (NaN -> False)
    if hit.any():
        h = np.nonzero(hit)[0]
        xb, vb, ab = resolve_collision(x[h], v[h], x1[h], v1[h], a[h], seg[h], attractor, c, dt)
        x1 = x1.copy()
        v1 = v1.copy()
        a = a.copy()
        x1[h], v1[h], a[h] = xb, vb, ab
    return x1, v1, a, idx, hit, dist


def run_sim(
    sim_time,
    n_steps: int,
    segments,
    coordinate_init,
    velocity_init,
    attractor,
    constants,
    want_trajectory: bool = True,
):
    """``_run_sim`` (ball_sim.py:212-253) over a batch.

    Returns dict with x_T, v_T [B,2]; trajectory BallState of [n,B,2] (or None);
    idx [n,B] int32; hit [n,B] bool.
    """
    c = constants if isinstance(constants, Consts) else make_constants(constants)
    dt = f32(sim_time / n_steps)  # :237 host double, rounded once
    x = _f32(coordinate_init).reshape(-1, 2).copy()
    v = _f32(velocity_init).reshape(-1, 2).copy()
    attractor = _f32(attractor).reshape(1, 2)
    segments = _f32(segments)
    B = x.shape[0]
    idx_log = np.empty((n_steps, B), np.int32)
    hit_log = np.empty((n_steps, B), bool)
    tx = np.empty((n_steps, B, 2), np.float32) if want_trajectory else None
    tv = np.empty((n_steps, B, 2), np.float32) if want_trajectory else None
    ta = np.empty((n_steps, B, 2), np.float32) if want_trajectory else None
    for t in range(n_steps):  # lax.scan :245
        x, v, a, idx, hit, _ = sim_step(x, v, segments, attractor, c, dt)
        idx_log[t], hit_log[t] = idx, hit
        if want_trajectory:
            tx[t], tv[t], ta[t] = x, v, a
    traj = BallState(tx, tv, ta) if want_trajectory else None
    return {"x_T": x, "v_T": v, "trajectory": traj, "idx": idx_log, "hit": hit_log}


def compute_loss(x_T: np.ndarray, target) -> np.ndarray:
    """Per-ball L1 end-point loss (ball_sim.py:286): ``|dx| + |dy|`` -> [B]."""
    d = np.abs(x_T - _f32(target).reshape(1, 2))
    return d[:, 0] + d[:, 1]


def reduce_loss(per_ball: np.ndarray) -> np.float32:
    """Sum over balls (ball_sim.py:324). Reduction order is XLA's, hence only a
    toleranced comparison is meaningful; we use sequential f32 ``np.sum`` pairwise."""
    return np.sum(per_ball, dtype=np.float32)
