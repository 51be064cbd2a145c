"""Seeded synthetic scenes and initial states.

The reference's map assets are git-lfs stubs and its samplers are unseeded
(``doper/scenes/single_scene.py:39,59``, ``doper/trainers/ball_trainer.py:52``), so
every benchmark / parity input is generated here.  Scenes keep the reference's
invariants: triangles only, counter-clockwise, polygon-major segment order
``[(p0,p1),(p1,p2),(p2,p0)]`` (``doper/utils/assets.py:116-121``,
``doper/sim/jax_geometry.py:113``), edge lengths on the scale of
``scripts/generate_maps.py:87-99``.

This is input generation in plain numpy (host side, float64 geometry, rounded to
f32 once); it is not part of the measured path and does not touch ``oracle/``.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

# configs/jax_ball.yaml:2-6,11-12 (reference)
REFERENCE_CONSTANTS: Dict[str, float] = {
    "radius": 0.1,
    "density": 3000,
    "rolling_friction_coefficient": 0.007,
    "walls_elasticity": 0.8,
    "attractor_strength": 0.01,
}
REFERENCE_TARGET = (6.0, 4.0)
REFERENCE_ATTRACTOR = (-3.0, 2.0)
BASE_DOMAIN = (-12.0, 12.0, 0.0, 20.0)  # x_lo, x_hi, y_lo, y_hi for 67 triangles
BASE_TRIANGLES = 67


def reference_constants() -> Dict[str, float]:
    """Constants dict as ``BallAttractorTrainer.__init__`` leaves it
    (``ball_trainer.py:21-23``): volume and mass derived in host double."""
    c = dict(REFERENCE_CONSTANTS)
    c["volume"] = 4 * np.pi * (c["radius"] ** 3) / 3
    c["mass"] = c["volume"] * c["density"]
    return c


def domain_for(n_triangles: int) -> Tuple[float, float, float, float]:
    """Scale the base domain so the obstacle area fraction stays constant."""
    s = float(np.sqrt(n_triangles / BASE_TRIANGLES))
    x_lo, x_hi, y_lo, y_hi = BASE_DOMAIN
    cx, cy = 0.5 * (x_lo + x_hi), 0.5 * (y_lo + y_hi)
    hx, hy = 0.5 * (x_hi - x_lo) * s, 0.5 * (y_hi - y_lo) * s
    return (cx - hx, cx + hx, cy - hy, cy + hy)


def _triangles(rng: np.random.Generator, shape, domain) -> np.ndarray:
    """CCW triangles [..., 3, 2] float64."""
    x_lo, x_hi, y_lo, y_hi = domain
    centre = np.stack([rng.uniform(x_lo, x_hi, shape), rng.uniform(y_lo, y_hi, shape)], axis=-1)
    ang = np.sort(rng.uniform(0.0, 2 * np.pi, (*shape, 3)), axis=-1)
    rad = rng.uniform(0.3, 1.8, (*shape, 3))
    pts = centre[..., None, :] + rad[..., None] * np.stack([np.cos(ang), np.sin(ang)], axis=-1)
    # sorted polar angles are CCW unless one angular gap exceeds pi: fix by orientation
    a, b, c = pts[..., 0, :], pts[..., 1, :], pts[..., 2, :]
    area2 = (b[..., 0] - a[..., 0]) * (c[..., 1] - a[..., 1]) - (b[..., 1] - a[..., 1]) * (c[..., 0] - a[..., 0])
    cw = area2 < 0
    pts[cw] = pts[cw][..., [0, 2, 1], :]
    return pts


def _tri_to_segments(pts: np.ndarray) -> np.ndarray:
    """[..., T, 3, 2] -> polygon-major segments [..., 3T, 2, 2] float32."""
    nxt = np.roll(pts, -1, axis=-2)
    seg = np.stack([pts, nxt], axis=-2)  # [..., T, 3, 2(endpoint), 2(xy)]
    return seg.reshape(*pts.shape[:-3], pts.shape[-3] * 3, 2, 2).astype(np.float32)


def make_map(seed: int, n_triangles: int = BASE_TRIANGLES, domain=None) -> np.ndarray:
    """One scene: segments [3T,2,2] float32."""
    domain = domain or domain_for(n_triangles)
    rng = np.random.default_rng(seed)
    return _tri_to_segments(_triangles(rng, (n_triangles,), domain))


def make_env_maps(seed: int, n_envs: int, n_triangles: int = BASE_TRIANGLES, domain=None) -> np.ndarray:
    """Per-environment scenes: segments [E,3T,2,2] float32 (one generator, vectorised)."""
    domain = domain or domain_for(n_triangles)
    rng = np.random.default_rng(seed)
    return _tri_to_segments(_triangles(rng, (n_envs, n_triangles), domain))


def _dist_and_inside(p: np.ndarray, seg: np.ndarray):
    """float64 helper for rejection sampling: p [N,2], seg [S,2,2] or [N,S,2,2]."""
    seg = seg.astype(np.float64)
    s0, s1 = seg[..., 0, :], seg[..., 1, :]
    sv = s1 - s0
    pv = p[:, None, :] - s0
    L = np.maximum((sv * sv).sum(-1), 1e-300)
    t = np.clip((pv * sv).sum(-1) / L, 0.0, 1.0)
    d = np.linalg.norm(pv - t[..., None] * sv, axis=-1)
    cross = sv[..., 1] * pv[..., 0] - sv[..., 0] * pv[..., 1]  # (p-s0).[sv_y,-sv_x]
    neg = (cross < 0).reshape(p.shape[0], -1, 3)
    return d.min(axis=1), neg.all(axis=-1).any(axis=-1)


def make_init_states(seed: int, n_balls: int, segments: np.ndarray, radius: float = 0.1,
                     impulse_sigma: float = 10.0, v_max: float = 50.0,
                     chunk: int = 4096) -> Tuple[np.ndarray, np.ndarray]:
    """(x0, v0) float32 [B,2].

    x0 is rejection-sampled like ``single_scene.py:24-61``: uniform over the scene's
    bounding box (+1 m on the first proposal), rejected when inside a triangle or
    closer than ``radius + radius/2`` to any segment.  v0 = U(-1,2) (``ball_trainer.py:52``)
    + N(0, impulse_sigma^2) as a stand-in for ``impulse/mass``, clamped to ``v_max`` so
    that ``|v| dt < radius`` (no tunnelling).  ``segments`` may be per-env [B,S,2,2].
    """
    rng = np.random.default_rng(seed)
    per_env = segments.ndim == 4
    eps = radius / 2
    x0 = np.empty((n_balls, 2), np.float64)
    for lo in range(0, n_balls, chunk):
        hi = min(n_balls, lo + chunk)
        n = hi - lo
        seg = segments[lo:hi] if per_env else segments
        if per_env:
            mn = seg.reshape(n, -1, 2).min(axis=1).astype(np.float64)
            mx = seg.reshape(n, -1, 2).max(axis=1).astype(np.float64)
        else:
            mn = np.broadcast_to(seg.reshape(-1, 2).min(axis=0).astype(np.float64), (n, 2))
            mx = np.broadcast_to(seg.reshape(-1, 2).max(axis=0).astype(np.float64), (n, 2))
        todo = np.arange(n)
        prop = rng.uniform(mn - 1, mx + 1)
        out = np.empty((n, 2))
        while True:
            # judge acceptance on the f32-rounded proposal, which is what gets used
            p32 = prop.astype(np.float32).astype(np.float64)
            d, inside = _dist_and_inside(p32, seg[todo] if per_env else seg)
            ok = ~(inside | (d <= radius + eps))
            out[todo[ok]] = p32[ok]
            if ok.all():
                break
            todo = todo[~ok]
            prop = rng.uniform(mn[todo], mx[todo])
        x0[lo:hi] = out
    v0 = rng.uniform(-1.0, 2.0, (n_balls, 2))
    if impulse_sigma > 0:
        v0 = v0 + rng.normal(0.0, impulse_sigma, (n_balls, 2))
    speed = np.linalg.norm(v0, axis=1, keepdims=True)
    v0 = v0 * np.minimum(1.0, v_max / np.maximum(speed, 1e-30))
    return x0.astype(np.float32), v0.astype(np.float32)


# ---- named workloads of BASELINE.json -----------------------
WORKLOADS = {
    # name: (balls, triangles, n_steps, per_env, map_seed, state_seed)
    "c1": (16, 67, 300, False, 1001, 2001),
    "c2": (4096, 67, 500, False, 1002, 2002),
    "c3": (65536, 67, 1000, True, 10000, 2003),
    "c4": (16384, 3334, 500, False, 1004, 2004),
    "c5": (1048576, 67, 2000, False, 1005, 2005),
}


def make_workload(name: str, n_balls: int = None, lo: int = 0, hi: int = None, impulse_sigma: float = 10.0):
    """Inputs of a named workload; ``[lo:hi)`` slices balls/envs (rank sharding keeps
    results independent of the GPU count: global arrays are generated, then sliced).
    ``impulse_sigma=0`` is SURVEY.md §8d's "reference regime": v0 = U(-1,2)^2 only, as
    ``ball_trainer.py:52`` draws it before the controller's impulse."""
    balls, tri, n_steps, per_env, map_seed, state_seed = WORKLOADS[name]
    balls = n_balls or balls
    segs = make_env_maps(map_seed, balls, tri) if per_env else make_map(map_seed, tri)
    x0, v0 = make_init_states(state_seed, balls, segs, impulse_sigma=impulse_sigma)
    hi = balls if hi is None else hi
    if per_env:
        segs = segs[lo:hi]
    return {
        "name": name, "segments": segs, "x0": x0[lo:hi], "v0": v0[lo:hi], "n_steps": n_steps,
        "sim_time": n_steps / 1000.0, "per_env": per_env, "constants": reference_constants(),
        "target": np.array(REFERENCE_TARGET, np.float32),
        "attractor": np.array(REFERENCE_ATTRACTOR, np.float32),
    }
