"""ORACLE (test infrastructure only — never imported by doper_b200): numpy float32 restatement of
the reference's range-sensor path, operation for operation.

  line_ray_intersection_point / batch_line_ray_intersection_point
      doper/sim/jax_geometry.py:180-261   (JAX, float32: x64 is not enabled, scripts/train.py:7-27)
  SimpleRangeSensor / UndirectedRangeSensor.get_observation
      doper/agents/observations.py:29-110 (host numpy on the JAX result)
  RangeSensingAgent.get_observations
      doper/agents/range_sensor_agent.py:17-41

Parity status: UNPINNED by the reference (it has no tests or recorded outputs for this path and
cannot be run here: JAX is absent).  Pinned only by this restatement plus hand-computed cases in
tests/test_sensor_oracle.py.

Dtype note.  The reference computes the ranges on the host as ``norm(points_f32 - position)``: in
float32 when the position array is float32, in float64 when it is float64 (numpy promotion).  The
kernels and this oracle take float32 positions (what the rollout consumes and produces), i.e. the
float32 case; `ranges_f64_positions` gives the float64 variant for comparison (<= 1 ulp apart).
"""
from __future__ import annotations

import numpy as np

F = np.float32
INF = F(np.inf)


def ray_directions(heading_direction, angle_range_deg: float, angle_step_deg: float) -> np.ndarray:
    """observations.py:50-58: float64 numpy angles (arange end may or may not include a 19th ray)."""
    angle_range_rad = angle_range_deg / 180 * np.pi
    angle_step_rad = angle_step_deg / 180 * np.pi
    heading_angle = np.arctan2(heading_direction[1], heading_direction[0])
    ray_angles = np.arange(heading_angle - angle_range_rad / 2, heading_angle + angle_range_rad / 2, angle_step_rad)
    return np.concatenate([np.cos(ray_angles)[..., np.newaxis], np.sin(ray_angles)[..., np.newaxis]], axis=-1)


def batch_line_ray_intersection_point(ray_origins, ray_directions_, segments) -> np.ndarray:
    """[R,2] x [R,2] x [S,2,2] -> [R,S,2] float32 (jax_geometry.py:180-261), vectorised."""
    o = np.asarray(ray_origins, F).reshape(-1, 1, 2)
    d = np.asarray(ray_directions_, F).reshape(-1, 1, 2)
    seg = np.asarray(segments, F).reshape(1, -1, 2, 2)
    with np.errstate(all="ignore"):
        nrm = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1])          # :196 np.linalg.norm
        dx, dy = d[..., 0] / nrm, d[..., 1] / nrm
        v1x, v1y = o[..., 0] - seg[..., 0, 0], o[..., 1] - seg[..., 0, 1]     # :197
        v2x, v2y = seg[..., 1, 0] - seg[..., 0, 0], seg[..., 1, 1] - seg[..., 0, 1]  # :198
        v3x, v3y = -dy, dx                                                     # :199
        denom = v2x * v3x + v2y * v3y                                          # :200
        t1 = (v2x * v1y - v2y * v1x) / denom                                   # :203 cross of 2-vectors
        t2 = (v1x * v3x + v1y * v3y) / denom                                   # :204
        miss = (denom == 0) | (t1 < 0) | (t2 < 0) | (t2 > 1)                   # :205, :213-214
        px = o[..., 0] + t1 * dx                                               # :209
        py = o[..., 1] + t1 * dy
    px = np.where(miss, INF, px).astype(F)
    py = np.where(miss, INF, py).astype(F)
    return np.stack([px, py], axis=-1)


def sensor_observation(position, ray_dirs, segments, distance_range: float, positions_dtype=np.float32):
    """SimpleRangeSensor.get_observation(return_intersection_points=True) for ONE position
    (observations.py:59-76) -> (ranges [R], points [R,2], closest index [R])."""
    position = np.asarray(position, positions_dtype)
    R = len(ray_dirs)
    origins = position.reshape(1, -1).repeat(R, axis=0)
    pts = batch_line_ray_intersection_point(origins, ray_dirs, segments)       # f32, like the JAX result
    with np.errstate(all="ignore"):
        diff = pts - position.reshape(1, 1, -1)
        ranges = np.sqrt(diff[..., 0] * diff[..., 0] + diff[..., 1] * diff[..., 1])  # :66 norm over 2 entries
    closest = np.argmin(ranges, axis=-1)                                        # :69 (first minimum)
    ray_idx = np.arange(R)
    points = pts[ray_idx, closest, :].copy()
    rng = ranges[ray_idx, closest].copy()
    points[rng > distance_range] = np.inf                                       # :72
    rng[rng > distance_range] = distance_range                                  # :73
    return rng, points, closest.astype(np.int32)


def sensor_observations(positions, ray_dirs, segments, distance_range: float):
    out = [sensor_observation(p, ray_dirs, segments, distance_range) for p in np.asarray(positions, F)]
    return (np.stack([o[0] for o in out]).astype(F), np.stack([o[1] for o in out]).astype(F),
            np.stack([o[2] for o in out]))


def agent_observations(coordinate_init, velocity_init, ray_dirs, segments, distance_range: float, target):
    """RangeSensingAgent.get_observations (range_sensor_agent.py:32-41) -> float32 [B, 7 + R]."""
    coordinate_init = np.asarray(coordinate_init, F)
    velocity_init = np.asarray(velocity_init, F)
    tgt = np.array(target)  # float64, like onp.array(config[...])
    dist = np.linalg.norm(coordinate_init - tgt, axis=1).reshape(-1, 1)
    direction = (coordinate_init - tgt) / dist
    range_obs, _, _ = sensor_observations(coordinate_init, ray_dirs, segments, distance_range)
    range_obs = range_obs / F(distance_range)                                   # `/=` on a float32 array
    parts = [dist, direction, coordinate_init, velocity_init, range_obs]
    return np.concatenate([np.asarray(p).astype(F) for p in parts], axis=1)
