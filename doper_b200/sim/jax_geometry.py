"""Drop-in mirror of the reference ``doper/sim/jax_geometry.py`` for the hot-path functions.

Same names, positional signatures and return nesting as the reference
(``jax_geometry.py:8-9`` containers, ``:97-116`` inner check, ``:159-177`` closest segment,
``:281-305`` scene construction), computed by the sm_100a kernels behind the C ABI.
Inputs may be numpy arrays / lists (host: results come back as numpy) or torch CUDA tensors
(device: results stay on the device, zero-copy).  The module keeps the reference's file name
so that ``from doper.sim.jax_geometry import ...`` call sites only change the package name.
"""
from __future__ import annotations

from collections import namedtuple
from typing import List, Sequence, Tuple, Union

import numpy as np
import torch

from .. import _abi
from .._abi import check, lib
from .._device import as_f32_device, is_device_tensor, ptr, require_cuda, stream_ptr

__all__ = [
    "Polygon", "JaxScene", "create_polygon", "create_scene", "create_env_scenes",
    "find_closest_segment_to_point", "find_closest_segment_to_points_batch",
    "if_points_inside_any_polygon",
    "line_ray_intersection_point", "batch_line_ray_intersection_point", "range_sensor",
]

Polygon = namedtuple("Polygon", ["segments"])  # jax_geometry.py:8


class JaxScene:
    """``JaxScene(segments, polygons)`` (jax_geometry.py:9) plus the prepared device scene.

    ``segments`` is ``[S,2,2]`` (shared scene, the reference's only form) or ``[E,S,2,2]``
    (one scene per environment — an extension, the reference broadcasts the scene,
    ball_sim.py:332).  The staging layout the kernels read (``doper_scene_prepare``) is built
    lazily, once per device, and cached on the object.
    """

    __slots__ = ("segments", "polygons", "_dev")

    def __init__(self, segments, polygons=None):
        self.segments = segments
        self.polygons = polygons if polygons is not None else []
        self._dev = {}

    # namedtuple-style access used by reference code: scene.segments / scene.polygons / unpacking
    def __iter__(self):
        yield self.segments
        yield self.polygons

    def __len__(self):
        return 2

    def __getitem__(self, i):
        return (self.segments, self.polygons)[i]

    @property
    def per_env(self) -> bool:
        return len(self.segments.shape) == 4

    @property
    def n_segments(self) -> int:
        return int(self.segments.shape[-3])

    @property
    def n_scenes(self) -> int:
        return int(self.segments.shape[0]) if self.per_env else 1

    def prepared(self, device: torch.device = None, radius: float = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """(raw segments f32 on device ``[..., S, 4]``, prepared workspace bytes).

        ``radius``: the ball radius of the rollout that is about to use the workspace.  A shared scene carries
        baked per-cell candidate lists valid for radii up to the one it was prepared with
        (``doper_scene_prepare(max_radius)``); asking for a larger one re-bakes in place, stream-ordered."""
        device = device or require_cuda()
        key = (device.type, device.index)
        hit = self._dev.get(key)
        if hit is not None and (radius is None or (hit[2] is not None and radius <= hit[2])):
            return hit[0], hit[1]
        if hit is None:
            raw = as_f32_device(self.segments, device)
            if raw.shape[-2:] != (2, 2) or raw.dim() not in (3, 4):
                raise ValueError(f"segments must be [S,2,2] or [E,S,2,2], got {tuple(raw.shape)}")
            raw = raw.reshape(*raw.shape[:-2], 4).contiguous()
            S = int(raw.shape[-2])
            n = int(raw.shape[0]) if raw.dim() == 3 else 1
            ws = torch.empty(lib.doper_scene_workspace_bytes(n, S), dtype=torch.uint8, device=device)
        else:
            raw, ws = hit[0], hit[1]
            S = int(raw.shape[-2])
            n = int(raw.shape[0]) if raw.dim() == 3 else 1
        r = float(radius) if radius is not None else 0.0
        check(lib.doper_scene_prepare(stream_ptr(), n, S, ptr(raw), r, ptr(ws)), "doper_scene_prepare")
        self._dev[key] = (raw, ws, (r if r > 0.0 else None))
        return raw, ws


def create_polygon(segments) -> Polygon:
    """jax_geometry.py:281-299."""
    if isinstance(segments, (np.ndarray, torch.Tensor)):
        return Polygon(segments=segments)
    return Polygon(segments=np.asarray(segments, dtype=np.float32))


def create_scene(polygons: List[Polygon]) -> JaxScene:
    """jax_geometry.py:302-305: concatenate the polygons' segments, polygon-major."""
    segs = [p.segments for p in polygons]
    if any(isinstance(s, torch.Tensor) for s in segs):
        cat = torch.cat([torch.as_tensor(s, dtype=torch.float32) for s in segs], dim=0)
    else:
        cat = np.concatenate([np.asarray(s, dtype=np.float32) for s in segs], axis=0)
    return JaxScene(segments=cat, polygons=polygons)


def create_env_scenes(segments) -> JaxScene:
    """Extension: per-environment scenes ``[E,S,2,2]`` (3 segments per CCW triangle each)."""
    return JaxScene(segments=segments, polygons=None)


def _as_scene(segments_or_scene) -> JaxScene:
    if isinstance(segments_or_scene, JaxScene):
        return segments_or_scene
    return JaxScene(segments=segments_or_scene, polygons=None)


def _closest(points, segments):
    device = require_cuda()
    on_device = is_device_tensor(points)
    scene = _as_scene(segments)
    raw, ws = scene.prepared(device)
    pts = as_f32_device(points, device).reshape(-1, 2)
    N, S = int(pts.shape[0]), scene.n_segments
    idx = torch.empty(N, dtype=torch.int32, device=device)
    dist = torch.empty(N, dtype=torch.float32, device=device)
    seg = torch.empty((N, 2, 2), dtype=torch.float32, device=device)
    if scene.per_env:  # point i against scene i ([E,S,2,2] scenes: one environment per ball, ball_sim.py:332 vmapped over maps)
        if N != scene.n_scenes:
            raise ValueError(f"per-environment scenes: {N} points for {scene.n_scenes} scenes")
        check(lib.doper_closest_segment_env(stream_ptr(), N, S, ptr(pts), ptr(raw), ptr(ws), ptr(idx), ptr(dist),
                                            ptr(seg)), "doper_closest_segment_env")
    else:
        check(lib.doper_closest_segment(stream_ptr(), N, S, ptr(pts), ptr(raw), ptr(ws), ptr(idx), ptr(dist),
                                        ptr(seg)), "doper_closest_segment")
    if on_device:
        return seg, dist, idx
    return seg.cpu().numpy(), dist.cpu().numpy(), idx.cpu().numpy()


def find_closest_segment_to_points_batch(points, segments_batch):
    """jax_geometry.py:177: ``[N,2] x [S,2,2] -> (closest segments [N,2,2], distances [N])``."""
    seg, dist, _ = _closest(points, segments_batch)
    return seg, dist


def find_closest_segment_to_point(point, segments_batch):
    """jax_geometry.py:159-174: ``[2] x [S,2,2] -> (closest segment [2,2], distance scalar)``."""
    seg, dist, _ = _closest(point, segments_batch)
    return seg[0], dist[0]


def closest_segment_index(points, segments_batch):
    """Extension: also return ``argmin`` itself (int32 ``[N]``), which the reference discards."""
    return _closest(points, segments_batch)


def if_points_inside_any_polygon(points, scene: JaxScene):
    """jax_geometry.py:97-116: bool ``[N]``; a 1-D point is treated as a batch of one (:108-109)."""
    device = require_cuda()
    on_device = is_device_tensor(points)
    scene = _as_scene(scene)
    raw, _ = scene.prepared(device)
    pts = as_f32_device(points, device).reshape(-1, 2)
    N, S = int(pts.shape[0]), scene.n_segments
    flag = torch.empty(N, dtype=torch.uint8, device=device)
    if scene.per_env:
        if N != scene.n_scenes:
            raise ValueError(f"per-environment scenes: {N} points for {scene.n_scenes} scenes")
        check(lib.doper_points_inside_env(stream_ptr(), N, S, ptr(pts), ptr(raw), ptr(flag)), "doper_points_inside_env")
    else:
        check(lib.doper_points_inside(stream_ptr(), N, S, ptr(pts), ptr(raw), ptr(flag)), "doper_points_inside")
    out = flag.to(torch.bool)
    return out if on_device else out.cpu().numpy()


# ------------------------------------------------------------------------------------------
# ray casting (reference jax_geometry.py:180-261)
# ------------------------------------------------------------------------------------------
def batch_line_ray_intersection_point(ray_origins, ray_directions, segments):
    """jax_geometry.py:237-261: ``[R,2] x [R,2] x [S,2,2] -> [R,S,2]`` intersection points,
    ``inf`` where ray and segment do not meet (directions need not be normalised)."""
    device = require_cuda()
    on_device = is_device_tensor(ray_origins) and is_device_tensor(ray_directions)
    scene = _as_scene(segments)
    if scene.per_env:
        raise ValueError("batch_line_ray_intersection_point: per-environment scenes are not supported here")
    raw, _ = scene.prepared(device)
    o = as_f32_device(ray_origins, device).reshape(-1, 2)
    d = as_f32_device(ray_directions, device).reshape(-1, 2)
    if o.shape[0] != d.shape[0]:
        raise ValueError("ray_origins and ray_directions must have the same length")
    R, S = int(o.shape[0]), scene.n_segments
    out = torch.empty((R, S, 2), dtype=torch.float32, device=device)
    check(lib.doper_ray_segment_points(stream_ptr(), R, S, ptr(o), ptr(d), ptr(raw), ptr(out)),
          "doper_ray_segment_points")
    return out if on_device else out.cpu().numpy()


def line_ray_intersection_point(ray_origin, ray_direction, segment):
    """jax_geometry.py:180-218: one ray against one ``[2,2]`` segment -> ``[2]`` point (or inf, inf)."""
    seg = segment.reshape(1, 2, 2) if hasattr(segment, "reshape") else np.asarray(segment, np.float32).reshape(1, 2, 2)
    o = ray_origin.reshape(1, 2) if hasattr(ray_origin, "reshape") else np.asarray(ray_origin, np.float32).reshape(1, 2)
    d = ray_direction.reshape(1, 2) if hasattr(ray_direction, "reshape") else \
        np.asarray(ray_direction, np.float32).reshape(1, 2)
    return batch_line_ray_intersection_point(o, d, seg)[0, 0]


def range_sensor(positions, ray_directions, scene, distance_range: float, want_points: bool = False):
    """All sensors at once (extension of observations.py:29-79, which handles one position per
    call): ``positions [N,2]``, ``ray_directions [R,2]`` shared by every position ->
    ``ranges [N,R]`` (and ``points [N,R,2]``, ``closest index [N,R]`` when asked)."""
    device = require_cuda()
    on_device = is_device_tensor(positions)
    scene = _as_scene(scene)
    _, ws = scene.prepared(device)
    # float64 positions keep their precision in the ranges, like the reference's numpy code (observations.py:50,67)
    f64 = (isinstance(positions, torch.Tensor) and positions.dtype == torch.float64) or \
        (not isinstance(positions, torch.Tensor) and np.asarray(positions).dtype == np.float64)
    if f64:
        pos = torch.as_tensor(np.asarray(positions) if not isinstance(positions, torch.Tensor) else positions,
                              dtype=torch.float64, device=device).reshape(-1, 2).contiguous()
    else:
        pos = as_f32_device(positions, device).reshape(-1, 2)
    dirs = as_f32_device(ray_directions, device).reshape(-1, 2)
    N, R, S = int(pos.shape[0]), int(dirs.shape[0]), scene.n_segments
    if scene.per_env and N != scene.n_scenes:
        raise ValueError(f"per-environment scenes: {N} positions for {scene.n_scenes} scenes")
    ranges = torch.empty((N, R), dtype=torch.float64 if f64 else torch.float32, device=device)
    points = torch.empty((N, R, 2), dtype=torch.float32, device=device) if want_points else None
    idx = torch.empty((N, R), dtype=torch.int32, device=device) if want_points else None
    if f64:
        check(lib.doper_range_sensor_f64(stream_ptr(), N, R, S, 1 if scene.per_env else 0, ptr(pos), ptr(dirs), ptr(ws),
                                         float(distance_range), ptr(ranges), ptr(points), ptr(idx)), "doper_range_sensor_f64")
    else:
        fn = lib.doper_range_sensor_env if scene.per_env else lib.doper_range_sensor
        check(fn(stream_ptr(), N, R, S, ptr(pos), ptr(dirs), ptr(ws), float(distance_range),
                 ptr(ranges), ptr(points), ptr(idx)), "doper_range_sensor")
    if not want_points:
        return ranges if on_device else ranges.cpu().numpy()
    if on_device:
        return ranges, points, idx
    return ranges.cpu().numpy(), points.cpu().numpy(), idx.cpu().numpy()
