"""GPU parity of the range-sensor path against oracle/ray_np.py, through the
mirror of the reference API (doper_b200.sim.jax_geometry / doper_b200.agents).  Bar: intersection
points, ranges, closest-segment indices and observation rows bit-exact (float32)."""
import numpy as np
import pytest
import torch

from doper_b200 import synthetic as syn
from oracle import ray_np as orc
from tests.helpers import assert_bitexact

pytestmark = pytest.mark.gpu


def same_with_inf(a, b, what):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.array_equal(np.isinf(a), np.isinf(b)), f"{what}: hit / miss pattern differs"
    assert_bitexact(a, b, what)


@pytest.mark.parametrize("R,T", [(1, 1), (37, 67), (700, 67), (64, 3334)])
def test_batch_line_ray_intersection_point_bitexact(R, T):
    from doper_b200.sim import jax_geometry as geo
    segs = syn.make_map(3, T)
    rng = np.random.default_rng(R)
    lo, hi = segs.reshape(-1, 2).min(0), segs.reshape(-1, 2).max(0)
    o = rng.uniform(lo, hi, (R, 2)).astype(np.float32)
    d = rng.normal(size=(R, 2)).astype(np.float32) * rng.uniform(0.1, 30, (R, 1)).astype(np.float32)
    # boundary cases: rays that start ON a segment endpoint, axis-aligned rays, rays through vertices
    k = R // 4
    if k:
        o[:k] = segs[rng.integers(0, len(segs), k), 0]
        d[k:2 * k] = np.array([1.0, 0.0], np.float32)
        v = segs[rng.integers(0, len(segs), k), 1]
        d[2 * k:3 * k] = v - o[2 * k:3 * k]
    ref = orc.batch_line_ray_intersection_point(o, d, segs)
    got = geo.batch_line_ray_intersection_point(o, d, segs)
    assert isinstance(got, np.ndarray) and got.shape == (R, len(segs), 2)
    same_with_inf(got, ref, "intersection points")
    assert np.isfinite(ref).any()
    got_dev = geo.batch_line_ray_intersection_point(torch.tensor(o, device="cuda"), torch.tensor(d, device="cuda"), segs)
    assert got_dev.is_cuda
    same_with_inf(got_dev.cpu().numpy(), ref, "intersection points (device inputs)")
    p1 = geo.line_ray_intersection_point(o[0], d[0], segs[0])
    same_with_inf(p1, ref[0, 0], "single ray / single segment")


@pytest.mark.parametrize("N,T,dr", [(1, 67, 1000.0), (333, 67, 1000.0), (5000, 67, 3.0), (200, 3334, 1000.0)])
def test_range_sensor_bitexact(N, T, dr):
    from doper_b200.agents import UndirectedRangeSensor
    segs = syn.make_map(5, T)
    rng = np.random.default_rng(N)
    lo, hi = segs.reshape(-1, 2).min(0), segs.reshape(-1, 2).max(0)
    pos = rng.uniform(lo - 1, hi + 1, (N, 2)).astype(np.float32)
    pos[: N // 5] = segs[rng.integers(0, len(segs), N // 5), 0]  # sensors sitting on vertices: exact ties
    sensor = UndirectedRangeSensor(dr, 20)
    dirs = orc.ray_directions((1, 0), 360, 20)
    r_ref, p_ref, i_ref = orc.sensor_observations(pos, dirs.astype(np.float32), segs, dr)
    ranges, points = sensor.get_observations_batch(pos, segs, return_intersection_points=True)
    assert ranges.shape == (N, len(dirs))
    assert_bitexact(ranges, r_ref, "ranges")
    same_with_inf(points, p_ref, "closest intersection points")
    from doper_b200.sim.jax_geometry import range_sensor
    _, _, idx = range_sensor(pos, dirs.astype(np.float32), segs, dr, want_points=True)
    assert np.array_equal(idx, i_ref), "closest-segment index (first on ties, 0 when nothing is hit)"
    # the reference's one-position API
    r1 = sensor.get_observation(pos[0], segs)
    assert r1.shape == (len(dirs),)
    assert_bitexact(r1, r_ref[0], "get_observation(position)")
    r1b, p1b = sensor.get_observation(tuple(pos[0]), segs, return_intersection_points=True)
    same_with_inf(p1b, p_ref[0], "get_observation(..., return_intersection_points=True)")


def test_agent_observations_bitexact():
    from doper_b200.agents import RangeSensingAgent
    cfg = {"sim": {"coordinate_target": [6.0, 4.0], "agent_params": {"distance_range": 1000, "angle_step": 20}}}
    w = syn.make_workload("c2", n_balls=777)
    agent = RangeSensingAgent(cfg)

    class Handler:  # doper/scenes/*: anything with a .jax_scene
        pass
    from doper_b200.sim.jax_geometry import JaxScene
    h = Handler()
    h.jax_scene = JaxScene(w["segments"])
    obs = agent.get_observations(w["x0"], w["v0"], h)
    dirs = orc.ray_directions((1, 0), 360, 20).astype(np.float32)
    ref = orc.agent_observations(w["x0"], w["v0"], dirs, w["segments"], 1000.0, [6.0, 4.0])
    assert isinstance(obs, torch.Tensor) and not obs.is_cuda and obs.dtype == torch.float32
    assert tuple(obs.shape) == ref.shape == (777, 7 + len(dirs)) and ref.shape[1] == 25  # model.input_dim: 25
    assert_bitexact(obs.numpy(), ref, "controller input rows")
    obs_dev = agent.get_observations(torch.tensor(w["x0"], device="cuda"), torch.tensor(w["v0"], device="cuda"), h)
    assert obs_dev.is_cuda
    assert_bitexact(obs_dev.cpu().numpy(), ref, "controller input rows (device inputs)")


# ------------------------------------------------------------------------------------------------
# per-environment scenes (BASELINE configs[2]: every ball has its own map), float64 positions, scenes larger
# than shared memory
# ------------------------------------------------------------------------------------------------
def _env_scenes(E, T, seed=11):
    return np.stack([syn.make_map(seed + e, T) for e in range(E)])  # [E, 3T, 2, 2]


def test_per_environment_queries_and_sensor_match_one_scene_at_a_time():
    """Point / position i against scene i in ONE launch == E separate shared-scene calls == the oracle."""
    from doper_b200.agents import RangeSensingAgent, UndirectedRangeSensor
    from doper_b200.sim import jax_geometry as geo
    from oracle.cref import CRef
    E, T = 37, 20
    segs = _env_scenes(E, T)
    rng = np.random.default_rng(5)
    lo, hi = segs.reshape(E, -1, 2).min(1), segs.reshape(E, -1, 2).max(1)
    pts = rng.uniform(lo - 0.5, hi + 0.5).astype(np.float32)
    pts[::5] = segs[::5, 7, 0]  # on a vertex: exact ties
    scene = geo.JaxScene(segs)
    assert scene.per_env and scene.n_scenes == E
    seg, dist, idx = geo.closest_segment_index(pts, scene)
    inside = geo.if_points_inside_any_polygon(pts, scene)
    cr = CRef("portable")
    for e in range(E):
        i_ref, d_ref = cr.closest_segment(pts[e:e + 1], segs[e])
        assert idx[e] == i_ref[0] and np.float32(dist[e]).view(np.uint32) == d_ref[0].view(np.uint32), e
        assert_bitexact(seg[e], segs[e][i_ref[0]], "closest segment")
        assert bool(inside[e]) == bool(cr.points_inside(pts[e:e + 1], segs[e])[0]), e
    dr = 6.0
    sensor = UndirectedRangeSensor(dr, 20)
    dirs = orc.ray_directions((1, 0), 360, 20).astype(np.float32)
    ranges, points = sensor.get_observations_batch(pts, scene, return_intersection_points=True)
    for e in range(E):
        r_ref, p_ref, _ = orc.sensor_observation(pts[e], dirs, segs[e], dr)
        assert_bitexact(ranges[e], r_ref.astype(np.float32), f"ranges of environment {e}")
        same_with_inf(points[e], p_ref, f"points of environment {e}")
    cfg = {"sim": {"coordinate_target": [6.0, 4.0], "agent_params": {"distance_range": dr, "angle_step": 20}}}
    vel = rng.normal(size=(E, 2)).astype(np.float32)
    obs = RangeSensingAgent(cfg).get_observations(pts, vel, scene)
    ref = np.concatenate([orc.agent_observations(pts[e:e + 1], vel[e:e + 1], dirs, segs[e], dr, [6.0, 4.0]) for e in range(E)])
    assert_bitexact(obs.numpy(), ref, "controller input rows, one scene per agent")
    with pytest.raises(ValueError):
        geo.closest_segment_index(pts[:5], scene)


def test_float64_positions_keep_float64_ranges_like_the_reference():
    """observations.py:50,67: numpy keeps a float64 position, the ranges are float64 norms of (float32 points -
    float64 position); the intersections come from the float32-rounded origins (JAX)."""
    from doper_b200.agents import UndirectedRangeSensor
    from doper_b200.sim.jax_geometry import range_sensor
    segs = syn.make_map(5, 67)
    rng = np.random.default_rng(3)
    lo, hi = segs.reshape(-1, 2).min(0), segs.reshape(-1, 2).max(0)
    pos = rng.uniform(lo - 1, hi + 1, (300, 2))  # float64, not representable in float32
    pos[:40] = segs[rng.integers(0, len(segs), 40), 0].astype(np.float64)  # on vertices: ties
    dr = 7.5
    dirs = orc.ray_directions((1, 0), 360, 20).astype(np.float32)
    ranges, points, idx = range_sensor(pos, dirs, segs, dr, want_points=True)
    assert ranges.dtype == np.float64
    differs = 0
    for k in range(len(pos)):
        r_ref, p_ref, i_ref = orc.sensor_observation(pos[k], dirs, segs, dr, positions_dtype=np.float64)
        assert r_ref.dtype == np.float64
        assert np.array_equal(ranges[k].view(np.uint64), r_ref.view(np.uint64)), k
        same_with_inf(points[k], p_ref, "points")
        assert np.array_equal(idx[k], i_ref), k
        r32, _, _ = orc.sensor_observation(pos[k].astype(np.float32), dirs, segs, dr)
        differs += int((r32.astype(np.float64) != r_ref).sum())
    assert differs > 0  # the float64 branch is not the float32 one
    sensor = UndirectedRangeSensor(dr, 20)
    r1 = sensor.get_observation((float(pos[0, 0]), float(pos[0, 1])), segs)  # python floats -> float64, like onp.array
    assert r1.dtype == np.float64 and np.array_equal(r1, ranges[0])
    r2 = sensor.get_observation(pos[0].astype(np.float32), segs)
    assert r2.dtype == np.float32


def test_queries_and_sensor_on_a_scene_larger_than_shared_memory():
    """20,001 segments (400 KB staged) do not fit shared memory: the kernels read the scene in place."""
    from doper_b200.agents import UndirectedRangeSensor
    from doper_b200.sim import jax_geometry as geo
    from oracle.cref import CRef
    segs = syn.make_map(9, 6667)
    assert len(segs) == 20001
    rng = np.random.default_rng(1)
    lo, hi = segs.reshape(-1, 2).min(0), segs.reshape(-1, 2).max(0)
    pts = rng.uniform(lo, hi, (500, 2)).astype(np.float32)
    _, dist, idx = geo.closest_segment_index(pts, segs)
    i_ref, d_ref = CRef("portable").closest_segment(pts, segs)
    assert np.array_equal(idx, i_ref)
    assert_bitexact(dist, d_ref, "distances")
    dirs = orc.ray_directions((1, 0), 360, 20).astype(np.float32)
    sensor = UndirectedRangeSensor(50.0, 20)
    ranges = sensor.get_observations_batch(pts[:40], segs)
    r_ref, _, _ = orc.sensor_observations(pts[:40], dirs, segs, 50.0)
    assert_bitexact(ranges, r_ref, "ranges")
