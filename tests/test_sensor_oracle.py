"""CPU tests of the range-sensor oracle (oracle/ray_np.py): hand-computed cases and float64 geometry.
The reference has no tests or recorded outputs for this path: parity is pinned by this restatement
of doper/sim/jax_geometry.py:180-261, doper/agents/observations.py:29-110 and
doper/agents/range_sensor_agent.py:17-41 only."""
import numpy as np

from doper_b200 import synthetic as syn
from oracle import ray_np as orc


def test_known_intersections():
    seg = np.array([[[2.0, -1.0], [2.0, 1.0]],      # vertical wall at x = 2
                    [[-3.0, -1.0], [-3.0, 1.0]],    # behind the ray
                    [[0.0, 5.0], [4.0, 5.0]],       # parallel to the ray
                    [[1.0, 1.0], [1.0, 3.0]]],      # beside the ray (t2 < 0)
                   np.float32)
    pts = orc.batch_line_ray_intersection_point([[0.0, 0.0]], [[3.0, 0.0]], seg)  # unnormalised direction
    assert pts.shape == (1, 4, 2) and pts.dtype == np.float32
    assert np.array_equal(pts[0, 0], np.array([2.0, 0.0], np.float32))
    assert np.isinf(pts[0, 1:]).all()
    # endpoint hits count (t2 == 0 and t2 == 1 are inside the closed interval, jax_geometry.py:205)
    pts = orc.batch_line_ray_intersection_point([[0.0, 1.0], [0.0, -1.0]], [[1.0, 0.0], [1.0, 0.0]], seg[:1])
    assert np.array_equal(pts[:, 0], np.array([[2.0, 1.0], [2.0, -1.0]], np.float32))


def test_sensor_observation_matches_float64_geometry():
    segs = syn.make_map(7, 40)
    rng = np.random.default_rng(0)
    dirs = orc.ray_directions((1, 0), 360, 20)
    assert dirs.shape in ((18, 2), (19, 2)) and dirs.dtype == np.float64
    pos = rng.uniform([-10, 1], [10, 19], (50, 2)).astype(np.float32)
    ranges, points, idx = orc.sensor_observations(pos, dirs, segs, 1000.0)
    assert ranges.shape == (50, len(dirs)) and points.shape == (50, len(dirs), 2)
    # float64 ray casting
    s0, s1 = segs[:, 0].astype(np.float64), segs[:, 1].astype(np.float64)
    for i in range(50):
        o = pos[i].astype(np.float64)
        for r in range(len(dirs)):
            d = dirs[r].astype(np.float32).astype(np.float64)
            d = d / np.linalg.norm(d)
            v1, v2, v3 = o - s0, s1 - s0, np.array([-d[1], d[0]])
            den = v2 @ v3
            with np.errstate(all="ignore"):
                t1 = (v2[:, 0] * v1[:, 1] - v2[:, 1] * v1[:, 0]) / den
                t2 = (v1 @ v3) / den
            ok = (den != 0) & (t1 >= 0) & (t2 >= 0) & (t2 <= 1)
            best = t1[ok].min() if ok.any() else np.inf
            got = ranges[i, r]
            if np.isfinite(best) and best < 1000:
                assert abs(got - best) <= 1e-4 * max(1.0, best), (i, r, got, best)
            else:
                assert got == np.float32(1000.0) and np.isinf(points[i, r]).all()


def test_agent_observation_layout():
    segs = syn.make_map(7, 20)
    pos = np.array([[1.0, 2.0], [-4.0, 7.5]], np.float32)
    vel = np.array([[0.5, -1.0], [3.0, 2.0]], np.float32)
    dirs = orc.ray_directions((1, 0), 360, 20).astype(np.float32)
    obs = orc.agent_observations(pos, vel, dirs, segs, 1000.0, [6.0, 4.0])
    assert obs.shape == (2, 7 + len(dirs)) and obs.dtype == np.float32
    d = pos.astype(np.float64) - np.array([6.0, 4.0])
    assert np.array_equal(obs[:, 0], np.linalg.norm(d, axis=1).astype(np.float32))
    assert np.array_equal(obs[:, 3:5], pos) and np.array_equal(obs[:, 5:7], vel)
    assert (obs[:, 7:] > 0).all() and (obs[:, 7:] <= 1).all()
