"""Scene ingest: SVG subset reader, outline ordering, ear clipping.
Host-only code; the reference (doper/utils/assets.py) has no tests and depends on svgpathtools /
tripy, which are absent: what is checked here are the invariants the kernels rely on."""
import numpy as np
import pytest

from doper_b200.utils import assets as A
from oracle import ball_sim_np as onp_oracle


def poly_area(pts):
    p = np.asarray(pts, np.float64)
    x, y = p[:, 0], p[:, 1]
    return 0.5 * float(np.sum(x * np.roll(y, -1) - np.roll(x, -1) * y))


def tri_segments_area(seg):
    a, b, c = seg[0, 0], seg[1, 0], seg[2, 0]
    return 0.5 * float((b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]))


FIGURES_PX = [  # SVG pixel coordinates (y down), px_per_meter = 10, canvas 300 x 200
    [(20, 20), (120, 30), (130, 110), (70, 150), (15, 90)],                 # convex pentagon
    [(160, 20), (280, 20), (280, 180), (160, 180), (160, 120), (230, 120), (230, 70), (160, 70)],  # concave "C"
    [(30, 170), (60, 160), (50, 190)],                                       # triangle
]


def test_path_data_subset():
    assert A._path_lines("M 0,0 L 10,0 L 10,10 Z") == [(0j, 10 + 0j), (10 + 0j, 10 + 10j), (10 + 10j, 0j)]
    rel = A._path_lines("m 5,5 l 10,0 v 10 h -10 z")
    assert rel[0] == (5 + 5j, 15 + 5j) and rel[-1] == (5 + 15j, 5 + 5j) and len(rel) == 4
    with pytest.raises(ValueError):
        A._path_lines("M 0,0 C 1,1 2,2 3,3")


def test_svg_round_trip_invariants(tmp_path):
    f = tmp_path / "map.svg"
    A.write_svg(f, FIGURES_PX, 300, 200)
    scene = A.get_svg_scene(f, px_per_meter=10)
    seg = np.asarray(scene.segments)
    assert seg.dtype == np.float32 and seg.shape[1:] == (2, 2) and seg.shape[0] % 3 == 0
    n_tri = sum(len(fig) - 2 for fig in FIGURES_PX)
    assert seg.shape[0] == 3 * n_tri == 3 * len(scene.polygons)
    # every triangle: closed chain (p0,p1),(p1,p2),(p2,p0), counter-clockwise (positive area)
    for t in range(0, len(seg), 3):
        tri = seg[t:t + 3]
        assert np.array_equal(tri[0, 1], tri[1, 0]) and np.array_equal(tri[1, 1], tri[2, 0]) and \
            np.array_equal(tri[2, 1], tri[0, 0])
        assert tri_segments_area(tri) > 0
    # the triangles tile the figures: same total area (metres, y flipped)
    want = sum(abs(poly_area([(x / 10, (200 - y) / 10) for x, y in fig])) for fig in FIGURES_PX)
    got = sum(tri_segments_area(seg[t:t + 3]) for t in range(0, len(seg), 3))
    assert abs(got - want) < 1e-3 * want
    # and the inner check (the reference arithmetic, oracle) agrees with the figures at sample points
    inside_pt = np.array([[7.0, 12.0], [25.5, 16.0], [4.5, 2.7]], np.float32)   # inside pentagon, C, triangle
    outside_pt = np.array([[20.0, 10.5], [14.5, 10.0], [0.5, 19.5]], np.float32)  # C's notch, gap, corner
    assert onp_oracle.points_inside_any_polygon(inside_pt, seg).all()
    assert not onp_oracle.points_inside_any_polygon(outside_pt, seg).any()


def test_sort_segments_chains_a_shuffled_outline():
    """The reference's start heuristic swaps only the x coordinates of its two candidates
    (assets.py:48-51), so the walking direction it picks depends on how the segments happen to be
    stored; ear clipping re-orients clockwise outlines afterwards (tripy.earclip), which is why scenes
    still come out counter-clockwise (round-trip test above).  What sort_segments itself guarantees
    is a head-to-tail chain covering every segment once."""
    pts = np.array([(0, 0), (4, 0), (5, 3), (2, 5), (-1, 2)], np.float64)
    segs = np.stack([pts, np.roll(pts, -1, axis=0)], axis=1)
    rng = np.random.default_rng(0)
    for _ in range(20):
        perm = rng.permutation(len(segs))
        shuffled = segs[perm].copy()
        flip = rng.random(len(segs)) < 0.5
        shuffled[flip] = shuffled[flip][:, ::-1]
        for orientation in ("counterclockwise", "clockwise"):
            out = A.sort_segments(shuffled, orientation)
            assert np.allclose(out[:, 1], np.roll(out[:, 0], -1, axis=0))  # head-to-tail chain
            assert abs(abs(poly_area(out[:, 0])) - abs(poly_area(pts))) < 1e-9  # same outline
        tris = A.earclip([s[0] for s in A.sort_segments(shuffled, "counterclockwise")])
        assert len(tris) == 3 and all(poly_area(t) > 0 for t in tris)
    # stored left-to-right along the lower edge, the documented intent holds: counter-clockwise
    assert poly_area(A.sort_segments(segs, "counterclockwise")[:, 0]) > 0
    assert poly_area(A.sort_segments(segs, "clockwise")[:, 0]) < 0
    with pytest.raises(ValueError):
        A.sort_segments(segs, "sideways")


def test_rotate_transform_and_cache(tmp_path):
    f = tmp_path / "rot.svg"
    A.write_svg(f, [FIGURES_PX[2]], 300, 200, transforms=["rotate(90, 45, 175)"])
    a = np.asarray(A.get_svg_scene(f, px_per_meter=10).segments)
    A.write_svg(f, [FIGURES_PX[2]], 300, 200)
    b = np.asarray(A.get_svg_scene(f, px_per_meter=10).segments)
    assert a.shape == b.shape == (3, 2, 2) and not np.allclose(a, b)
    c = np.array([4.5, 2.5])
    assert np.allclose(np.linalg.norm(a.reshape(-1, 2) - c, axis=1), np.linalg.norm(b.reshape(-1, 2) - c, axis=1),
                       atol=1e-5)  # a rotation about (cx, h - cy) / px_per_meter keeps distances to the centre
    A.save_scene(tmp_path / "scene.npy", A.get_svg_scene(f, px_per_meter=10))
    assert np.array_equal(np.asarray(A.load_scene(tmp_path / "scene.npy").segments), b)
