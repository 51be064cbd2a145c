"""No-GPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol the
header declares, and the host-side mirror keeps the reference's names and signatures."""
import ctypes
import inspect
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _header_symbols():
    text = (ROOT / "include" / "doper_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(doper_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_header_symbols():
    from doper_b200 import _abi
    syms = _header_symbols()
    assert len(syms) >= 12
    for s in syms:
        assert hasattr(_abi.lib, s), f"{s} declared in include/doper_b200.h but not exported"
    assert set(_abi.EXPORTS) == set(syms)
    assert _abi.lib.doper_abi_version() == _abi.ABI_VERSION


def test_error_strings_and_size_queries_without_gpu():
    from doper_b200 import _abi
    assert _abi.error_string(0) == "ok"
    for code in range(1, 8):
        assert "unknown" not in _abi.error_string(code)
    assert _abi.lib.doper_scene_workspace_bytes(1, 201) % 256 == 0
    assert _abi.lib.doper_scene_workspace_bytes(0, 201) == 0
    d = _abi.RolloutDesc()
    d.B, d.S, d.n_steps, d.ckpt_every = 4096, 201, 500, 32
    n = _abi.lib.doper_rollout_workspace_bytes(ctypes.byref(d))
    assert n >= 16 * 4096 * 16 + 500 * 4096 * 4
    d.lanes_per_ball = 3  # not a power of two
    assert _abi.lib.doper_rollout_workspace_bytes(ctypes.byref(d)) == 0


def test_argument_validation_happens_before_any_launch():
    """Bad arguments are rejected with distinct codes and never reach CUDA (safe without a GPU)."""
    from doper_b200 import _abi
    L = _abi.lib
    assert L.doper_scene_prepare(None, 0, 201, None, 0.1, None) == 1  # BAD_SHAPE
    assert L.doper_scene_prepare(None, 1, 201, None, 0.1, None) == 2  # NULL_PTR
    assert L.doper_scene_prepare(None, 1, 201, 24, 0.1, 256) == 3  # MISALIGNED (segments % 16)
    assert L.doper_points_inside(None, 4, 200, 8, 16, 1) == 7  # S % 3 != 0 -> UNSUPPORTED
    d = _abi.RolloutDesc()
    d.B, d.S, d.n_steps = 0, 201, 10
    assert L.doper_rollout_fwd(None, ctypes.byref(d), 256, 8, 8, 8, 8, None, None, None, None) == 1
    d.B = 4
    assert L.doper_rollout_fwd(None, ctypes.byref(d), None, 8, 8, 8, 8, None, None, None, None) == 2
    assert L.doper_rollout_fwd(None, ctypes.byref(d), 256, 4, 8, 8, 8, None, None, None, None) == 3
    d.S = 70000  # beyond the 16-bit candidate-list indices
    assert L.doper_rollout_fwd(None, ctypes.byref(d), 256, 8, 8, 8, 8, None, None, None, None) == 7
    assert L.doper_rollout_workspace_bytes(ctypes.byref(d)) == 0
    assert L.doper_range_sensor(None, 4, 0, 201, 8, 8, 256, 1000.0, 8, None, None) == 1   # R <= 0
    assert L.doper_ray_segment_points(None, 4, 201, 8, 8, 16, 4) == 3                      # points % 8


def test_mirror_keeps_reference_signatures():
    """Positional parameter names of the reference (ball_sim.py:212-220,256-265,289-298;
    jax_geometry.py:98,160)."""
    from doper_b200.sim import ball_sim, jax_geometry
    ref_run_sim = ["sim_time", "n_steps", "scene", "coordinate_init", "velocity_init", "attractor", "constants"]
    ref_loss = ["sim_time", "n_steps", "scene", "coordinate_init", "velocity_init", "target_coordinate",
                "attractor", "constants"]
    assert list(inspect.signature(ball_sim.run_sim).parameters)[:7] == ref_run_sim
    for fn in (ball_sim.compute_loss, ball_sim.vmapped_loss, ball_sim.reduce_loss, ball_sim.vmapped_grad_and_value):
        assert list(inspect.signature(fn).parameters)[:8] == ref_loss
    assert list(inspect.signature(jax_geometry.find_closest_segment_to_point).parameters) == ["point", "segments_batch"]
    assert list(inspect.signature(jax_geometry.if_points_inside_any_polygon).parameters) == ["points", "scene"]
    assert ball_sim.BallState._fields == ("coordinate", "velocity", "acceleration")
    sc = jax_geometry.create_scene([jax_geometry.create_polygon(np.zeros((3, 2, 2), np.float32))] * 2)
    segments, polygons = sc  # namedtuple-style unpacking like JaxScene(segments, polygons)
    assert segments.shape == (6, 2, 2) and len(polygons) == 2 and len(sc.polygons) == 2


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from doper_b200.sim import ball_sim
    from doper_b200 import synthetic as syn
    w = syn.make_workload("c1")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ball_sim.vmapped_grad_and_value(0.3, 300, w["segments"], w["x0"], w["v0"], w["target"], w["attractor"],
                                        w["constants"])


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under doper_b200/ may import it."""
    for py in (ROOT / "doper_b200").rglob("*.py"):
        text = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), py
