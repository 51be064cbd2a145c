"""ctypes binding of ``libdoper_b200.so`` (include/doper_b200.h).

The product has exactly one backend: the sm_100a CUDA library.  Importing this module fails
loudly (``ImportError``) when the library has not been built — there is no CPU or eager
fallback anywhere in ``doper_b200``.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_size_t, c_uint8, c_void_p
from pathlib import Path

import os

# DOPER_B200_LIB points the binding at another build of the same library (kernel A/B tuning)
LIB_PATH = Path(os.environ.get("DOPER_B200_LIB") or (Path(__file__).resolve().parent / "_lib" / "libdoper_b200.so"))

ABI_VERSION = 3

EXPORTS = (
    "doper_abi_version", "doper_fp_policy", "doper_error_string", "doper_scene_workspace_bytes", "doper_scene_prepare",
    "doper_closest_segment", "doper_points_inside", "doper_rollout_ckpt_every", "doper_rollout_log_ball_major",
    "doper_rollout_workspace_bytes", "doper_rollout_bwd_scratch_bytes", "doper_rollout_plan_name",
    "doper_rollout_fwd", "doper_rollout_bwd", "doper_l1_loss", "doper_value_and_grad",
    "doper_fp32_probe",
    "doper_ray_segment_points", "doper_range_sensor", "doper_agent_observations",
    "doper_closest_segment_env", "doper_points_inside_env", "doper_range_sensor_env", "doper_range_sensor_f64",
    "doper_agent_observations_env",
    "doper_controller_scratch_bytes", "doper_controller_fwd", "doper_controller_bwd", "doper_adam_clip_step",
    "doper_peer_buffer_bytes", "doper_peer_alloc", "doper_peer_open", "doper_peer_close", "doper_peer_free",
    "doper_allreduce_oneshot",
)


class RolloutDesc(ctypes.Structure):
    """``doper_rollout_desc`` (include/doper_b200.h)."""

    _fields_ = [
        ("B", c_int64),
        ("S", c_int32),
        ("per_env", c_int32),
        ("n_steps", c_int32),
        ("ckpt_every", c_int32),
        ("dt", c_float),
        ("consts", c_float * 5),
        ("attractor", c_float * 2),
        ("lanes_per_ball", c_int32),
        ("flags", c_int32),
        ("hit_capacity", c_int64),
    ]


class DoperError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = code
        super().__init__(f"{where}: doper_b200 error {code}: {error_string(code)}")


def _load() -> ctypes.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"doper_b200: CUDA library not built ({LIB_PATH}). Run `python -m doper_b200.build` "
            "(or __graft_entry__.build()). There is no CPU fallback."
        )
    try:
        lib = ctypes.CDLL(str(LIB_PATH))
    except OSError as e:  # pragma: no cover - depends on the box
        raise ImportError(f"doper_b200: cannot load {LIB_PATH}: {e}") from e
    missing = [s for s in EXPORTS if not hasattr(lib, s)]
    if missing:
        raise ImportError(f"doper_b200: {LIB_PATH} lacks symbols {missing}")
    vp, fp = c_void_p, c_void_p  # device pointers are passed as integers
    lib.doper_abi_version.restype = ctypes.c_int
    lib.doper_fp_policy.restype = ctypes.c_int
    lib.doper_error_string.restype = c_char_p
    lib.doper_error_string.argtypes = [ctypes.c_int]
    lib.doper_scene_workspace_bytes.restype = c_size_t
    lib.doper_scene_workspace_bytes.argtypes = [c_int64, c_int32]
    lib.doper_scene_prepare.argtypes = [vp, c_int64, c_int32, fp, c_float, vp]
    lib.doper_closest_segment.argtypes = [vp, c_int64, c_int32, fp, fp, vp, vp, fp, fp]
    lib.doper_points_inside.argtypes = [vp, c_int64, c_int32, fp, fp, vp]
    lib.doper_closest_segment_env.argtypes = lib.doper_closest_segment.argtypes
    lib.doper_points_inside_env.argtypes = lib.doper_points_inside.argtypes
    lib.doper_rollout_ckpt_every.restype = c_int32
    lib.doper_rollout_ckpt_every.argtypes = [POINTER(RolloutDesc)]
    lib.doper_rollout_log_ball_major.restype = c_int32
    lib.doper_rollout_log_ball_major.argtypes = [POINTER(RolloutDesc)]
    lib.doper_rollout_plan_name.restype = c_char_p
    lib.doper_rollout_plan_name.argtypes = [POINTER(RolloutDesc), c_int32]
    lib.doper_rollout_workspace_bytes.restype = c_size_t
    lib.doper_rollout_workspace_bytes.argtypes = [POINTER(RolloutDesc)]
    lib.doper_rollout_bwd_scratch_bytes.restype = c_size_t
    lib.doper_rollout_bwd_scratch_bytes.argtypes = [POINTER(RolloutDesc)]
    lib.doper_rollout_fwd.argtypes = [vp, POINTER(RolloutDesc), vp, fp, fp, fp, fp, fp, fp, fp, vp]
    lib.doper_rollout_bwd.argtypes = [vp, POINTER(RolloutDesc), vp, vp, vp, fp, fp, fp, fp]
    lib.doper_l1_loss.argtypes = [vp, c_int64, fp, POINTER(c_float), fp, fp, fp]
    lib.doper_value_and_grad.argtypes = [vp, POINTER(RolloutDesc), vp, fp, fp, POINTER(c_float), vp, vp, fp,
                                         fp, fp, fp, fp, fp, fp]
    lib.doper_fp32_probe.argtypes = [vp, c_int32, c_int32, c_int32, c_int32, fp]
    lib.doper_ray_segment_points.argtypes = [vp, c_int64, c_int32, fp, fp, fp, fp]
    lib.doper_range_sensor.argtypes = [vp, c_int64, c_int32, c_int32, fp, fp, vp, c_float, fp, fp, vp]
    lib.doper_agent_observations.argtypes = [vp, c_int64, c_int32, c_int32, fp, fp, fp, vp, c_float,
                                             POINTER(ctypes.c_double), fp]
    lib.doper_range_sensor_env.argtypes = lib.doper_range_sensor.argtypes
    lib.doper_agent_observations_env.argtypes = lib.doper_agent_observations.argtypes
    i4 = POINTER(c_int32)
    lib.doper_controller_scratch_bytes.restype = c_size_t
    lib.doper_controller_scratch_bytes.argtypes = [i4, c_int64]
    lib.doper_controller_fwd.argtypes = [vp, i4, c_int64, fp, fp, fp, c_float, fp, fp]
    lib.doper_controller_bwd.argtypes = [vp, i4, c_int64, fp, fp, fp, c_float, vp, fp]
    lib.doper_adam_clip_step.argtypes = [vp, c_int32, fp, fp, fp, fp, fp, c_float, c_float, c_float, c_float, c_float, fp]
    lib.doper_range_sensor_f64.argtypes = [vp, c_int64, c_int32, c_int32, c_int32, vp, fp, vp, c_float, vp, fp, vp]
    lib.doper_peer_buffer_bytes.restype = c_size_t
    lib.doper_peer_buffer_bytes.argtypes = [c_int32]
    lib.doper_peer_alloc.argtypes = [c_int32, POINTER(c_void_p), ctypes.c_char_p]
    lib.doper_peer_open.argtypes = [ctypes.c_char_p, POINTER(c_void_p)]
    lib.doper_peer_close.argtypes = [c_void_p]
    lib.doper_peer_free.argtypes = [c_void_p]
    lib.doper_allreduce_oneshot.argtypes = [vp, c_int32, c_int32, POINTER(c_void_p), c_int32, fp, fp, vp, ctypes.c_double]
    for name in EXPORTS:
        if name not in ("doper_controller_scratch_bytes", "doper_peer_buffer_bytes", "doper_error_string", "doper_scene_workspace_bytes",
                        "doper_rollout_workspace_bytes", "doper_rollout_bwd_scratch_bytes", "doper_abi_version", "doper_rollout_ckpt_every",
                        "doper_rollout_log_ball_major", "doper_rollout_plan_name"):
            getattr(lib, name).restype = ctypes.c_int
    if lib.doper_abi_version() != ABI_VERSION:
        raise ImportError(f"doper_b200: ABI version mismatch: library {lib.doper_abi_version()}, "
                          f"binding {ABI_VERSION}")
    return lib


lib = _load()


def error_string(code: int) -> str:
    return lib.doper_error_string(int(code)).decode()


def check(code: int, where: str) -> None:
    if code != 0:
        raise DoperError(code, where)
