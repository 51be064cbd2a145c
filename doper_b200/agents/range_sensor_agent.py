"""Drop-in mirror of the reference ``doper/agents/range_sensor_agent.py``.

``RangeSensingAgent.get_observations`` returns the controller input ``[B, 7 + R]`` as a torch
tensor like the reference (range_sensor_agent.py:17-41), but builds it in ONE kernel launch
(``doper_agent_observations``) instead of a Python loop over balls with a host round trip per
ball: distance and direction to the target (float64 arithmetic, cast once, like the reference's
numpy code), coordinates, velocities and the normalised ranges.  CUDA inputs stay on the device.
"""
from __future__ import annotations

import ctypes

import numpy as onp
import torch

from .._abi import check, lib
from .._device import as_f32_device, is_device_tensor, ptr, require_cuda, stream_ptr
from ..sim.jax_geometry import _as_scene
from .observations import UndirectedRangeSensor

__all__ = ["RangeSensingAgent"]


class RangeSensingAgent:
    def __init__(self, config: dict):
        self.config = config
        self.sensor = UndirectedRangeSensor(
            self.config["sim"]["agent_params"]["distance_range"],
            self.config["sim"]["agent_params"]["angle_step"],
        )
        self._dirs = None

    def get_observations(self, coordinate_init, velocity_init, scene_handler: object) -> torch.Tensor:
        """-> torch.Tensor ``[batch_size, observation_len]`` (CPU for host inputs, like the reference;
        CUDA when the coordinates are CUDA tensors).  ``scene_handler`` is anything with a
        ``jax_scene`` attribute (doper/scenes/*), or a scene itself."""
        device = require_cuda()
        scene = _as_scene(getattr(scene_handler, "jax_scene", scene_handler))
        _, ws = scene.prepared(device)
        on_device = is_device_tensor(coordinate_init)
        pos = as_f32_device(coordinate_init, device).reshape(-1, 2)
        vel = as_f32_device(velocity_init, device).reshape(-1, 2)
        if self._dirs is None or self._dirs.device != device:
            self._dirs = torch.as_tensor(self.sensor.ray_directions((1, 0)).astype(onp.float32), device=device)
        N, R, S = int(pos.shape[0]), int(self._dirs.shape[0]), scene.n_segments
        tgt = onp.array(self.config["sim"]["coordinate_target"], dtype=onp.float64).reshape(2)
        obs = torch.empty((N, 7 + R), dtype=torch.float32, device=device)
        if scene.per_env and N != scene.n_scenes:
            raise ValueError(f"per-environment scenes: {N} agents for {scene.n_scenes} scenes")
        fn = lib.doper_agent_observations_env if scene.per_env else lib.doper_agent_observations  # ball i looks at scene i
        check(fn(stream_ptr(), N, R, S, ptr(pos), ptr(vel), ptr(self._dirs), ptr(ws),
                                           float(self.config["sim"]["agent_params"]["distance_range"]),
                                           (ctypes.c_double * 2)(tgt[0], tgt[1]), ptr(obs)),
              "doper_agent_observations")
        return obs if on_device else obs.cpu()
