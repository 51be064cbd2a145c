"""Mirror of the reference ``doper/scenes/single_scene.py``.

``SingleScene.get_init_state`` is the second caller of the hot-path geometry functions
(single_scene.py:44-47): it rejection-samples start positions that are outside every triangle and
farther than ``radius + radius / 2`` from every segment.  Here the whole loop runs on the device:
proposals are drawn by a torch CUDA generator, the inner check and the closest-segment distance
are the sm_100a kernels, and only the "all accepted?" flag crosses to the host per iteration (the
reference converts numpy <-> JAX arrays every iteration).  Sampling boxes follow the reference:
``[min - 1, max + 1]`` for the first proposal (:39-41), ``[min, max]`` for resamples (:58-60).
The reference's sampler is unseeded numpy; pass ``generator=`` for reproducible draws.
"""
__all__ = ["SingleScene"]

import logging

import numpy as onp
import torch

from .._device import require_cuda
from ..sim.jax_geometry import JaxScene, _as_scene, find_closest_segment_to_points_batch, \
    if_points_inside_any_polygon

logger = logging.getLogger(__name__)


class SingleScene:
    def __init__(self, config: dict, jax_scene: JaxScene = None):
        self.config = config
        if jax_scene is None:
            from ..utils.assets import get_svg_scene
            jax_scene = get_svg_scene(
                config["sim"]["scene_params"]["svg_scene_path"],
                px_per_meter=config["sim"]["scene_params"]["px_per_meter"],
            )
        self.jax_scene = _as_scene(jax_scene)

    def get_init_state(self, batch_size: int, generator: torch.Generator = None) -> torch.Tensor:
        """-> CUDA tensor ``[batch_size, 2]`` float32 of admissible start coordinates."""
        device = require_cuda()
        seg = onp.asarray(self.jax_scene.segments.cpu() if isinstance(self.jax_scene.segments, torch.Tensor)
                          else self.jax_scene.segments)
        eps = self.config["constants"]["radius"] / 2
        limit = self.config["constants"]["radius"] + eps
        if self.jax_scene.per_env:
            return self._get_init_state_per_env(batch_size, seg, limit, device, generator)
        lo = torch.tensor([seg[:, :, 0].min(), seg[:, :, 1].min()], dtype=torch.float32, device=device)
        hi = torch.tensor([seg[:, :, 0].max(), seg[:, :, 1].max()], dtype=torch.float32, device=device)

        def draw(n, a, b):
            u = torch.rand((n, 2), dtype=torch.float32, device=device, generator=generator)
            return a + (b - a) * u

        out = torch.empty((batch_size, 2), dtype=torch.float32, device=device)
        remaining = torch.arange(batch_size, device=device)
        proposal = draw(batch_size, lo - 1, hi + 1)
        while True:
            is_inner = if_points_inside_any_polygon(proposal, self.jax_scene)
            _, distance = find_closest_segment_to_points_batch(proposal, self.jax_scene)
            acceptable = torch.logical_not(torch.logical_or(is_inner, distance <= limit))
            out[remaining[acceptable]] = proposal[acceptable]
            if bool(acceptable.all()):
                break
            logger.debug("Resampling starting position")
            remaining = remaining[torch.logical_not(acceptable)]
            proposal = draw(len(remaining), lo, hi)
        return out

    def _get_init_state_per_env(self, batch_size, seg, limit, device, generator):
        """One environment per ball ([E,S,2,2] scenes, BASELINE configs[2]): ball i is sampled in the box of ITS map
        and tested against ITS triangles and segments (per-environment kernels: point i <-> scene i).  Same
        acceptance rule and sampling boxes as single_scene.py:39-61, applied per environment; every iteration
        re-tests the whole batch (accepted balls keep their position), so point i always meets scene i."""
        E = self.jax_scene.n_scenes
        if batch_size != E:
            raise ValueError(f"per-environment scenes: batch_size {batch_size} != {E} environments")
        lo = torch.tensor(seg.reshape(E, -1, 2).min(axis=1), dtype=torch.float32, device=device)
        hi = torch.tensor(seg.reshape(E, -1, 2).max(axis=1), dtype=torch.float32, device=device)

        def draw(a, b):
            return a + (b - a) * torch.rand((E, 2), dtype=torch.float32, device=device, generator=generator)

        out = draw(lo - 1, hi + 1)
        while True:
            is_inner = if_points_inside_any_polygon(out, self.jax_scene)
            _, distance = find_closest_segment_to_points_batch(out, self.jax_scene)
            acceptable = torch.logical_not(torch.logical_or(is_inner, distance <= limit))
            if bool(acceptable.all()):
                return out
            logger.debug("Resampling starting position")
            out = torch.where(acceptable[:, None], out, draw(lo, hi))
