"""Drop-in mirror of the reference ``doper/agents/observations.py``: same classes, constructor
arguments and ``get_observation`` signatures / return nesting.  The ray / segment intersections,
the per-ray minimum and the range clamp run in one sm_100a kernel (``doper_range_sensor``); the
ray directions are built on the host exactly like the reference builds them (float64 numpy
``arange`` / ``cos`` / ``sin``, observations.py:50-58), so the number of rays matches too.

``get_observations_batch`` is the batched entry point the agent uses (the reference loops over
balls in Python, range_sensor_agent.py:36-39).
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import numpy as onp

from ..sim.jax_geometry import JaxScene, range_sensor

__all__ = ["Sensor", "SimpleRangeSensor", "UndirectedRangeSensor"]


class Sensor:
    """Dummy sensor class to use in typehints (observations.py:6-13)."""

    def get_observation(self) -> None:
        pass


class SimpleRangeSensor(Sensor):
    def __init__(self, distance_range: float, angle_range: float, angle_step: float) -> None:
        self._distance_range = distance_range
        self._angle_range_rad = angle_range / 180 * onp.pi
        self._angle_step_rad = angle_step / 180 * onp.pi

    def ray_directions(self, heading_direction) -> onp.ndarray:
        """observations.py:50-58 (float64, unnormalised-safe)."""
        heading_angle = onp.arctan2(heading_direction[1], heading_direction[0])
        ray_angles = onp.arange(
            heading_angle - self._angle_range_rad / 2,
            heading_angle + self._angle_range_rad / 2,
            self._angle_step_rad,
        )
        return onp.concatenate(
            [onp.cos(ray_angles)[..., onp.newaxis], onp.sin(ray_angles)[..., onp.newaxis]], axis=-1
        )

    def get_observations_batch(self, positions, heading_direction, scene: JaxScene,
                               return_intersection_points: Optional[bool] = False):
        """All positions in one launch: ``[N,2] -> ranges [N,R]`` (or ``(ranges, points [N,R,2])``)."""
        return self._batch(positions, heading_direction, scene, return_intersection_points)

    def _batch(self, positions, heading_direction, scene, return_intersection_points):
        dirs = self.ray_directions(heading_direction).astype(onp.float32)
        if return_intersection_points:
            ranges, points, _ = range_sensor(positions, dirs, scene, self._distance_range, want_points=True)
            return ranges, points
        return range_sensor(positions, dirs, scene, self._distance_range)

    def get_observation(
        self,
        position: Union[onp.ndarray, Tuple[float, float]],
        heading_direction: Union[onp.ndarray, Tuple[float, float]],
        scene: JaxScene,
        return_intersection_points: Optional[bool] = False,
    ) -> Union[onp.ndarray, Tuple[onp.ndarray, onp.ndarray]]:
        """observations.py:29-79 for one position: ranges ``[R]`` or ``(ranges, points [R,2])``."""
        # observations.py:50 `onp.array(position)`: a float32 array stays float32, anything else (python floats,
        # float64 arrays: e.g. the sampler's proposals) is float64 and the ranges come out in float64 (:67)
        position = onp.array(position)
        position = (position if position.dtype == onp.float32 else position.astype(onp.float64)).reshape(1, 2)
        out = self._batch(position, heading_direction, scene, return_intersection_points)
        if return_intersection_points:
            return out[0][0], out[1][0]
        return out[0]


class UndirectedRangeSensor(SimpleRangeSensor):
    def __init__(self, distance_range: float, angle_step: float) -> None:
        super().__init__(distance_range=distance_range, angle_range=360, angle_step=angle_step)

    def get_observations_batch(self, positions, scene: JaxScene, return_intersection_points: Optional[bool] = False):
        return self._batch(positions, (1, 0), scene, return_intersection_points)

    def get_observation(
        self,
        position: Union[onp.ndarray, Tuple[float, float]],
        scene: JaxScene,
        return_intersection_points: Optional[bool] = False,
    ) -> Union[onp.ndarray, Tuple[onp.ndarray, onp.ndarray]]:
        """observations.py:82-110."""
        return super().get_observation(position, (1, 0), scene, return_intersection_points)
