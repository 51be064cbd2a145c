"""Mirror of the reference ``doper/agents`` package (observation path)."""
from .observations import Sensor, SimpleRangeSensor, UndirectedRangeSensor  # noqa: F401
from .range_sensor_agent import RangeSensingAgent  # noqa: F401

__all__ = ["RangeSensingAgent"]
