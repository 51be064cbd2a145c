from . import ball_sim, jax_geometry  # noqa: F401
