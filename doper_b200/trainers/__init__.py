from .ball_trainer import *  # noqa: F401,F403
