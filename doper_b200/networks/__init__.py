from .controllers import *  # noqa: F401,F403
