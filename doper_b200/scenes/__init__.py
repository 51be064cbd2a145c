from .single_scene import *  # noqa: F401,F403
from .multiple_scenes import *  # noqa: F401,F403
