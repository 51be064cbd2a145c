"""Mirror of the reference ``doper/scenes/multiple_scenes.py``: every SVG matching the glob becomes
a ``SingleScene``; each call picks one at random and exposes its scene as ``.jax_scene``
(multiple_scenes.py:14-27).  Unlike the reference, each single scene gets its OWN copy of the
nested config (the reference shares the nested dict, so all its copies end up pointing at the last
path, :19-21 — harmless there because the scene is loaded in the constructor)."""
__all__ = ["MultipleScenes"]

import copy
import logging
from glob import glob

import numpy as onp

from .single_scene import SingleScene

logger = logging.getLogger(__name__)


class MultipleScenes:
    def __init__(self, config: dict, scenes=None):
        self.config = config
        self.single_scenes = []
        if scenes is not None:  # pre-built scenes (synthetic maps: the reference's SVG assets are git-lfs stubs)
            self.single_scenes = [SingleScene(config, jax_scene=s) for s in scenes]
        else:
            for scene_path in sorted(glob(config["sim"]["scene_params"]["svg_scene_path"])):
                single_config = copy.deepcopy(config)
                single_config["sim"]["scene_params"]["svg_scene_path"] = scene_path
                self.single_scenes.append(SingleScene(single_config))
        if not self.single_scenes:
            raise ValueError("MultipleScenes: no scene found")
        self.jax_scene = self.single_scenes[0].jax_scene

    def get_init_state(self, batch_size: int, generator=None):
        single_scene = self.single_scenes[int(onp.random.randint(len(self.single_scenes)))]
        self.jax_scene = single_scene.jax_scene
        return single_scene.get_init_state(batch_size, generator=generator)
