"""Mirror of the reference ``doper/trainers/ball_trainer.py``: same class,
constructor, ``forward`` and ``optimize_parameters`` signatures, same control flow — 8 actions per
iteration, ``impulse.backward(dL/dv0)`` after every action, no gradient between actions
(ball_trainer.py:72), gradient clip 5, Adam.  What changes is where the data lives: controller,
observations, rollout and its gradient all stay on the GPU (the reference hops through host numpy
at every action, doper/utils/connectors.py), and the rollout is the sm_100a path.

Reference quirk kept on purpose: ``velocity_init = velocity_init`` (:73) — the final velocity of an
action is discarded and the next action starts again from the initial random velocity.

Multi-GPU: with ``torch.distributed`` initialised, every rank simulates its own
batch and the controller gradients + loss are all-reduced once per ``optimize_parameters``.
"""
__all__ = ["BallAttractorTrainer"]

import logging

import numpy as onp
import torch

from ..distributed import PolicyGradReducer
from ..networks import LinearReLUController
from ..sim.ball_sim import run_sim, vmapped_grad_and_value

logger = logging.getLogger(__name__)


class BallAttractorTrainer:
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.device = torch.device(self.config["train"]["device"] if "cuda" in str(self.config["train"]["device"])
                                   else "cuda")  # the rollout has one backend; "cpu" configs run on cuda:current
        self.constants = self.config["constants"]
        self.constants["volume"] = 4 * onp.pi * (self.constants["radius"] ** 3) / 3  # ball_trainer.py:22 (host double)
        self.constants["mass"] = self.constants["volume"] * self.constants["density"]
        self.controller = LinearReLUController(config["model"]).to(self.device)
        self.sim_time = config["sim"]["sim_time"]
        self.n_steps = config["sim"]["n_steps"]
        self.optimizer = torch.optim.Adam(self.controller.parameters())
        self.batch_size = self.config["train"]["batch_size"]
        self.num_actions = self.config["train"]["num_actions"]
        self._reducer = None

    def forward(self, observation, coordinate_init, velocity_init, scene):
        """Validation rollout of the FIRST ball (ball_trainer.py:31-47): ``(x_T, v_T, trajectory)``."""
        impulse = self.controller(observation.to(self.device))
        v0 = torch.as_tensor(onp.asarray(velocity_init.detach().cpu() if isinstance(velocity_init, torch.Tensor)
                                         else velocity_init), dtype=torch.float32, device=self.device)
        x0 = torch.as_tensor(onp.asarray(coordinate_init.detach().cpu() if isinstance(coordinate_init, torch.Tensor)
                                         else coordinate_init), dtype=torch.float32, device=self.device)
        return run_sim(
            self.sim_time, self.n_steps, scene.jax_scene, x0[0],
            (v0 + impulse.detach() / self.constants["mass"])[0],
            onp.array(self.config["sim"]["attractor_coordinate"]), self.constants,
        )

    def optimize_parameters(self, agent, scene_handler, grad_clip=5, generator=None):
        coordinate_init = scene_handler.get_init_state(self.batch_size)
        self.optimizer.zero_grad()
        if generator is None:
            velocity_init = torch.as_tensor(onp.random.uniform(-1, 2, (self.batch_size, 2)), dtype=torch.float32,
                                            device=self.device)  # ball_trainer.py:52
        else:
            velocity_init = -1 + 3 * torch.rand((self.batch_size, 2), device=self.device, generator=generator)
        coordinate_init = torch.as_tensor(coordinate_init, dtype=torch.float32, device=self.device)
        loss_val = None
        for action_id in range(self.num_actions):
            observation = agent.get_observations(coordinate_init, velocity_init, scene_handler)
            impulse = self.controller(observation.to(self.device))
            (loss_val, (coord, velocity)), v_grad = vmapped_grad_and_value(
                self.sim_time, self.n_steps, scene_handler.jax_scene, coordinate_init,
                (velocity_init + impulse / self.constants["mass"]).detach(),
                onp.array(self.config["sim"]["coordinate_target"]),
                onp.array(self.config["sim"]["attractor_coordinate"]),
                self.constants,
            )
            impulse.backward(v_grad)  # ball_trainer.py:70-71: dL/dv0 handed to the controller as dL/dimpulse
            coordinate_init = coord
            velocity_init = velocity_init  # sic (ball_trainer.py:73)
        if torch.distributed.is_available() and torch.distributed.is_initialized() and \
                torch.distributed.get_world_size() > 1:
            if self._reducer is None:
                self._reducer = PolicyGradReducer(self.controller.parameters(), self.device)
            loss_val = self._reducer.reduce(loss_val)
        torch.nn.utils.clip_grad_norm_(self.controller.parameters(), grad_clip)
        self.optimizer.step()
        return loss_val
