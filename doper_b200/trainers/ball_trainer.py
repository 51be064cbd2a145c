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

``graphed_iteration`` captures the whole action loop of one iteration — 8 x
(observation kernel, controller forward, rollout forward + loss + backward, controller backward) +
gradient clip + Adam — into ONE CUDA graph: at the reference's batch size (16 balls) an iteration is
launch-bound (~200 small launches), and a graph replay removes the per-launch host cost.
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
        # capturable: the step counter lives on the device, so the optimiser can run inside a CUDA graph
        self.optimizer = torch.optim.Adam(self.controller.parameters(), capturable=True)
        self.batch_size = self.config["train"]["batch_size"]
        self.num_actions = self.config["train"]["num_actions"]
        self._reducer = None
        self.reduce_transport = "auto"  # distributed.PolicyGradReducer: "auto" | "nvlink" | "collective"
        self.data_parallel = True  # False: never exchange, even inside an initialised process group
        self._fused = None  # trainers/fused.py: the action loop on the sm_100a kernels only (created on first use)

    def fused_loop(self, grad_clip=5):
        """The action loop without torch operators (controller, gradient clip and Adam as kernels of this library).
        From the first call on the controller's parameters are views into one flat vector and the fused Adam state
        replaces ``self.optimizer``'s: do not mix with ``optimize_parameters`` afterwards."""
        if self._fused is None:
            from .fused import FusedActionLoop
            self._fused = FusedActionLoop(self, grad_clip)
        return self._fused

    def optimize_parameters_fused(self, agent, scene_handler, grad_clip=5, generator=None):
        """``optimize_parameters`` on the fused kernels (same control flow, same update rule)."""
        coordinate_init = torch.as_tensor(scene_handler.get_init_state(self.batch_size), dtype=torch.float32, device=self.device)
        if generator is None:
            velocity_init = torch.as_tensor(onp.random.uniform(-1, 2, (self.batch_size, 2)), dtype=torch.float32, device=self.device)
        else:
            velocity_init = -1 + 3 * torch.rand((self.batch_size, 2), device=self.device, generator=generator)
        return self.fused_loop(grad_clip).run(agent, scene_handler, coordinate_init, velocity_init)

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
        loss_val = self._exchange(loss_val)
        torch.nn.utils.clip_grad_norm_(self.controller.parameters(), grad_clip)
        self.optimizer.step()
        return loss_val

    def _exchange(self, loss_val):
        """Data-parallel ranks (one process per GPU, each with its own balls): sum the controller
        gradients and the loss over ranks — one small kernel over NVLink peer memory
        (``distributed.PeerAllReducer``), capturable in the iteration's CUDA graph."""
        if not (self.data_parallel and torch.distributed.is_available() and torch.distributed.is_initialized()
                and torch.distributed.get_world_size() > 1):
            return loss_val
        if self._reducer is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("BallAttractorTrainer: the gradient reducer must exist before graph capture")
            self._reducer = PolicyGradReducer(self.controller.parameters(), self.device, self.reduce_transport)
        return self._reducer.reduce(loss_val)

    # ------------------------------------------------------------------------------------------
    # one iteration as ONE CUDA graph
    # ------------------------------------------------------------------------------------------
    def _action_loop(self, agent, scene_handler, coordinate_init, velocity_init, grad_clip):
        """ball_trainer.py:53-75 on device tensors; no host synchronisation anywhere (capturable)."""
        loss_val = None
        for action_id in range(self.num_actions):
            observation = agent.get_observations(coordinate_init, velocity_init, scene_handler)
            impulse = self.controller(observation)
            (loss_val, (coord, velocity)), v_grad = vmapped_grad_and_value(
                self.sim_time, self.n_steps, scene_handler.jax_scene, coordinate_init,
                (velocity_init + impulse / self.constants["mass"]).detach(),
                onp.array(self.config["sim"]["coordinate_target"]),
                onp.array(self.config["sim"]["attractor_coordinate"]),
                self.constants,
            )
            impulse.backward(v_grad)
            coordinate_init = coord
        loss_val = self._exchange(loss_val)
        torch.nn.utils.clip_grad_norm_(self.controller.parameters(), grad_clip)
        self.optimizer.step()
        return loss_val

    def graphed_iteration(self, agent, scene_handler, grad_clip=5, warmup=3, fused=False):
        """-> ``run(coordinate_init [B,2], velocity_init [B,2]) -> loss`` replaying one captured CUDA
        graph per call (inputs are copied into static buffers; the returned loss is a static 0-dim
        tensor that the next replay overwrites).  The scene must not change between replays (capture
        one callable per scene); ``scene_handler.get_init_state`` stays outside the graph (its
        rejection loop reads a flag on the host).  ``fused=True``: the captured loop is ``fused_loop().run`` (no
        torch operator inside the graph: ~9 kernels per action instead of ~40)."""
        if fused:
            return self._graphed_fused(agent, scene_handler, grad_clip, warmup)
        B = self.batch_size
        x_static = torch.zeros((B, 2), dtype=torch.float32, device=self.device)
        v_static = torch.zeros((B, 2), dtype=torch.float32, device=self.device)
        x_static.copy_(torch.as_tensor(scene_handler.get_init_state(B), dtype=torch.float32, device=self.device))
        v_static.uniform_(-1, 2)
        if self._reducer is None and self.data_parallel and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:  # peer buffers are mapped outside the capture
            self._reducer = PolicyGradReducer(self.controller.parameters(), self.device, self.reduce_transport)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        # warm-up on a side stream (allocates the cached rollout buffers, lazily-initialised optimiser
        # state, cuBLAS workspaces) — the parameter updates of these iterations are rolled back
        saved = [p.detach().clone() for p in self.controller.parameters()]
        # optimiser state that already exists (a trainer that was trained eagerly first) survives the capture
        saved_opt = {id(t): t.detach().clone() for st in self.optimizer.state.values() for t in st.values()
                     if isinstance(t, torch.Tensor)}
        from ..sim.ball_sim import _Buffers
        log_start = len(_Buffers._capture_log)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.optimizer.zero_grad(set_to_none=True)
                self._action_loop(agent, scene_handler, x_static, v_static, grad_clip)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph):
            loss_static = self._action_loop(agent, scene_handler, x_static, v_static, grad_clip)
        with torch.no_grad():
            for p, q in zip(self.controller.parameters(), saved):
                p.copy_(q)
        # Adam's moments / step, in place (the graph holds pointers to these tensors): back to what they were
        # before the warm-up, zero where the warm-up created them
        with torch.no_grad():
            for st in self.optimizer.state.values():
                for k, v in st.items():
                    if isinstance(v, torch.Tensor):
                        if id(v) in saved_opt:
                            v.copy_(saved_opt[id(v)])
                        else:
                            v.zero_()
        del saved_opt
        # the graph holds raw pointers into the cached rollout buffers: keep them (and the scene whose
        # prepared workspace was captured) alive for as long as the runner lives, whatever the cache evicts
        held = list(_Buffers._capture_log[log_start:])
        del _Buffers._capture_log[log_start:]
        scene_captured = scene_handler.jax_scene
        peer = self._reducer.peer if self._reducer is not None else None
        replays = [0]

        def run(coordinate_init, velocity_init):
            if scene_handler.jax_scene is not scene_captured:
                raise RuntimeError("graphed_iteration: the scene handler switched to another scene; the graph was "
                                   "captured for one scene (capture one runner per SingleScene)")
            x_static.copy_(torch.as_tensor(coordinate_init, dtype=torch.float32, device=self.device))
            v_static.copy_(torch.as_tensor(velocity_init, dtype=torch.float32, device=self.device))
            graph.replay()
            replays[0] += 1
            if peer is not None and replays[0] % run.check_every == 0:
                peer.check()  # a timed-out exchange left NaN gradients on this rank: stop before the ranks drift
            return loss_static

        run.graph = graph
        run.buffers = held
        run.check_every = 16
        return run

    def _graphed_fused(self, agent, scene_handler, grad_clip, warmup):
        B = self.batch_size
        fl = self.fused_loop(grad_clip)
        fl.prepare_exchange()
        x_static = torch.zeros((B, 2), dtype=torch.float32, device=self.device)
        v_static = torch.zeros((B, 2), dtype=torch.float32, device=self.device)
        x_static.copy_(torch.as_tensor(scene_handler.get_init_state(B), dtype=torch.float32, device=self.device))
        v_static.uniform_(-1, 2)
        state = [t.clone() for t in (fl.pflat, fl.exp_avg, fl.exp_avg_sq, fl.step)]
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        from ..sim.ball_sim import _Buffers
        log_start = len(_Buffers._capture_log)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fl.run(agent, scene_handler, x_static, v_static)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss_static = fl.run(agent, scene_handler, x_static, v_static)
        with torch.no_grad():  # the warm-up iterations are rolled back (parameters, Adam moments, step counter)
            for t, s0 in zip((fl.pflat, fl.exp_avg, fl.exp_avg_sq, fl.step), state):
                t.copy_(s0)
            fl.gflat.zero_()
        held = list(_Buffers._capture_log[log_start:])
        del _Buffers._capture_log[log_start:]
        scene_captured = scene_handler.jax_scene
        replays = [0]

        def run(coordinate_init, velocity_init):
            if scene_handler.jax_scene is not scene_captured:
                raise RuntimeError("graphed_iteration: the scene handler switched to another scene; the graph was "
                                   "captured for one scene (capture one runner per SingleScene)")
            x_static.copy_(torch.as_tensor(coordinate_init, dtype=torch.float32, device=self.device))
            v_static.copy_(torch.as_tensor(velocity_init, dtype=torch.float32, device=self.device))
            graph.replay()
            replays[0] += 1
            if fl._peer and replays[0] % run.check_every == 0:
                fl._peer.check()
            return loss_static

        run.graph = graph
        run.buffers = held
        run.check_every = 16
        return run
