"""The trainer's action loop on the sm_100a kernels only.

``FusedActionLoop`` runs ``BallAttractorTrainer``'s iteration (doper/trainers/ball_trainer.py:49-76) without any
torch operator between the kernels: per action one observation kernel (``doper_agent_observations``), one
controller-forward kernel that writes ``v0 = velocity_init + impulse / mass`` directly, the rollout
(forward + loss + backward), one controller-backward kernel pair that ADDS into the flat gradient buffer; per
iteration one exchange (that same buffer, gradients ‖ loss: nothing is packed) and one gradient-clip + Adam
kernel.  The controller's parameters and gradients are views into two flat vectors, so ``trainer.controller``
(the reference's ``LinearReLUController``) keeps working for ``forward`` / ``state_dict`` and sees the updates.

Everything launches on torch's current stream and synchronises nowhere: the loop can be captured in a CUDA graph
(``BallAttractorTrainer.graphed_iteration(fused=True)``).
"""
from __future__ import annotations

import ctypes

import numpy as onp
import torch

from .._abi import check, lib
from .._device import ptr, stream_ptr
from ..sim import ball_sim
from ..sim.jax_geometry import _as_scene

__all__ = ["FusedActionLoop"]


class FusedActionLoop:
    def __init__(self, trainer, grad_clip: float = 5.0):
        self.tr = trainer
        dev = trainer.device
        lin = [m for m in trainer.controller.controller if isinstance(m, torch.nn.Linear)]
        if len(lin) != 3 or lin[2].out_features < 2:
            raise ValueError("FusedActionLoop: the controller must be Linear-ReLU-Linear-ReLU-Linear (controllers.py:12-18)")
        self.dims = (ctypes.c_int32 * 4)(lin[0].in_features, lin[0].out_features, lin[1].out_features, lin[2].out_features)
        params = [p for m in lin for p in (m.weight, m.bias)]
        self.n = sum(p.numel() for p in params)
        self.pflat = torch.empty(self.n, dtype=torch.float32, device=dev)
        self.gflat = torch.zeros(self.n + 2, dtype=torch.float32, device=dev)  # gradients ‖ loss (‖ pad): the exchange buffer
        o = 0
        with torch.no_grad():
            for p in params:  # parameters and gradients become views into the flat vectors
                k = p.numel()
                self.pflat[o:o + k].copy_(p.detach().reshape(-1))
                p.data = self.pflat[o:o + k].view(p.shape)
                p.grad = self.gflat[o:o + k].view(p.shape)
                o += k
        self.exp_avg = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.step = torch.zeros(1, dtype=torch.float32, device=dev)
        self.total_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        g = trainer.optimizer.param_groups[0]
        self.lr, (self.beta1, self.beta2), self.eps = float(g["lr"]), g["betas"], float(g["eps"])
        self.grad_clip = float(grad_clip)
        self.mass32 = float(onp.float32(trainer.constants["mass"]))
        self._bufs = {}
        self._peer = None

    # -- buffers of one (scene, batch) ----------------------------------------------------------------
    def _plan(self, scene, B):
        key = (id(scene), B)
        pl = self._bufs.get(key)
        if pl is None:
            tr, dev = self.tr, self.tr.device
            desc = ball_sim._make_desc(B, scene, tr.sim_time, tr.n_steps, onp.array(tr.config["sim"]["attractor_coordinate"]),
                                       tr.constants, ball_sim._DEFAULT_OPTIONS)
            _, sws = scene.prepared(dev, float(desc.consts[0]))
            bufs = ball_sim._Buffers.get(dev, desc)
            pl = {
                "desc": desc, "sws": sws, "bufs": bufs, "scene": scene,
                "v0": torch.empty((B, 2), dtype=torch.float32, device=dev),
                "x": [torch.empty((B, 2), dtype=torch.float32, device=dev) for _ in range(2)],
                "vT": torch.empty((B, 2), dtype=torch.float32, device=dev),
                "g": torch.empty((B, 2), dtype=torch.float32, device=dev),
                "scratch": torch.empty(max(16, lib.doper_controller_scratch_bytes(self.dims, B)), dtype=torch.uint8, device=dev),
                "tgt": ball_sim._target2(onp.array(tr.config["sim"]["coordinate_target"])),
            }
            self._bufs[key] = pl
        return pl

    def run(self, agent, scene_handler, coordinate_init: torch.Tensor, velocity_init: torch.Tensor) -> torch.Tensor:
        """One iteration (num_actions actions + exchange + clip + Adam); returns the loss of the last action as a
        0-dim view into the exchange buffer (summed over ranks when data parallel)."""
        tr = self.tr
        scene = _as_scene(scene_handler.jax_scene)
        B = int(coordinate_init.shape[0])
        pl = self._plan(scene, B)
        desc, sws, bufs = pl["desc"], pl["sws"], pl["bufs"]
        st = stream_ptr()
        x = coordinate_init
        loss = self.gflat[self.n:self.n + 1]
        for a in range(tr.num_actions):
            obs = agent.get_observations(x, velocity_init, scene_handler)  # [B, d_in] on the device, one kernel
            check(lib.doper_controller_fwd(st, self.dims, B, ptr(self.pflat), ptr(obs), ptr(velocity_init), self.mass32, 0,
                                           ptr(pl["v0"])), "doper_controller_fwd")
            xT = pl["x"][a & 1]
            check(lib.doper_value_and_grad(st, ctypes.byref(desc), ptr(sws), ptr(x), ptr(pl["v0"]), pl["tgt"], ptr(bufs.ws),
                                           ptr(bufs.bwd_scratch), ptr(bufs.scratch), ptr(loss), 0, ptr(xT), ptr(pl["vT"]),
                                           ptr(pl["g"]), 0), "doper_value_and_grad")
            check(lib.doper_controller_bwd(st, self.dims, B, ptr(self.pflat), ptr(obs), ptr(pl["g"]), self.mass32,
                                           ptr(pl["scratch"]), ptr(self.gflat)), "doper_controller_bwd")
            x = xT  # ball_trainer.py:72; the velocity is NOT carried over (:73, sic)
        self._exchange()
        check(lib.doper_adam_clip_step(st, self.n, ptr(self.pflat), ptr(self.gflat), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                       ptr(self.step), self.grad_clip, self.lr, float(self.beta1), float(self.beta2), self.eps,
                                       ptr(self.total_norm)), "doper_adam_clip_step")
        return loss[0]

    def _exchange(self):
        tr = self.tr
        dist = torch.distributed
        if not (tr.data_parallel and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        if self._peer is None and tr.reduce_transport != "collective":
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("FusedActionLoop: call prepare_exchange() before graph capture")
            self.prepare_exchange()
        if self._peer:
            self._peer.all_reduce(self.gflat, out=self.gflat)  # gradients ‖ loss, in place, one kernel
        else:
            dist.all_reduce(self.gflat, op=dist.ReduceOp.SUM)

    def prepare_exchange(self):
        """Map the peer buffers (outside any capture)."""
        from ..distributed import PeerAllReducer
        dist = torch.distributed
        if self._peer is None and self.tr.data_parallel and dist.is_available() and dist.is_initialized() and \
                dist.get_world_size() > 1 and self.tr.reduce_transport != "collective":
            try:
                self._peer = PeerAllReducer(self.n + 2, self.tr.device)
            except RuntimeError:
                if self.tr.reduce_transport == "nvlink":
                    raise
                self._peer = False
        return self._peer
