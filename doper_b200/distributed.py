"""Multi-GPU plumbing of the rollout path: shard balls/environments, reduce the policy gradient.

The path shards trivially: balls are independent (``vmap``, reference
``ball_sim.py:332``) and the loss is a plain sum over balls (``:324``).  One process per GPU
(``torchrun``); rank r simulates a contiguous slice of the batch on its own device and keeps the
per-ball ``dL/dv0`` local (it flows into the local controller backward).  Exactly one exchange
happens per optimiser step: the sum over ranks of the controller's parameter gradients
(6,602 fp32 for the reference's 25-50-100-2 MLP, ``doper/networks/controllers.py:12-18``) and of
the scalar loss — one all-reduce of one flat buffer (NCCL over NVLink on GPUs, gloo in CPU tests).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """Initialise ``torch.distributed`` from torchrun's environment -> (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return rank, world, local_rank


def ball_shard(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice ``[lo, hi)`` of the global batch owned by ``rank`` (remainder to the
    first ranks), so that results do not depend on the number of GPUs."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(int(n_total), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class PolicyGradReducer:
    """All-reduce(sum) of the controller gradients and the scalar loss in ONE flat buffer."""

    def __init__(self, params: Iterable[torch.nn.Parameter], device: torch.device = None):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        self.numel = sum(p.numel() for p in self.params)
        device = device or (self.params[0].device if self.params else torch.device("cpu"))
        self.buf = torch.zeros(self.numel + 1, dtype=torch.float32, device=device)

    @torch.no_grad()
    def reduce(self, local_loss) -> torch.Tensor:
        """Sums ``p.grad`` over ranks in place; returns the global loss (0-dim tensor)."""
        o = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.buf[o:o + n].zero_()
            else:
                self.buf[o:o + n].copy_(p.grad.reshape(-1))
            o += n
        self.buf[o] = float(local_loss) if not isinstance(local_loss, torch.Tensor) else local_loss.to(self.buf.device)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.buf, op=dist.ReduceOp.SUM)
        o = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                p.grad = self.buf[o:o + n].reshape(p.shape).clone()
            else:
                p.grad.copy_(self.buf[o:o + n].reshape(p.shape))
            o += n
        return self.buf[o].clone()
