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
    """All-reduce(sum) of the controller gradients and the scalar loss in ONE flat buffer.

    ``transport``: "nvlink" = the one-shot peer-memory kernel (``PeerAllReducer``; CUDA, one node),
    "collective" = ``torch.distributed.all_reduce`` (NCCL on GPUs, gloo in the CPU tests), "auto" =
    nvlink when the parameters are on CUDA and a multi-rank group is initialised, else collective."""

    def __init__(self, params: Iterable[torch.nn.Parameter], device: torch.device = None, transport: str = "auto"):
        self.params: List[torch.nn.Parameter] = [p for p in params]
        self.numel = sum(p.numel() for p in self.params)
        device = device or (self.params[0].device if self.params else torch.device("cpu"))
        self.buf = torch.zeros(self.numel + 1, dtype=torch.float32, device=device)
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if transport == "auto":
            transport = "auto-nvlink" if (multi and torch.device(device).type == "cuda") else "collective"  # falls back
        self.peer = None
        if transport in ("nvlink", "auto-nvlink") and multi:
            try:
                self.peer = PeerAllReducer(self.numel + 1, torch.device(device))
            except RuntimeError:
                if transport == "nvlink":  # asked for explicitly: do not fall back silently
                    raise

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
        if self.peer is not None:
            self.peer.all_reduce(self.buf, out=self.buf)
            if not torch.cuda.is_current_stream_capturing():
                self.peer.check()  # eager call: one 4-byte read; captured calls are checked by the graph's runner
        elif dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
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


class PeerAllReducer:
    """One-shot all-reduce(sum) of a small fp32 vector over NVLink peer memory (csrc/peer.cuh):
    every rank's buffer is mapped into every other rank through CUDA IPC once, and each call is ONE
    small kernel per rank (push the input into every peer with the step number inside each 16-byte
    store, poll the own buffer for the peers' words, sum in rank order).  For the trainer's 26 KB
    exchange this replaces a NCCL all-reduce whose latency (25 us at 2 GPUs, 31 us at 8) is a visible
    part of a 0.4 ms step.  The step counter lives on the device, so ``all_reduce`` can be captured in
    a CUDA graph.  Needs one process per GPU on ONE node and an initialised ``torch.distributed``
    group (used once, to exchange the handles).
    """

    def __init__(self, numel: int, device: torch.device, timeout_seconds: float = 60.0):
        import ctypes

        from ._abi import check, lib
        self._lib, self._check, self._ct = lib, check, ctypes
        self.n = int(numel)
        self.device = device
        self.timeout_seconds = float(timeout_seconds)  # a peer that never arrives: give up, flag, NaN (see check())
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if self.world > 16:
            raise ValueError("PeerAllReducer: at most 16 ranks")
        torch.cuda.set_device(device)
        self._own, self._opened = None, []
        self._ptrs = (ctypes.c_void_p * self.world)()
        handle = ctypes.create_string_buffer(64)
        err = None
        try:
            own = ctypes.c_void_p()
            check(lib.doper_peer_alloc(self.n, ctypes.byref(own), handle), "doper_peer_alloc")
            self._own = own.value
        except Exception as e:  # still take part in the collectives below: the group must fail together
            err = e
        mine = torch.tensor(list(handle.raw), dtype=torch.uint8, device=device)
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine)
        if err is None:
            try:
                if os.environ.get("DOPER_PEER_FORCE_FAIL") == str(self.rank):
                    raise RuntimeError("forced failure (DOPER_PEER_FORCE_FAIL)")
                for r, h in enumerate(gathered):
                    if r == self.rank:
                        self._ptrs[r] = self._own
                        continue
                    p = ctypes.c_void_p()
                    check(lib.doper_peer_open(bytes(h.cpu().tolist()), ctypes.byref(p)), "doper_peer_open")
                    self._ptrs[r] = p.value
                    self._opened.append(p.value)
            except Exception as e:
                err = e
        # every rank learns whether EVERY rank mapped all its peers; otherwise all of them give up together
        ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            self.close(barrier=False)
            raise RuntimeError(f"PeerAllReducer: peer memory is not available on every rank ({err or 'another rank failed'})")
        self.out = torch.empty(self.n, dtype=torch.float32, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        dist.barrier(device_ids=[device.index])  # every peer mapping exists before the first kernel

    def all_reduce(self, x: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
        """Sum over ranks of ``x`` (fp32, ``numel`` elements, contiguous, on this device) -> ``out``
        (default ``self.out``, overwritten by the next call; ``out`` may be ``x``).  Enqueued on the
        current stream; no host synchronisation; every rank must make the same sequence of calls."""
        out = self.out if out is None else out
        if x.dtype != torch.float32 or x.numel() != self.n or not x.is_contiguous() or x.device != self.out.device:
            raise ValueError("PeerAllReducer.all_reduce: x must be a contiguous fp32 tensor of numel elements on the device")
        if out.dtype != torch.float32 or out.numel() != self.n or not out.is_contiguous() or out.device != self.out.device:
            raise ValueError("PeerAllReducer.all_reduce: bad out tensor")
        self._check(self._lib.doper_allreduce_oneshot(torch.cuda.current_stream().cuda_stream, self.rank, self.world,
                                                      self._ptrs, self.n, x.data_ptr(), out.data_ptr(),
                                                      self.status.data_ptr(), self.timeout_seconds),
                    "doper_allreduce_oneshot")
        return out

    def check(self) -> None:
        """Host synchronisation point: raises if any call since the last check timed out waiting for a
        peer (that call's result is NaN on this rank while the peers may hold a full sum: the ranks'
        parameters have diverged and must be re-broadcast before training continues)."""
        if int(self.status.item()) != 0:
            self.status.zero_()
            raise RuntimeError("PeerAllReducer: a peer did not arrive within "
                               f"{self.timeout_seconds:g} s; this rank's result was NaN-filled. Re-synchronise "
                               "the parameters (broadcast from rank 0) before the next optimiser step.")

    def close(self, barrier: bool = True) -> None:
        torch.cuda.synchronize(self.device)
        if barrier and dist.is_initialized():
            dist.barrier(device_ids=[self.device.index])  # nobody unmaps while a peer may still read
        for p in self._opened:
            self._lib.doper_peer_close(p)
        self._opened = []
        if self._own:
            self._lib.doper_peer_free(self._own)
            self._own = None
