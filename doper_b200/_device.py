"""Device-memory plumbing (torch is used for allocation, streams and copies only)."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("doper_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def is_device_tensor(a) -> bool:
    return isinstance(a, torch.Tensor) and a.is_cuda


def as_f32_device(a, device: torch.device) -> torch.Tensor:
    """float32, contiguous, on ``device``; zero-copy when ``a`` already is."""
    if isinstance(a, torch.Tensor):
        t = a.detach()
        if t.dtype != torch.float32:
            t = t.to(torch.float32)
        if not t.is_cuda:
            t = t.to(device, non_blocking=True)
        return t.contiguous()
    if hasattr(a, "__cuda_array_interface__"):
        return torch.as_tensor(a, device=device).to(torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32))).to(device)


def host_f32(a) -> np.ndarray:
    if isinstance(a, torch.Tensor):
        return a.detach().to("cpu", torch.float32).numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


class PinnedPool:
    """Page-locked staging buffers, reused across calls (pinning is expensive)."""

    def __init__(self):
        self._bufs: Dict[Tuple[str, int], torch.Tensor] = {}

    def get(self, tag: str, n_floats: int) -> torch.Tensor:
        key = (tag, n_floats)
        b = self._bufs.get(key)
        if b is None:
            b = torch.empty(n_floats, dtype=torch.float32, pin_memory=True)
            self._bufs[key] = b
        return b


PINNED = PinnedPool()
