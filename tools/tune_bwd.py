"""Backward-kernel timing vs chunk count (checkpoint period) on a workload."""
import ctypes
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from doper_b200 import synthetic as syn  # noqa: E402
from doper_b200._abi import check, lib  # noqa: E402
from doper_b200._device import ptr  # noqa: E402
from doper_b200.sim import ball_sim  # noqa: E402
from doper_b200.sim.jax_geometry import JaxScene  # noqa: E402
from tools.tune_fwd import time_ms  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
balls = int(sys.argv[2]) if len(sys.argv) > 2 else None
w = syn.make_workload(name, n_balls=balls)
B, n = w["x0"].shape[0], w["n_steps"]
dev = torch.device("cuda", 0)
scene = JaxScene(w["segments"])
_, sws = scene.prepared(dev, 0.1)
x0, v0 = torch.tensor(w["x0"], device=dev), torch.tensor(w["v0"], device=dev)
st = torch.cuda.current_stream().cuda_stream
tgt = ball_sim._target2(w["target"])
for K, seq in ((0, False), (32, False), (16, False), (8, False), (4, False), (0, True)):
    opts = ball_sim.RolloutOptions(ckpt_every=K, sequential_backward=seq)
    desc = ball_sim._make_desc(B, scene, w["sim_time"], n, w["attractor"], w["constants"], opts)
    bufs = ball_sim._Buffers.get(dev, desc)
    xT = torch.empty((B, 2), device=dev); vT = torch.empty((B, 2), device=dev)
    bar = torch.empty((B, 2), device=dev); gv = torch.empty((B, 2), device=dev); gx = torch.empty((B, 2), device=dev)
    ls = torch.empty(1, device=dev)
    check(lib.doper_rollout_fwd(st, ctypes.byref(desc), ptr(sws), ptr(x0), ptr(v0), ptr(xT), ptr(vT), 0, 0, 0, ptr(bufs.ws)), "fwd")
    check(lib.doper_l1_loss(st, B, ptr(xT), tgt, 0, ptr(ls), ptr(bar)), "loss")
    ms = time_ms(lambda: check(lib.doper_rollout_bwd(st, ctypes.byref(desc), ptr(sws), ptr(bufs.ws), ptr(bufs.bwd_scratch), ptr(bar), 0, ptr(gx), ptr(gv)), "bwd"))
    print(f"{name} B={B} K={lib.doper_rollout_ckpt_every(ctypes.byref(desc))} sequential={seq} kernel={lib.doper_rollout_plan_name(ctypes.byref(desc), 1).decode()}: bwd {ms:.4f} ms", flush=True)
