"""Launch one forward-kernel variant a few times (the command ncu wraps).

    python tools/prof_fwd.py <workload> <default|nocull|tp1|cullG|tpcG> [n_balls] [reps]
"""
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

name, var = sys.argv[1], sys.argv[2]
balls = int(sys.argv[3]) if len(sys.argv) > 3 and int(sys.argv[3]) > 0 else None
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
lo = int(sys.argv[5]) if len(sys.argv) > 6 else 0
hi = int(sys.argv[6]) if len(sys.argv) > 6 else None
w = syn.make_workload(name, n_balls=balls, lo=lo, hi=hi)
B, n = w["x0"].shape[0], w["n_steps"]
dev = torch.device("cuda", 0)
scene = JaxScene(w["segments"])
_, sws = scene.prepared(dev, 0.1)
x0, v0 = torch.tensor(w["x0"], device=dev), torch.tensor(w["v0"], device=dev)
st = torch.cuda.current_stream().cuda_stream
if var == "default":
    kw = {}
elif var.startswith("tpc"):
    kw = dict(lanes_per_ball=int(var[3:]), broad_phase=True, broad_phase_time_parallel=True)
elif var.startswith("cull"):
    kw = dict(lanes_per_ball=int(var[4:]), broad_phase=True)
elif var == "tp1":
    kw = dict(smem_forward=True)
elif var == "nocull":
    kw = dict(broad_phase=False)
else:
    raise SystemExit(f"unknown variant {var}")
opts = ball_sim.RolloutOptions(**kw)
desc = ball_sim._make_desc(B, scene, w["sim_time"], n, w["attractor"], w["constants"], opts)
bufs = ball_sim._Buffers.get(dev, desc)
xT = torch.empty((B, 2), device=dev); vT = torch.empty((B, 2), device=dev)
bar = torch.empty((B, 2), device=dev); gv = torch.empty((B, 2), device=dev); gx = torch.empty((B, 2), device=dev)
ls = torch.empty(1, device=dev)
tgt = ball_sim._target2(w["target"])
for _ in range(reps):
    check(lib.doper_rollout_fwd(st, ctypes.byref(desc), ptr(sws), ptr(x0), ptr(v0), ptr(xT), ptr(vT), 0, 0, 0,
                                ptr(bufs.ws)), "fwd")
    check(lib.doper_l1_loss(st, B, ptr(xT), tgt, 0, ptr(ls), ptr(bar)), "loss")
    check(lib.doper_rollout_bwd(st, ctypes.byref(desc), ptr(sws), ptr(bufs.ws), ptr(bufs.bwd_scratch), ptr(bar), 0, ptr(gx), ptr(gv)), "bwd")
torch.cuda.synchronize()
print("ok", name, var, B, n, float(ls))
