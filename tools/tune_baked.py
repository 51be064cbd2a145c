"""Forward-kernel timing of the baked broad phase at every lane count (and the per-ball-list kernels for
comparison): python tools/tune_baked.py [workload] [balls]."""
import ctypes
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from doper_b200 import synthetic as syn
from doper_b200._abi import check, lib
from doper_b200._device import ptr
from doper_b200.sim import ball_sim
from doper_b200.sim.jax_geometry import JaxScene

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
balls = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda")
w = syn.make_workload(name, n_balls=balls or None) if balls else syn.make_workload(name)
if name == "c5" and not balls:
    w = syn.make_workload(name, n_balls=131072)
B, n = w["x0"].shape[0], w["n_steps"]
scene = JaxScene(w["segments"])
x0, v0 = torch.tensor(w["x0"], device=dev), torch.tensor(w["v0"], device=dev)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {"workload": name, "B": B, "n_steps": n, "S": int(w["segments"].shape[-3])}


def time_fwd(opts, reps=10):
    desc = ball_sim._make_desc(B, scene, w["sim_time"], n, w["attractor"], w["constants"], opts)
    _, sws = scene.prepared(dev, float(desc.consts[0]))
    bufs = ball_sim._Buffers.get(dev, desc)
    out = torch.empty((2, B, 2), device=dev)
    f = lambda: check(lib.doper_rollout_fwd(st, ctypes.byref(desc), ptr(sws), ptr(x0), ptr(v0), ptr(out[0]), ptr(out[1]), 0, 0, 0,
                                            ptr(bufs.ws)), "fwd")
    for _ in range(3):
        f()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return round(ts[len(ts) // 2], 4), lib.doper_rollout_plan_name(ctypes.byref(desc), 0).decode()


res["default"] = time_fwd(ball_sim.RolloutOptions())
for u in (8, 16, 32):
    if B <= 300_000:
        res[f"pipe_{u}"] = time_fwd(ball_sim.RolloutOptions(pipe_window=u))
for g in (1, 2, 4, 8, 16):
    if B * g <= 4_500_000:
        res[f"baked_{g}"] = time_fwd(ball_sim.RolloutOptions(baked_lanes=g))
res["per_ball_lists"] = time_fwd(ball_sim.RolloutOptions(baked_broad_phase=False))
print(json.dumps(res))
