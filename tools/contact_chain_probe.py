"""How long is the serial chain of ONE ball in sliding contact?  Runs the ball of BASELINE configs[4] that collides
on 1,972 of its 2,000 steps (found with the oracle) ALONE on the GPU through several forward kernels.
python tools/contact_chain_probe.py [ball index]"""
import ctypes
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from doper_b200 import synthetic as syn
from doper_b200._abi import lib
from doper_b200._device import ptr
from doper_b200.sim import ball_sim
from doper_b200.sim.jax_geometry import JaxScene

idx = int(sys.argv[1]) if len(sys.argv) > 1 else 36034
only_first = len(sys.argv) > 2 and sys.argv[2] == "profile"  # the command ncu wraps: first variant only
dev = torch.device("cuda")
w = syn.make_workload("c5", n_balls=131072, lo=idx, hi=idx + 1)
B, n = w["x0"].shape[0], w["n_steps"]
scene = JaxScene(w["segments"])
x0, v0 = torch.tensor(w["x0"], device=dev), torch.tensor(w["v0"], device=dev)
st = torch.cuda.current_stream().cuda_stream
res = {"ball": idx, "n_steps": n, "forward_ms": {}}
for name, opts in (("one thread per ball, no hand-over", ball_sim.RolloutOptions(baked_lanes=1, contact_handover=False)),
                   ("one thread per ball, hand-over", ball_sim.RolloutOptions(baked_lanes=1)),
                   ("lane = step, 16 lanes", ball_sim.RolloutOptions(baked_lanes=16, pipelined=False)),
                   ("pipelined", ball_sim.RolloutOptions()))[: 1 if only_first else 4]:
    desc = ball_sim._make_desc(B, scene, w["sim_time"], n, w["attractor"], w["constants"], opts)
    _, sws = scene.prepared(dev, float(desc.consts[0]))
    bufs = ball_sim._Buffers.get(dev, desc)
    out = torch.empty((2, B, 2), device=dev)
    f = lambda: lib.doper_rollout_fwd(st, ctypes.byref(desc), ptr(sws), ptr(x0), ptr(v0), ptr(out[0]), ptr(out[1]), 0, 0, 0, ptr(bufs.ws))
    assert f() == 0
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    hits = int(bufs.ws[:4].view(torch.int32).item())
    res["forward_ms"][name] = {"ms": round(ts[len(ts) // 2], 4), "plan": lib.doper_rollout_plan_name(ctypes.byref(desc), 0).decode(),
                               "collisions": hits, "us_per_step": round(1e3 * ts[len(ts) // 2] / n, 3)}
print(json.dumps(res, indent=1))
