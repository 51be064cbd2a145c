"""Is the pipelined forward bound by the FEW balls with many collisions or by the collisions of all balls?  Times the
forward of BASELINE configs[1] with every ball that collides more than m times replaced by a copy of a ball that never
collides, m = inf, 8, 4, 2, 1, 0 (collision counts from the oracle).
python tools/collision_cap_probe.py [workload] [balls]"""
import ctypes
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from doper_b200 import synthetic as syn
from doper_b200._abi import lib
from doper_b200._device import ptr
from doper_b200.sim import ball_sim
from doper_b200.sim.jax_geometry import JaxScene

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
balls = int(sys.argv[2]) if len(sys.argv) > 2 else None
w = syn.make_workload(name, n_balls=balls)
B, n = w["x0"].shape[0], w["n_steps"]
hits = ball_sim.rollout_history(w["sim_time"], n, w["segments"], w["x0"], w["v0"], w["attractor"], w["constants"])["hit"].sum(0)
quiet = int(np.nonzero(hits == 0)[0][0])
dev = torch.device("cuda")
scene = JaxScene(w["segments"])
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {"workload": name, "B": B, "n_steps": n, "collisions": int(hits.sum()), "max_per_ball": int(hits.max()), "caps": {}}
for cap in (10 ** 9, 8, 4, 2, 1, 0):
    x0, v0 = w["x0"].copy(), w["v0"].copy()
    bad = hits > cap
    x0[bad], v0[bad] = x0[quiet], v0[quiet]
    xd, vd = torch.tensor(x0, device=dev), torch.tensor(v0, device=dev)
    desc = ball_sim._make_desc(B, scene, w["sim_time"], n, w["attractor"], w["constants"], ball_sim.RolloutOptions())
    _, sws = scene.prepared(dev, float(desc.consts[0]))
    bufs = ball_sim._Buffers.get(dev, desc)
    out = torch.empty((2, B, 2), device=dev)
    f = lambda: lib.doper_rollout_fwd(st, ctypes.byref(desc), ptr(sws), ptr(xd), ptr(vd), ptr(out[0]), ptr(out[1]), 0, 0, 0, ptr(bufs.ws))
    for _ in range(3):
        assert f() == 0
    ts = []
    for _ in range(9):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    res["caps"]["inf" if cap > 10 ** 6 else cap] = {"balls_replaced": int(bad.sum()), "collisions_left": int(hits[~bad].sum()),
                                                    "forward_ms": round(ts[len(ts) // 2], 4)}
res["plan"] = lib.doper_rollout_plan_name(ctypes.byref(desc), 0).decode()
print(json.dumps(res))
