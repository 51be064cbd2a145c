"""Forward time of a workload's default plan under several values of a launcher tuning variable (read per launch):
python tools/tune_env.py DOPER_PIPE_BALLS 16,12,8,4 [workload] [balls] [window]      (also DOPER_BAKED_THREADS)
Checks that x_T / v_T do not depend on the value."""
import ctypes
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from doper_b200 import synthetic as syn
from doper_b200._abi import lib
from doper_b200._device import ptr
from doper_b200.sim import ball_sim
from doper_b200.sim.jax_geometry import JaxScene

env, values = sys.argv[1], sys.argv[2].split(",")
name = sys.argv[3] if len(sys.argv) > 3 else "c2"
balls = int(sys.argv[4]) if len(sys.argv) > 4 else 0
window = int(sys.argv[5]) if len(sys.argv) > 5 else 0  # pin the pipelined kernel's window length (8, 16, 32)
dev = torch.device("cuda")
w = syn.make_workload(name, n_balls=balls) if balls else syn.make_workload(name)
B, n = w["x0"].shape[0], w["n_steps"]
scene = JaxScene(w["segments"])
x0, v0 = torch.tensor(w["x0"], device=dev), torch.tensor(w["v0"], device=dev)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
desc = ball_sim._make_desc(B, scene, w["sim_time"], n, w["attractor"], w["constants"], ball_sim.RolloutOptions(pipe_window=window) if window else ball_sim.RolloutOptions())
_, sws = scene.prepared(dev, float(desc.consts[0]))
bufs = ball_sim._Buffers.get(dev, desc)
out = torch.empty((2, B, 2), device=dev)
res = {"workload": name, "B": B, "n_steps": n, "plan": lib.doper_rollout_plan_name(ctypes.byref(desc), 0).decode(), env: {}}
ref = None
for v in ["default"] + values:
    if v == "default":
        os.environ.pop(env, None)
    else:
        os.environ[env] = v
    f = lambda: lib.doper_rollout_fwd(st, ctypes.byref(desc), ptr(sws), ptr(x0), ptr(v0), ptr(out[0]), ptr(out[1]), 0, 0, 0, ptr(bufs.ws))
    for _ in range(3):
        assert f() == 0
    ts = []
    for _ in range(9):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    res[env][v] = round(ts[len(ts) // 2], 4)
    o = out.clone()
    ref = o if ref is None else ref
    assert torch.equal(o.view(torch.int32), ref.view(torch.int32)), "results depend on the tuning variable"
os.environ.pop(env, None)
print(json.dumps(res))
