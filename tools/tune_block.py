"""Block size of the one-wave baked forward kernels (kernels.cu balanced_block): forward time of a workload with
DOPER_BAKED_THREADS pinned to fixed sizes against the balanced default.
python tools/tune_block.py [workload] [balls] [lanes]"""
import ctypes
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from doper_b200 import synthetic as syn
from doper_b200._abi import check, lib
from doper_b200._device import ptr
from doper_b200.sim import ball_sim
from doper_b200.sim.jax_geometry import JaxScene

name = sys.argv[1] if len(sys.argv) > 1 else "c5"
balls = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda")
if name == "c5" and not balls:
    balls = 131072
w = syn.make_workload(name, n_balls=balls) if balls else syn.make_workload(name)
B, n = w["x0"].shape[0], w["n_steps"]
scene = JaxScene(w["segments"])
x0, v0 = torch.tensor(w["x0"], device=dev), torch.tensor(w["v0"], device=dev)
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {"workload": name, "B": B, "n_steps": n, "S": int(w["segments"].shape[-3]), "forward_ms": {}}
opts = ball_sim.RolloutOptions(baked_lanes=lanes) if lanes else ball_sim.RolloutOptions()
desc = ball_sim._make_desc(B, scene, w["sim_time"], n, w["attractor"], w["constants"], opts)
_, sws = scene.prepared(dev, float(desc.consts[0]))
bufs = ball_sim._Buffers.get(dev, desc)
out = torch.empty((2, B, 2), device=dev)
res["plan"] = lib.doper_rollout_plan_name(ctypes.byref(desc), 0).decode()
ref = None
for threads in (0, 64, 128, 256, 512, 1024):
    if threads:
        os.environ["DOPER_BAKED_THREADS"] = str(threads)
    else:
        os.environ.pop("DOPER_BAKED_THREADS", None)
    f = lambda: lib.doper_rollout_fwd(st, ctypes.byref(desc), ptr(sws), ptr(x0), ptr(v0), ptr(out[0]), ptr(out[1]), 0, 0, 0,
                                      ptr(bufs.ws))
    if f() != 0:
        torch.cuda.synchronize()
        res["forward_ms"][threads or "balanced"] = None
        continue
    for _ in range(2):
        f()
    ts = []
    for _ in range(7):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    res["forward_ms"][threads or "balanced"] = round(ts[len(ts) // 2], 4)
    o = out.clone()
    if ref is None:
        ref = o
    assert torch.equal(o.view(torch.int32), ref.view(torch.int32)), "results depend on the block size"
os.environ.pop("DOPER_BAKED_THREADS", None)
print(json.dumps(res))
