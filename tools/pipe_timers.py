"""Where do the producer and consumer warps of rollout_fwd_pipe_kernel spend their cycles?  Debug build with
-DDOPER_PIPE_TIMERS (clock64 at the phase boundaries, lane 0 of every warp):

    python tools/pipe_timers.py build           # here (CPU): doper_b200/_lib/libdoper_b200_pipet.so
    python tools/pipe_timers.py run [c2] [B] [window]   # on the GPU box
"""
import ctypes
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
DBG = ROOT / "doper_b200" / "_lib" / "libdoper_b200_pipet.so"
P_PH = ["mailbox + window header", "free-flight chain (U steps)", "barrier wait (consumer busy)"]
C_PH = ["barrier wait (producer / slowest consumer busy)", "loads + cell look-up", "first hit + collision resolve",
        "commit (log, checkpoints) + vote", "narrow phase (list of the cell)"]

if sys.argv[1] == "build":
    sys.path.insert(0, str(ROOT))
    from doper_b200.build import CSRC, NVCC_FLAGS
    subprocess.run(["nvcc", *NVCC_FLAGS, "-DDOPER_PIPE_TIMERS", "-o", str(DBG), str(CSRC / "kernels.cu")], check=True)
    print(DBG)
    raise SystemExit(0)

os.environ["DOPER_B200_LIB"] = str(DBG)
import numpy as np  # noqa: E402
import torch  # noqa: E402

sys.path.insert(0, str(ROOT))
from doper_b200 import synthetic as syn  # noqa: E402
from doper_b200._abi import check, lib  # noqa: E402
from doper_b200._device import ptr  # noqa: E402
from doper_b200.sim import ball_sim  # noqa: E402
from doper_b200.sim.jax_geometry import JaxScene  # noqa: E402

name = sys.argv[2] if len(sys.argv) > 2 else "c2"
balls = int(sys.argv[3]) if len(sys.argv) > 3 and int(sys.argv[3]) > 0 else None
window = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dev = torch.device("cuda", 0)
sigma = float(sys.argv[5]) if len(sys.argv) > 5 else 10.0  # 0 = reference regime (slow balls, rare collisions)
w = syn.make_workload(name, n_balls=balls, impulse_sigma=sigma)
B, n = w["x0"].shape[0], w["n_steps"]
scene = JaxScene(w["segments"])
_, sws = scene.prepared(dev, 0.1)
x0, v0 = torch.tensor(w["x0"], device=dev), torch.tensor(w["v0"], device=dev)
st = torch.cuda.current_stream().cuda_stream
desc = ball_sim._make_desc(B, scene, w["sim_time"], n, w["attractor"], w["constants"], ball_sim.RolloutOptions(pipe_window=window))
bufs = ball_sim._Buffers.get(dev, desc)
xT = torch.empty((B, 2), device=dev); vT = torch.empty((B, 2), device=dev)
dbg = torch.zeros(16, dtype=torch.int64, device=dev)
for rep in range(3):
    dbg.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    check(lib.doper_rollout_fwd(st, ctypes.byref(desc), ptr(sws), ptr(x0), ptr(v0), ptr(xT), ptr(vT), 0, 0, ptr(dbg), ptr(bufs.ws)), "fwd")
    b.record()
    torch.cuda.synchronize()
raw = dbg.cpu().numpy().astype(np.float64)
groups = (B + 15) // 16
pw, cwn = raw[7], raw[15]
hits = ball_sim.rollout_history(w["sim_time"], n, scene, w["x0"], w["v0"], w["attractor"], w["constants"])["hit"]
out = {"workload": name, "impulse_sigma": sigma, "collisions": int(hits.sum()), "max_collisions_per_ball": int(hits.sum(0).max()), "B": B, "n_steps": n, "kernel": lib.doper_rollout_plan_name(ctypes.byref(desc), 0).decode(),
       "kernel_ms_with_timers": round(a.elapsed_time(b), 4), "groups": groups,
       "windows_per_group": round(pw / groups, 1),
       "producer_cycles_per_window": {P_PH[k]: round(raw[k] / pw, 1) for k in range(3)},
       "consumer_cycles_per_window_mean_over_consumer_warps": {C_PH[k]: round(raw[8 + k] / cwn, 1) for k in range(5)},
       "consumer_windows_with_a_list_in_some_lane": round(raw[13] / cwn, 4), "consumer_windows_with_narrow_phase_over_1000_cycles": round(raw[14] / cwn, 4)}
print(json.dumps(out, indent=1))
Path("gpurun_out").mkdir(exist_ok=True)
Path(f"gpurun_out/pipe_timers_{name}_{B}.json").write_text(json.dumps(out))
