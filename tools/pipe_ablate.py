"""Latency budget of rollout_fwd_pipe_kernel's consumers by ablation (debug builds, results are WRONG by
construction — only the kernel time is read): 0 = consumers only meet the barriers, 1 = + loads and cell
look-ups, 2 = + narrow phase and collisions, 3 = everything (= the product kernel).
    python tools/pipe_ablate.py build      # here
    python tools/pipe_ablate.py run [B] [window]   # GPU box
"""
import ctypes
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIBS = {k: ROOT / "doper_b200" / "_lib" / f"libdoper_b200_ablate{k}.so" for k in (0, 1, 2, 3)}
if sys.argv[1] == "build":
    sys.path.insert(0, str(ROOT))
    from doper_b200.build import CSRC, NVCC_FLAGS
    procs = [subprocess.Popen(["nvcc", *NVCC_FLAGS, f"-DDOPER_PIPE_ABLATE={k}", "-o", str(lib), str(CSRC / "kernels.cu")]) for k, lib in LIBS.items()]
    assert all(p.wait() == 0 for p in procs)
    raise SystemExit(0)
if sys.argv[1] == "run":
    res = {}
    for k, lib in LIBS.items():
        out = subprocess.run([sys.executable, __file__, "one", str(lib)] + sys.argv[2:], capture_output=True, text=True)
        res[k] = json.loads(out.stdout.strip().splitlines()[-1]) if out.returncode == 0 else out.stderr[-500:]
    print(json.dumps(res, indent=1))
    raise SystemExit(0)

os.environ["DOPER_B200_LIB"] = sys.argv[2]
import torch  # noqa: E402

sys.path.insert(0, str(ROOT))
from doper_b200 import synthetic as syn  # noqa: E402
from doper_b200._abi import check, lib  # noqa: E402
from doper_b200._device import ptr  # noqa: E402
from doper_b200.sim import ball_sim  # noqa: E402
from doper_b200.sim.jax_geometry import JaxScene  # noqa: E402

B = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
window = int(sys.argv[4]) if len(sys.argv) > 4 else 16
dev = torch.device("cuda", 0)
w = syn.make_workload("c2", n_balls=B)
n = w["n_steps"]
scene = JaxScene(w["segments"])
_, sws = scene.prepared(dev, 0.1)
x0, v0 = torch.tensor(w["x0"], device=dev), torch.tensor(w["v0"], device=dev)
st = torch.cuda.current_stream().cuda_stream
desc = ball_sim._make_desc(B, scene, w["sim_time"], n, w["attractor"], w["constants"], ball_sim.RolloutOptions(pipe_window=window))
bufs = ball_sim._Buffers.get(dev, desc)
xT = torch.empty((B, 2), device=dev); vT = torch.empty((B, 2), device=dev)
ts = []
for rep in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    check(lib.doper_rollout_fwd(st, ctypes.byref(desc), ptr(sws), ptr(x0), ptr(v0), ptr(xT), ptr(vT), 0, 0, 0, ptr(bufs.ws)), "fwd")
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
print(json.dumps({"ms": round(sorted(ts)[len(ts) // 2], 4)}))
