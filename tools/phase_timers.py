"""Where does the time-parallel forward kernel spend its cycles?  Builds a DEBUG copy of the library with
-DDOPER_PHASE_TIMERS (clock64 at the phase boundaries of every warp; build it here, run on the GPU box):

    python tools/phase_timers.py build          # here (CPU): writes doper_b200/_lib/libdoper_b200_phase.so
    python tools/phase_timers.py run [c2] [B]   # on the GPU box
"""
import ctypes
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
DBG = ROOT / "doper_b200" / "_lib" / "libdoper_b200_phase.so"
PHASES = ["contact mode", "list refresh", "free flight (Euler chain)", "list check / exact distance",
          "votes, shuffles, collision resolve", "commit (log, checkpoints)", "loop tail", "prologue / epilogue"]

if sys.argv[1] == "build":
    sys.path.insert(0, str(ROOT))
    from doper_b200.build import CSRC, NVCC_FLAGS
    subprocess.run(["nvcc", *NVCC_FLAGS, "-DDOPER_PHASE_TIMERS", "-o", str(DBG), str(CSRC / "kernels.cu")], check=True)
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
balls = int(sys.argv[3]) if len(sys.argv) > 3 else None
dev = torch.device("cuda", 0)
out = {}
for regime, sigma in (("default", 10.0), ("reference_regime", 0.0)):
    w = syn.make_workload(name, n_balls=balls, impulse_sigma=sigma)
    B, n = w["x0"].shape[0], w["n_steps"]
    scene = JaxScene(w["segments"])
    _, sws = scene.prepared(dev)
    x0, v0 = torch.tensor(w["x0"], device=dev), torch.tensor(w["v0"], device=dev)
    st = torch.cuda.current_stream().cuda_stream
    desc = ball_sim._make_desc(B, scene, w["sim_time"], n, w["attractor"], w["constants"], ball_sim.RolloutOptions())
    bufs = ball_sim._Buffers.get(dev, desc)
    xT = torch.empty((B, 2), device=dev); vT = torch.empty((B, 2), device=dev)
    dbg = torch.zeros(8, dtype=torch.int64, device=dev)
    for rep in range(3):
        dbg.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        check(lib.doper_rollout_fwd(st, ctypes.byref(desc), ptr(sws), ptr(x0), ptr(v0), ptr(xT), ptr(vT), 0, 0, ptr(dbg),
                                    ptr(bufs.ws)), "fwd")
        b.record()
        torch.cuda.synchronize()
    cyc = dbg.cpu().numpy().astype(np.float64)
    out[regime] = {"kernel_ms_with_timers": round(a.elapsed_time(b), 4),
                   "share_of_warp_cycles": {PHASES[k]: round(float(cyc[k] / cyc.sum()), 4) for k in range(8)},
                   "mean_warp_cycles": round(float(cyc.sum() / ((B + 1) // 2)), 1)}
print(json.dumps(out, indent=1))
Path("gpurun_out").mkdir(exist_ok=True)
Path(f"gpurun_out/phase_timers_{name}.json").write_text(json.dumps(out))
