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
    _, sws = scene.prepared(dev, 0.1)
    x0, v0 = torch.tensor(w["x0"], device=dev), torch.tensor(w["v0"], device=dev)
    st = torch.cuda.current_stream().cuda_stream
    desc = ball_sim._make_desc(B, scene, w["sim_time"], n, w["attractor"], w["constants"], ball_sim.RolloutOptions())
    bufs = ball_sim._Buffers.get(dev, desc)
    xT = torch.empty((B, 2), device=dev); vT = torch.empty((B, 2), device=dev)
    dbg = torch.zeros(16 + 4 * ((B + 1) // 2 + 64), dtype=torch.int64, device=dev)
    for rep in range(3):
        dbg.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        check(lib.doper_rollout_fwd(st, ctypes.byref(desc), ptr(sws), ptr(x0), ptr(v0), ptr(xT), ptr(vT), 0, 0, ptr(dbg),
                                    ptr(bufs.ws)), "fwd")
        b.record()
        torch.cuda.synchronize()
    raw = dbg.cpu().numpy().astype(np.float64)
    cyc = raw[:8]
    n_warps = (B + 1) // 2
    per = raw[16:16 + 4 * n_warps].reshape(n_warps, 4)
    hist = ball_sim.rollout_history(w["sim_time"], n, scene, w["x0"], w["v0"], w["attractor"], w["constants"])["hit"].sum(0)
    speed = np.linalg.norm(w["v0"], axis=1)
    order = np.argsort(-per[:, 0])[:12]
    slow = [{"warp": int(k), "cycles": int(per[k, 0]), "rounds": int(per[k, 1]), "refresh_passes": int(per[k, 2]),
             "refresh_cycles": int(per[k, 3]), "hits": [int(hist[2 * k]), int(hist[min(2 * k + 1, B - 1)])],
             "speed": [round(float(speed[2 * k]), 1), round(float(speed[min(2 * k + 1, B - 1)]), 1)]} for k in order]
    q = np.quantile(per[:, 0], [0.5, 0.9, 0.99, 1.0])
    np.save(f"gpurun_out/phase_per_warp_{name}_{regime}.npy", per)
    np.save(f"gpurun_out/phase_hits_{name}_{regime}.npy", hist)
    out[regime] = {"kernel_ms_with_timers": round(a.elapsed_time(b), 4),
                   "share_of_warp_cycles": {PHASES[k]: round(float(cyc[k] / cyc.sum()), 4) for k in range(8)},
                   "mean_warp_cycles": round(float(cyc.sum() / n_warps), 1),
                   "warp_cycles_quantiles_50_90_99_100": [int(v) for v in q], "slowest_warps": slow,
                   "corr_cycles_vs": {"rounds": round(float(np.corrcoef(per[:, 0], per[:, 1])[0, 1]), 3),
                                      "refresh_passes": round(float(np.corrcoef(per[:, 0], per[:, 2])[0, 1]), 3),
                                      "hits": round(float(np.corrcoef(per[:, 0], hist[0:2 * n_warps:2] + hist[1:2 * n_warps:2])[0, 1]), 3),
                                      "speed": round(float(np.corrcoef(per[:, 0], speed[0:2 * n_warps:2] + speed[1:2 * n_warps:2])[0, 1]), 3)},
                   "rounds_per_warp": round(float(raw[8] / n_warps), 2), "refresh_passes_per_warp": round(float(raw[9] / n_warps), 2)}
print(json.dumps(out, indent=1))
Path("gpurun_out").mkdir(exist_ok=True)
Path(f"gpurun_out/phase_timers_{name}.json").write_text(json.dumps(out))
