"""Forward-kernel A/B timing under gpurun: default plan vs full-sweep kernels vs the broad-phase
kernels at every lane count.

    python tools/tune_fwd.py c2 [n_balls] [lo hi]
"""
import ctypes
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from doper_b200 import synthetic as syn  # noqa: E402
from doper_b200._abi import check, lib  # noqa: E402
from doper_b200._device import ptr  # noqa: E402
from doper_b200.sim import ball_sim  # noqa: E402
from doper_b200.sim.jax_geometry import JaxScene  # noqa: E402


def time_ms(fn, reps=15, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2]


def main():
    pos = [a for a in sys.argv[1:] if "=" not in a]
    name = pos[0] if pos else "c2"
    balls = int(pos[1]) if len(pos) > 1 else None
    lo = int(pos[2]) if len(pos) > 3 else 0
    hi = int(pos[3]) if len(pos) > 3 else None
    w = syn.make_workload(name, n_balls=balls, lo=lo, hi=hi)
    B, n = w["x0"].shape[0], w["n_steps"]
    dev = torch.device("cuda", 0)
    scene = JaxScene(w["segments"])
    _, sws = scene.prepared(dev, 0.1)
    x0, v0 = torch.tensor(w["x0"], device=dev), torch.tensor(w["v0"], device=dev)
    st = torch.cuda.current_stream().cuda_stream
    tgt = ball_sim._target2(w["target"])
    out = {"workload": name, "B": B, "n_steps": n, "S": int(w["segments"].shape[-3]), "fwd_ms": {}}
    variants = [("default", {}), ("nocull", dict(broad_phase=False))]
    variants += [("tp1", dict(smem_forward=True))]
    variants += [(f"cull_G{g}", dict(lanes_per_ball=g, broad_phase=True)) for g in (1, 2, 4, 8, 16, 32)]
    variants += [(f"tpc_G{g}", dict(lanes_per_ball=g, broad_phase=True, broad_phase_time_parallel=True)) for g in (2, 4, 8, 16, 32)]
    only = [a[5:] for a in sys.argv if a.startswith("only=")]
    if only:
        variants = [v for v in variants if any(v[0].startswith(o) for o in only[0].split(","))]
    for tag, kw in variants:
        opts = ball_sim.RolloutOptions(**kw)
        desc = ball_sim._make_desc(B, scene, w["sim_time"], n, w["attractor"], w["constants"], opts)
        bufs = ball_sim._Buffers.get(dev, desc)
        xT = torch.empty((B, 2), device=dev); vT = torch.empty((B, 2), device=dev)
        bar = torch.empty((B, 2), device=dev); gv = torch.empty((B, 2), device=dev); gx = torch.empty((B, 2), device=dev)
        ls = torch.empty(1, device=dev)
        try:
            ms = time_ms(lambda: check(lib.doper_rollout_fwd(st, ctypes.byref(desc), ptr(sws), ptr(x0), ptr(v0),
                                                              ptr(xT), ptr(vT), 0, 0, 0, ptr(bufs.ws)), "fwd"))
            check(lib.doper_l1_loss(st, B, ptr(xT), tgt, 0, ptr(ls), ptr(bar)), "loss")
            bms = time_ms(lambda: check(lib.doper_rollout_bwd(st, ctypes.byref(desc), ptr(sws), ptr(bufs.ws), ptr(bufs.bwd_scratch), ptr(bar),
                                                               0, ptr(gx), ptr(gv)), "bwd"))
        except Exception as e:
            ms, bms = repr(e), None
        out["fwd_ms"][tag] = ms
        out.setdefault("bwd_ms", {})[tag] = bms
        print(f"{name} B={B} {tag:14s} fwd {ms} ms   bwd {bms} ms", flush=True)
    Path("gpurun_out").mkdir(exist_ok=True)
    Path(f"gpurun_out/tune_fwd_{name}_{B}_{lo}.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
