"""How much of a step is the serial chain of the few balls in sliding contact?  Times forward and
backward of a workload as is, then with the balls that collided on more than 10 % of their steps
replaced by copies of an ordinary ball."""
import ctypes
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from doper_b200 import synthetic as syn  # noqa: E402
from doper_b200._abi import check, lib  # noqa: E402
from doper_b200._device import ptr  # noqa: E402
from doper_b200.sim import ball_sim  # noqa: E402
from doper_b200.sim.jax_geometry import JaxScene  # noqa: E402


def time_ms(fn, reps=7, warm=2):
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
    name = sys.argv[1] if len(sys.argv) > 1 else "c5"
    balls = int(sys.argv[2]) if len(sys.argv) > 2 else None
    w = syn.make_workload(name, n_balls=balls)
    B, n = w["x0"].shape[0], w["n_steps"]
    dev = torch.device("cuda", 0)
    scene = JaxScene(w["segments"])
    _, sws = scene.prepared(dev, 0.1)
    st = torch.cuda.current_stream().cuda_stream
    tgt = ball_sim._target2(w["target"])
    desc = ball_sim._make_desc(B, scene, w["sim_time"], n, w["attractor"], w["constants"], ball_sim.RolloutOptions())
    bufs = ball_sim._Buffers.get(dev, desc)
    xT = torch.empty((B, 2), device=dev); vT = torch.empty((B, 2), device=dev)
    bar = torch.empty((B, 2), device=dev); gv = torch.empty((B, 2), device=dev); gx = torch.empty((B, 2), device=dev)
    ls = torch.empty(1, device=dev)
    out = {"workload": name, "B": B, "n_steps": n}
    x0h, v0h = w["x0"].copy(), w["v0"].copy()
    # optional third argument: hit-count thresholds, e.g. "8,4,2,0": after the run as is, balls with more
    # collisions than each threshold are replaced by copies of a ball that never collides
    thresholds = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else None
    tags = ("as_is", "without_sliding_balls") if thresholds is None else ("as_is",) + tuple(f"max_{h}_collisions_per_ball" for h in thresholds)
    per_ball = None
    for tag in tags:
        x0, v0 = torch.tensor(x0h, device=dev), torch.tensor(v0h, device=dev)
        fwd = lambda: check(lib.doper_rollout_fwd(st, ctypes.byref(desc), ptr(sws), ptr(x0), ptr(v0), ptr(xT), ptr(vT),
                                                  0, 0, 0, ptr(bufs.ws)), "fwd")
        f_ms = time_ms(fwd)
        check(lib.doper_l1_loss(st, B, ptr(xT), tgt, 0, ptr(ls), ptr(bar)), "loss")
        b_ms = time_ms(lambda: check(lib.doper_rollout_bwd(st, ctypes.byref(desc), ptr(sws), ptr(bufs.ws), ptr(bufs.bwd_scratch), ptr(bar), 0,
                                                           ptr(gx), ptr(gv)), "bwd"))
        out[tag] = {"fwd_ms": round(f_ms, 4), "bwd_ms": round(b_ms, 4)}
        if tag != "as_is" and thresholds is not None:
            pass
        if tag == "as_is":
            h = ball_sim.rollout_history(w["sim_time"], n, scene, x0h, v0h, w["attractor"], w["constants"])["hit"]
            per_ball = h.sum(0)
            heavy = np.nonzero(per_ball > 0.1 * n)[0]
            out["balls_colliding_on_over_10pct_of_steps"] = int(len(heavy))
            out["max_collisions_per_ball"] = int(per_ball.max())
            out["total_collisions"] = int(per_ball.sum())
            donor = int(np.argmin(per_ball))
            if thresholds is None:
                x0h[heavy], v0h[heavy] = x0h[donor], v0h[donor]
        if thresholds is not None:
            k = tags.index(tag)
            if k < len(thresholds):
                sel = np.nonzero(per_ball > thresholds[k])[0]
                # donors: random balls that never collide (same speed distribution, unlike one fixed donor)
                free = np.nonzero(per_ball == 0)[0]
                pick = np.random.default_rng(7).choice(free, size=len(sel))
                x0h[sel], v0h[sel] = w["x0"][pick], w["v0"][pick]
                out.setdefault("replaced", {})[str(thresholds[k])] = int(len(sel))
        print(tag, out[tag], flush=True)
    print(json.dumps(out))
    Path("gpurun_out").mkdir(exist_ok=True)
    Path(f"gpurun_out/tail_probe_{name}_{B}.json").write_text(json.dumps(out))


if __name__ == "__main__":
    main()
