"""GPU tuning / accuracy report (run under gpurun): forward time vs lanes-per-ball, backward
time chunk-parallel vs sequential, gradient error of the chunk-parallel backward."""
import ctypes
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from doper_b200 import synthetic as syn  # noqa: E402
from doper_b200._abi import check, lib  # noqa: E402
from doper_b200._device import ptr  # noqa: E402
from doper_b200.sim import ball_sim  # noqa: E402
from doper_b200.sim.jax_geometry import JaxScene  # noqa: E402


def time_ms(fn, reps=20, warm=3):
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
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    balls = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else None
    w = syn.make_workload(name, n_balls=balls)
    B, n = w["x0"].shape[0], w["n_steps"]
    dev = torch.device("cuda", 0)
    scene = JaxScene(w["segments"])
    _, sws = scene.prepared(dev, 0.1)
    x0, v0 = torch.tensor(w["x0"], device=dev), torch.tensor(w["v0"], device=dev)
    st = torch.cuda.current_stream().cuda_stream
    tgt = ball_sim._target2(w["target"])
    out = {"workload": name, "B": B, "n_steps": n, "S": int(w["segments"].shape[-3])}
    lanes_list = [0, 1, 2, 4, 8, 16, 32] if not w["per_env"] else [0, 2, 4, 8, 16, 32]
    if len(sys.argv) > 3:
        lanes_list = [int(v) for v in sys.argv[3].split(',')]
    for exact, smem in ((False, False), (False, True)):
        for lanes in lanes_list:
            opts = ball_sim.RolloutOptions(lanes_per_ball=lanes, force_exact_sweep=exact, smem_forward=smem)
            desc = ball_sim._make_desc(B, scene, w["sim_time"], n, w["attractor"], w["constants"], opts)
            bufs = ball_sim._Buffers.get(dev, desc)
            xT = torch.empty((B, 2), device=dev); vT = torch.empty((B, 2), device=dev)
            try:
                ms = time_ms(lambda: check(lib.doper_rollout_fwd(st, ctypes.byref(desc), ptr(sws), ptr(x0), ptr(v0),
                                                                  ptr(xT), ptr(vT), 0, 0, 0, ptr(bufs.ws)), "fwd"))
            except Exception as e:
                ms = repr(e)
            out[f"fwd_ms_lanes{lanes}{'_exact' if exact else ''}{'_smem' if smem else ''}"] = ms
            print(f"fwd lanes={lanes:2d} exact={exact} smem={smem}: {ms}", flush=True)
    # backward variants
    res = {}
    for seq in (True, False):
        opts = ball_sim.RolloutOptions(sequential_backward=seq)
        desc = ball_sim._make_desc(B, scene, w["sim_time"], n, w["attractor"], w["constants"], opts)
        K = lib.doper_rollout_ckpt_every(ctypes.byref(desc))
        ws = torch.empty(lib.doper_rollout_workspace_bytes(ctypes.byref(desc)), dtype=torch.uint8, device=dev)
        xT = torch.empty((B, 2), device=dev); vT = torch.empty((B, 2), device=dev)
        bar = torch.empty((B, 2), device=dev); gv = torch.empty((B, 2), device=dev); gx = torch.empty((B, 2), device=dev)
        ls = torch.empty(1, device=dev)
        check(lib.doper_rollout_fwd(st, ctypes.byref(desc), ptr(sws), ptr(x0), ptr(v0), ptr(xT), ptr(vT), 0, 0, 0,
                                    ptr(ws)), "fwd")
        check(lib.doper_l1_loss(st, B, ptr(xT), tgt, 0, ptr(ls), ptr(bar)), "loss")
        ms = time_ms(lambda: check(lib.doper_rollout_bwd(st, ctypes.byref(desc), ptr(sws), ptr(ws), ptr(bar), 0,
                                                          ptr(gx), ptr(gv)), "bwd"))
        lms = time_ms(lambda: check(lib.doper_l1_loss(st, B, ptr(xT), tgt, 0, ptr(ls), ptr(bar)), "loss"))
        res[seq] = gv.cpu().numpy().astype(np.float64)
        out[f"bwd_ms_{'sequential' if seq else 'chunked'}"] = ms
        out[f"K_{'sequential' if seq else 'chunked'}"] = int(K)
        out["loss_ms"] = lms
        print(f"bwd sequential={seq} K={K}: {ms} ms; loss {lms} ms", flush=True)
    nrm = np.linalg.norm(res[True], axis=1)
    err = np.abs(res[False] - res[True]).max(axis=1) / np.maximum(nrm, 1e-30)
    out["chunked_vs_sequential_relnorm"] = {"max": float(err.max()), "q999": float(np.quantile(err, 0.999)),
                                            "q99": float(np.quantile(err, 0.99)), "median": float(np.median(err)),
                                            "max_abs_grad": float(np.abs(res[True]).max())}
    print(json.dumps(out))
    Path("gpurun_out").mkdir(exist_ok=True)
    Path(f"gpurun_out/tune_{name}.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
