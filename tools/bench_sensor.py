"""Measurement of the range-sensor row: RangeSensingAgent.get_observations for
the C2 batch (4,096 balls x 18 rays x 201 segments), CUDA-event timed, against the numpy oracle on
the host.  Prints one JSON line (also written to gpurun_out/sensor_bench.json)."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from doper_b200 import synthetic as syn  # noqa: E402
from doper_b200.agents import RangeSensingAgent  # noqa: E402
from doper_b200.sim.jax_geometry import JaxScene  # noqa: E402
from oracle import ray_np as orc  # noqa: E402  (cpu baseline leg only)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    w = syn.make_workload(name)
    cfg = {"sim": {"coordinate_target": [6.0, 4.0], "agent_params": {"distance_range": 1000, "angle_step": 20}}}
    agent = RangeSensingAgent(cfg)
    scene = JaxScene(w["segments"])
    dev = torch.device("cuda", 0)
    x = torch.tensor(w["x0"], device=dev)
    v = torch.tensor(w["v0"], device=dev)
    for _ in range(5):
        obs = agent.get_observations(x, v, scene)
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    K = 50
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for a, b in ev:
        flush.fill_(1)
        a.record(); obs = agent.get_observations(x, v, scene); b.record()
    torch.cuda.synchronize()
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    B, R, S = x.shape[0], obs.shape[1] - 7, int(w["segments"].shape[0])
    # e2e: host arrays in, host tensor out
    for _ in range(3):
        agent.get_observations(w["x0"], w["v0"], scene)
    t0 = time.perf_counter()
    for _ in range(20):
        agent.get_observations(w["x0"], w["v0"], scene)
    e2e_ms = 1e3 * (time.perf_counter() - t0) / 20
    # cpu baseline: the numpy oracle (vectorised over segments, python loop over balls like the reference)
    n_cpu = 256
    dirs = orc.ray_directions((1, 0), 360, 20).astype(np.float32)
    t0 = time.perf_counter()
    orc.agent_observations(w["x0"][:n_cpu], w["v0"][:n_cpu], dirs, w["segments"], 1000.0, [6.0, 4.0])
    cpu_s = time.perf_counter() - t0
    pairs = B * R * S
    flops_per_pair = 35  # reference arithmetic per (ray, segment) pair incl. its per-pair normalisation and the range
    line = {
        "metric": "sensor_rays_per_sec", "value": round(B * R / (ms * 1e-3), 1), "unit": "rays/s",
        "config": {"workload": f"{name}: RangeSensingAgent.get_observations, {B} balls x {R} rays x {S} segments -> obs [{B},{7 + R}]"},
        "ms_per_call": round(ms, 5), "pairs_per_sec": round(pairs / (ms * 1e-3), 1),
        "e2e": {"value": round(B * R / (e2e_ms * 1e-3), 1), "unit": "rays/s", "ms_per_call": round(e2e_ms, 5),
                "h2d_bytes_per_call": int(B * 4 * 4), "d2h_bytes_per_call": int(B * (7 + R) * 4)},
        "roofline": {"bound": "fp32", "achieved_tflops": round(pairs * flops_per_pair / (ms * 1e-3) / 1e12, 3),
                     "note": "algorithmic flops of the reference per pair (35); most pairs are rejected after 12 "
                             "instructions by a conservative sign / range test, the survivors get the exact divisions"},
        "cpu_baseline": {"value": round(n_cpu * R / cpu_s, 1), "unit": "rays/s", "kind": "port", "cores": 1,
                         "sample": f"first {n_cpu} balls, oracle/ray_np.py (numpy, python loop over balls like "
                                   f"range_sensor_agent.py:36-39)"},
    }
    print(json.dumps(line))
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/sensor_bench.json").write_text(json.dumps(line))


if __name__ == "__main__":
    main()
