"""Per-kernel-call timing of ONE action of the fused trainer loop (CUDA events, median of 20): where the
iteration's time goes at a given batch size.  python tools/fused_breakdown.py [B]"""
import copy, ctypes, json, sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from doper_b200 import synthetic as syn
from doper_b200._abi import check, lib
from doper_b200._device import ptr, stream_ptr
from doper_b200.agents import RangeSensingAgent
from doper_b200.scenes import SingleScene
from doper_b200.sim.jax_geometry import JaxScene
from doper_b200.trainers import BallAttractorTrainer
from tools.bench_trainer import CONFIG

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = copy.deepcopy(CONFIG); cfg["train"]["batch_size"] = B
handler = SingleScene(cfg, jax_scene=JaxScene(syn.make_map(1002, 67)))
agent = RangeSensingAgent(cfg)
tr = BallAttractorTrainer(cfg)
fl = tr.fused_loop(5)
x = handler.get_init_state(B); v = -1 + 3 * torch.rand((B, 2), device="cuda")
for _ in range(3):
    fl.run(agent, handler, x, v)
pl = fl._plan(handler.jax_scene, B)
desc, sws, bufs = pl["desc"], pl["sws"], pl["bufs"]
st = stream_ptr()
loss = fl.gflat[fl.n:fl.n + 1]
obs = agent.get_observations(x, v, handler)
calls = {
    "observations (sensor kernel)": lambda: agent.get_observations(x, v, handler),
    "controller_fwd": lambda: check(lib.doper_controller_fwd(st, fl.dims, B, ptr(fl.pflat), ptr(obs), ptr(v), fl.mass32, 0, ptr(pl["v0"])), "f"),
    "rollout fwd": lambda: check(lib.doper_rollout_fwd(st, ctypes.byref(desc), ptr(sws), ptr(x), ptr(pl["v0"]), ptr(pl["x"][0]), ptr(pl["vT"]), 0, 0, 0, ptr(bufs.ws)), "r"),
    "l1 loss": lambda: check(lib.doper_l1_loss(st, B, ptr(pl["x"][0]), pl["tgt"], 0, ptr(loss), ptr(bufs.scratch)), "l"),
    "rollout bwd": lambda: check(lib.doper_rollout_bwd(st, ctypes.byref(desc), ptr(sws), ptr(bufs.ws), ptr(bufs.bwd_scratch), ptr(bufs.scratch), 0, 0, ptr(pl["g"])), "b"),
    "controller_bwd (+ reduce)": lambda: check(lib.doper_controller_bwd(st, fl.dims, B, ptr(fl.pflat), ptr(obs), ptr(pl["g"]), fl.mass32, ptr(pl["scratch"]), ptr(fl.gflat)), "cb"),
    "adam_clip_step": lambda: check(lib.doper_adam_clip_step(st, fl.n, ptr(fl.pflat), ptr(fl.gflat), ptr(fl.exp_avg), ptr(fl.exp_avg_sq), ptr(fl.step), 5.0, 1e-3, 0.9, 0.999, 1e-8, ptr(fl.total_norm)), "a"),
}
out = {"batch": B, "plan": lib.doper_rollout_plan_name(ctypes.byref(desc), 0).decode(), "us": {}}
for name, fn in calls.items():
    ts = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    out["us"][name] = round(1e3 * float(np.median(ts)), 1)
print(json.dumps(out, indent=1))
Path("gpurun_out").mkdir(exist_ok=True)
Path(f"gpurun_out/fused_breakdown_{B}.json").write_text(json.dumps(out))
