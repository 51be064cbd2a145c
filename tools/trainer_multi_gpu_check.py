"""torchrun --nproc-per-node N tools/trainer_multi_gpu_check.py : the data-parallel training iteration.

Every rank owns 16 balls of its own; the iteration (8 actions, clip, Adam) runs (a) eagerly with the
NCCL exchange, (b) as ONE CUDA graph per rank with the peer-memory exchange captured inside, and (c) on
one GPU with all N x 16 balls and no exchange.  (a) = (b) (parameters after 3 iterations), (b) is
bit-identical across ranks, and the first iteration's summed gradient equals (c)'s."""
import copy
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from doper_b200 import synthetic as syn  # noqa: E402
from doper_b200.agents import RangeSensingAgent  # noqa: E402
from doper_b200.scenes import SingleScene  # noqa: E402
from doper_b200.sim.jax_geometry import JaxScene  # noqa: E402
from doper_b200.trainers import BallAttractorTrainer  # noqa: E402

CONFIG = {
    "constants": {"radius": 0.1, "density": 3000, "rolling_friction_coefficient": 0.007,
                  "walls_elasticity": 0.8, "attractor_strength": 0.01},
    "sim": {"n_steps": 300, "sim_time": 0.3, "coordinate_target": [6.0, 4.0], "attractor_coordinate": [-3.0, 2.0],
            "scene_params": {"svg_scene_path": "unused", "px_per_meter": 1},
            "agent_params": {"distance_range": 1000, "angle_step": 20}},
    "train": {"device": "cuda", "batch_size": 16, "num_actions": 8},
    "model": {"input_dim": 25, "output_dim": 2, "hidden_dim": 50},
}

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = 16
cfg = copy.deepcopy(CONFIG)
cfg["train"]["device"] = f"cuda:{local}"
handler = SingleScene(cfg, jax_scene=JaxScene(syn.make_map(1002, 67)))
agent = RangeSensingAgent(cfg)
g = torch.Generator(device=dev); g.manual_seed(11 + rank)
draws = [(handler.get_init_state(B, generator=g), -1 + 3 * torch.rand((B, 2), device=dev, generator=g)) for _ in range(3)]


def trainer(transport, batch=B, data_parallel=True):
    torch.manual_seed(4)  # the same controller on every rank
    c = copy.deepcopy(cfg)
    c["train"]["batch_size"] = batch
    t = BallAttractorTrainer(c)
    t.reduce_transport, t.data_parallel = transport, data_parallel
    return t


# (a) eager, NCCL
eager = trainer("collective")
losses_a, grads_a = [], None
for x0, v0 in draws:
    eager.optimizer.zero_grad(set_to_none=True)
    losses_a.append(float(eager._action_loop(agent, handler, x0, v0, 5)))
    if grads_a is None:
        grads_a = [p.grad.detach().clone() for p in eager.controller.parameters()]

# (b) one CUDA graph per rank, peer-memory exchange inside
graphed = trainer("nvlink")
run = graphed.graphed_iteration(agent, handler)
losses_b = [float(run(x0, v0)) for x0, v0 in draws]
same_as_eager = all(torch.allclose(a, b, rtol=1e-5, atol=1e-7)
                    for a, b in zip(eager.controller.parameters(), graphed.controller.parameters()))
same_as_eager = same_as_eager and all(abs(a - b) <= 1e-6 * abs(a) for a, b in zip(losses_a, losses_b))
bits = True
for p in graphed.controller.parameters():
    q = p.detach().clone()
    dist.broadcast(q, 0)
    bits = bits and bool(torch.equal(q, p.detach()))

# (c) all balls on one GPU, no exchange: the first iteration's (clipped) gradient
xs = [torch.empty_like(draws[0][0]) for _ in range(world)]
vs = [torch.empty_like(draws[0][1]) for _ in range(world)]
dist.all_gather(xs, draws[0][0].contiguous())
dist.all_gather(vs, draws[0][1].contiguous())
big = trainer("collective", batch=B * world, data_parallel=False)
big.optimizer.zero_grad(set_to_none=True)
loss_c = float(big._action_loop(agent, handler, torch.cat(xs), torch.cat(vs), 5))
grad_err = 0.0
for a, c in zip(grads_a, [p.grad for p in big.controller.parameters()]):
    grad_err = max(grad_err, float((a - c).abs().max() / c.abs().max().clamp_min(1e-30)))
loss_err = abs(loss_c - losses_a[0]) / abs(loss_c)


def timeit(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier(device_ids=[local])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


# (d) the fused action loop (controller / clip / Adam as library kernels), one graph per rank, the flat gradient
# buffer exchanged in place by the peer kernel: parameters close to (a), identical bits on every rank
fusedt = trainer("nvlink")
runf = fusedt.graphed_iteration(agent, handler, fused=True)
losses_d = [float(runf(x0, v0)) for x0, v0 in draws]
fused_close = all(torch.allclose(a, b, rtol=1e-3, atol=2e-5) for a, b in zip(eager.controller.parameters(), fusedt.controller.parameters()))
fused_close = fused_close and all(abs(a - b) <= 1e-3 * abs(a) for a, b in zip(losses_a, losses_d))
fbits = True
q = fusedt.fused_loop().pflat.detach().clone()
dist.broadcast(q, 0)
fbits = bool(torch.equal(q, fusedt.fused_loop().pflat))

x0, v0 = draws[0]


def eager_it():
    eager.optimizer.zero_grad(set_to_none=True)
    eager._action_loop(agent, handler, x0, v0, 5)


t_eager, t_graph = timeit(eager_it), timeit(lambda: run(x0, v0))
t_fused = timeit(lambda: runf(x0, v0))
status = int(graphed._reducer.peer.status.item())
flags = torch.tensor([int(same_as_eager), int(bits), int(fused_close), int(fbits)], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    line = {"world": world, "balls_per_rank": B, "graph_equals_eager_nccl": bool(flags[0]), "same_bits_on_every_rank": bool(flags[1]),
            "first_iteration_grad_vs_single_gpu_big_batch_rel_err": grad_err, "loss_rel_err": loss_err, "peer_status": status,
            "fused_graph_close_to_eager_nccl": bool(flags[2]), "fused_same_bits_on_every_rank": bool(flags[3]),
            "eager_nccl_ms": round(t_eager, 4), "graph_peer_ms": round(t_graph, 4), "fused_graph_peer_ms": round(t_fused, 4)}
    print(json.dumps(line), flush=True)
    Path("gpurun_out").mkdir(exist_ok=True)
    Path(f"gpurun_out/trainer_multi_gpu_n{world}.json").write_text(json.dumps(line))
graphed._reducer.peer.close()
if fusedt.fused_loop()._peer:
    fusedt.fused_loop()._peer.close()
dist.destroy_process_group()
