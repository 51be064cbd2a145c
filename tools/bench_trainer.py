"""Training iteration (8 actions: observation -> controller -> rollout fwd+bwd -> controller bwd; clip;
Adam), eager vs one CUDA graph.  Prints one JSON line."""
import copy
import json
import sys
import time
from pathlib import Path

import torch

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


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / reps


def main():
    out = {"metric": "training_iteration_ms", "config": "configs/jax_ball.yaml shapes: 300 steps, 8 actions, 201-segment map", "rows": []}
    for B in (16, 4096):
        cfg = copy.deepcopy(CONFIG)
        cfg["train"]["batch_size"] = B
        handler = SingleScene(cfg, jax_scene=JaxScene(syn.make_map(1002, 67)))
        agent = RangeSensingAgent(cfg)
        torch.manual_seed(0)  # same controller, same draw in every run: the iteration time depends on the trajectories
        tr = BallAttractorTrainer(cfg)
        gen = torch.Generator(device="cuda"); gen.manual_seed(1)
        x0 = handler.get_init_state(B, generator=gen)
        v0 = -1 + 3 * torch.rand((B, 2), device="cuda", generator=gen)

        def eager():
            tr.optimizer.zero_grad(set_to_none=True)
            tr._action_loop(agent, handler, x0, v0, 5)

        e = timeit(eager, 20)
        run = tr.graphed_iteration(agent, handler)
        g = timeit(lambda: run(x0, v0), 50)
        # the fused action loop (controller / clip / Adam as kernels of this library), eager and as one graph
        torch.manual_seed(0)
        tr2 = BallAttractorTrainer(copy.deepcopy(cfg))
        fl = tr2.fused_loop(5)
        fe = timeit(lambda: fl.run(agent, handler, x0, v0), 20)
        run2 = tr2.graphed_iteration(agent, handler, fused=True)
        fg = timeit(lambda: run2(x0, v0), 50)
        out["rows"].append({"batch": B, "eager_ms": round(e, 4), "graph_ms": round(g, 4),
                            "fused_eager_ms": round(fe, 4), "fused_graph_ms": round(fg, 4),
                            "ball_steps_per_s_graph": round(B * 300 * 8 / (g * 1e-3), 1),
                            "ball_steps_per_s_fused_graph": round(B * 300 * 8 / (fg * 1e-3), 1)})
    print(json.dumps(out))
    Path("gpurun_out").mkdir(exist_ok=True)
    Path("gpurun_out/trainer_bench.json").write_text(json.dumps(out))


if __name__ == "__main__":
    main()
