"""A few iterations of the fused trainer loop at a fixed, seeded draw (the command ncu wraps).  python tools/fused_iter.py [B] [iters]"""
import copy, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from doper_b200 import synthetic as syn
from doper_b200.agents import RangeSensingAgent
from doper_b200.scenes import SingleScene
from doper_b200.sim.jax_geometry import JaxScene
from doper_b200.trainers import BallAttractorTrainer
from tools.bench_trainer import CONFIG
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
cfg = copy.deepcopy(CONFIG); cfg["train"]["batch_size"] = B
handler = SingleScene(cfg, jax_scene=JaxScene(syn.make_map(1002, 67)))
agent = RangeSensingAgent(cfg)
tr = BallAttractorTrainer(cfg)
fl = tr.fused_loop(5)
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = handler.get_init_state(B, generator=g); v = -1 + 3 * torch.rand((B, 2), device="cuda", generator=g)
for _ in range(iters):
    loss = fl.run(agent, handler, x, v)
torch.cuda.synchronize()
print("ok", float(loss))
