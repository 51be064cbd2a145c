import sys, ctypes, torch, numpy as np
sys.path.insert(0, "/root/repo")
from doper_b200 import synthetic as syn
from doper_b200.sim import ball_sim
from doper_b200.sim.jax_geometry import JaxScene
def t(w, tag):
    scene = JaxScene(w["segments"])
    args = (w["sim_time"], w["n_steps"], scene, torch.tensor(w["x0"], device="cuda"), torch.tensor(w["v0"], device="cuda"), w["target"], w["attractor"], w["constants"])
    for _ in range(3): ball_sim.vmapped_grad_and_value(*args)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): ball_sim.vmapped_grad_and_value(*args)
    b.record(); torch.cuda.synchronize()
    h = ball_sim.rollout_history(w["sim_time"], w["n_steps"], scene, w["x0"], w["v0"], w["attractor"], w["constants"])
    print(tag, "ms/step", a.elapsed_time(b) / 20, "hits", int(h["hit"].sum()), "max hits/ball", int(h["hit"].sum(0).max()), "speed max", float(np.linalg.norm(w["v0"], axis=1).max()))
t(syn.make_workload("c2"), "c2 native 4096")
t(syn.make_workload("c2", n_balls=8192, lo=0, hi=4096), "slice 0 of 8192")
t(syn.make_workload("c2", n_balls=8192, lo=4096, hi=8192), "slice 1 of 8192")
