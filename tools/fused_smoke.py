import numpy as np, torch
from doper_b200 import synthetic as syn
from doper_b200.sim import ball_sim
w = syn.make_workload("c2", n_balls=64)
n = 100
(l, (x, v)), g = ball_sim.vmapped_grad_and_value(n/1000, n, w["segments"], w["x0"], w["v0"], w["target"], w["attractor"], w["constants"])
print("ok", float(l), np.isfinite(g).all())
