"""world_size-2 gloo tests (CPU) of the N>1 host logic: ball sharding and the one exchange step
(policy-gradient + loss all-reduce).  The per-shard compute is the oracle here (no GPU in this
container); on the B200 box the same logic drives the CUDA path (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from doper_b200 import synthetic as syn
from doper_b200.distributed import PolicyGradReducer, ball_shard


def test_ball_shard_partitions_exactly():
    for n, world in ((4096, 8), (65536, 3), (10, 4), (7, 8), (1048576, 8)):
        pieces = [ball_shard(n, r, world) for r in range(world)]
        assert pieces[0][0] == 0 and pieces[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
        sizes = [hi - lo for lo, hi in pieces]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        ball_shard(10, 4, 4)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_balls, n_steps, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import ball_sim_np as onp
        from oracle.cref import CRef
        torch.manual_seed(0)  # identical controller replicas
        ctrl = torch.nn.Sequential(torch.nn.Linear(4, 50), torch.nn.ReLU(), torch.nn.Linear(50, 2))
        w = syn.make_workload("c2", n_balls=n_balls)  # global arrays, same on every rank
        lo, hi = ball_shard(n_balls, rank, world)
        x0, v_base = w["x0"][lo:hi], w["v0"][lo:hi]
        c = onp.make_constants(w["constants"])
        obs = torch.tensor(np.concatenate([x0, v_base], axis=1))
        impulse = ctrl(obs)
        v0 = torch.tensor(v_base) + impulse / float(c.mass)  # ball_trainer.py:61-63
        g = CRef("portable").value_and_grad(n_steps, np.float32((n_steps / 1000) / n_steps), c, w["attractor"],
                                            w["target"], w["segments"], x0, v0.detach().numpy())
        impulse.backward(torch.tensor(g["grad_v0"]) / float(c.mass))  # ball_trainer.py:70-71 (chain rule)
        red = PolicyGradReducer(ctrl.parameters())
        loss = red.reduce(float(g["loss_per_ball"].sum(dtype=np.float64)))
        if rank == 0:
            out.put((loss.item(), [p.grad.clone().numpy() for p in ctrl.parameters()]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_policy_gradient_equals_single_process():
    n_balls, n_steps = 96, 120
    ctx = mp.get_context("spawn")
    results = {}
    for world in (1, 2):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, n_balls, n_steps, q)) for r in range(world)]
        for p in procs:
            p.start()
        results[world] = q.get(timeout=240)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    (l1, g1), (l2, g2) = results[1], results[2]
    assert abs(l1 - l2) <= 1e-5 * abs(l1)
    for a, b in zip(g1, g2):
        np.testing.assert_allclose(a, b, rtol=2e-5, atol=1e-6 * max(1.0, float(np.abs(a).max())))
