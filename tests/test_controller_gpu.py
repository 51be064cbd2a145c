"""GPU tests of the trainer's glue kernels (csrc/controller.cuh): controller forward /
backward against torch (doper/networks/controllers.py:12-18), gradient clip + Adam against
torch.nn.utils.clip_grad_norm_ + torch.optim.Adam (doper/trainers/ball_trainer.py:74-76), and the fused action
loop against the torch action loop on the same draws.  fp32 both sides; the summation orders differ, so the bar
is rounding-level agreement, stated per test."""
import copy
import ctypes

import numpy as np
import pytest
import torch

from doper_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

from tests.test_trainer_gpu import CONFIG  # noqa: E402


def _flat(ctrl):
    return torch.cat([p.detach().reshape(-1) for p in ctrl.parameters()]).contiguous()


@pytest.mark.parametrize("B", [1, 16, 33, 4096])
def test_controller_forward_and_backward_match_torch(B):
    from doper_b200._abi import check, lib
    from doper_b200._device import ptr, stream_ptr
    from doper_b200.networks import LinearReLUController
    torch.manual_seed(B)
    ctrl = LinearReLUController(CONFIG["model"]).cuda()
    dims = (ctypes.c_int32 * 4)(25, 50, 100, 2)
    obs = torch.randn(B, 25, device="cuda") * 3
    v_init = torch.randn(B, 2, device="cuda")
    mass = float(np.float32(12.566370614359174))
    pflat = _flat(ctrl)
    assert pflat.numel() == 6602
    imp = torch.empty(B, 2, device="cuda"); v0 = torch.empty(B, 2, device="cuda")
    check(lib.doper_controller_fwd(stream_ptr(), dims, B, ptr(pflat), ptr(obs), ptr(v_init), mass, ptr(imp), ptr(v0)), "fwd")
    ref_imp = ctrl(obs)
    ref_v0 = v_init + ref_imp / mass
    assert torch.allclose(imp, ref_imp, rtol=2e-5, atol=2e-6)
    assert torch.allclose(v0, ref_v0, rtol=2e-5, atol=2e-6)
    g = torch.randn(B, 2, device="cuda")
    (ref_v0 * g).sum().backward()
    ref_grad = torch.cat([p.grad.reshape(-1) for p in ctrl.parameters()])
    grad = torch.full((6602 + 2,), 0.25, device="cuda")  # the kernel ADDS (8 actions accumulate)
    scratch = torch.empty(lib.doper_controller_scratch_bytes(dims, B), dtype=torch.uint8, device="cuda")
    check(lib.doper_controller_bwd(stream_ptr(), dims, B, ptr(pflat), ptr(obs), ptr(g), mass, ptr(scratch), ptr(grad)), "bwd")
    got = grad[:6602] - 0.25
    scale = ref_grad.abs().max()
    assert float((got - ref_grad).abs().max()) <= 2e-5 * float(scale) + 1e-6, float((got - ref_grad).abs().max() / scale)
    assert float(grad[6602]) == 0.25  # the loss slot is not touched
    # deterministic: block partials are added in block order
    grad2 = torch.full((6602 + 2,), 0.25, device="cuda")
    check(lib.doper_controller_bwd(stream_ptr(), dims, B, ptr(pflat), ptr(obs), ptr(g), mass, ptr(scratch), ptr(grad2)), "bwd")
    assert torch.equal(grad, grad2)


def test_adam_clip_step_matches_torch():
    from doper_b200._abi import check, lib
    from doper_b200._device import ptr, stream_ptr
    torch.manual_seed(0)
    n = 6602
    p0 = torch.randn(n, device="cuda") * 0.1
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref_p])
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    step, tn = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    for it, gscale in enumerate([30.0, 0.01, 3.0, 100.0]):  # clipped, not clipped, ...
        g = torch.randn(n, device="cuda") * gscale
        ref_p.grad = g.clone()
        ref_norm = torch.nn.utils.clip_grad_norm_([ref_p], 5)
        opt.step()
        grad = torch.cat([g, torch.tensor([123.0, 0.0], device="cuda")])
        check(lib.doper_adam_clip_step(stream_ptr(), n, ptr(p), ptr(grad), ptr(m), ptr(v), ptr(step), 5.0, 1e-3, 0.9, 0.999, 1e-8,
                                       ptr(tn)), "adam")
        assert abs(float(tn) - float(ref_norm)) <= 1e-5 * float(ref_norm)
        assert torch.allclose(p, ref_p.detach(), rtol=1e-5, atol=1e-7), (it, float((p - ref_p.detach()).abs().max()))
        assert float(grad[:n].abs().max()) == 0.0 and float(grad[n]) == 123.0  # gradients zeroed, loss slot kept
        assert float(step) == it + 1


def _setup(batch, num_actions=8, n_steps=300):
    from doper_b200.agents import RangeSensingAgent
    from doper_b200.scenes import SingleScene
    from doper_b200.sim.jax_geometry import JaxScene
    cfg = copy.deepcopy(CONFIG)
    cfg["train"]["batch_size"] = batch
    cfg["train"]["num_actions"] = num_actions
    cfg["sim"]["n_steps"], cfg["sim"]["sim_time"] = n_steps, n_steps / 1000
    handler = SingleScene(cfg, jax_scene=JaxScene(syn.make_map(1002, 67)))
    agent = RangeSensingAgent(cfg)
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    draws = [(handler.get_init_state(batch, generator=g), -1 + 3 * torch.rand((batch, 2), device="cuda", generator=g))
             for _ in range(3)]
    return cfg, handler, agent, draws


@pytest.mark.parametrize("batch", [16, 300])
def test_fused_action_loop_matches_the_torch_action_loop(batch):
    """Same draws, same initial controller: losses and parameters after three iterations agree to rounding (the two
    controllers differ in the last bits of v0, the rollout is a chaotic map of v0: 1e-3 on the loss, 1e-4 on the
    parameters is what the torch loop itself shows between two GEMM algorithms)."""
    from doper_b200.trainers import BallAttractorTrainer
    cfg, handler, agent, draws = _setup(batch)
    torch.manual_seed(4)
    eager = BallAttractorTrainer(copy.deepcopy(cfg))
    losses_e = []
    for x0, v0 in draws:
        eager.optimizer.zero_grad(set_to_none=True)
        losses_e.append(float(eager._action_loop(agent, handler, x0, v0, 5)))
    torch.manual_seed(4)
    fused = BallAttractorTrainer(copy.deepcopy(cfg))
    fl = fused.fused_loop(5)
    losses_f = [float(fl.run(agent, handler, x0, v0)) for x0, v0 in draws]
    assert np.allclose(losses_e, losses_f, rtol=1e-3), (losses_e, losses_f)
    for a, b in zip(eager.controller.parameters(), fused.controller.parameters()):
        assert torch.allclose(a, b, rtol=1e-3, atol=2e-5), float((a - b).abs().max())
    # the reference module sees the fused updates (its parameters are views into the flat vector)
    assert torch.equal(torch.cat([p.detach().reshape(-1) for p in fused.controller.parameters()]), fl.pflat)
    assert float(fl.step) == 3.0


def test_fused_iteration_as_one_graph_is_bit_identical_to_eager_fused():
    from doper_b200.trainers import BallAttractorTrainer
    cfg, handler, agent, draws = _setup(16)
    torch.manual_seed(4)
    a = BallAttractorTrainer(copy.deepcopy(cfg))
    fa = a.fused_loop(5)
    losses_a = [float(fa.run(agent, handler, x0, v0)) for x0, v0 in draws]
    torch.manual_seed(4)
    b = BallAttractorTrainer(copy.deepcopy(cfg))
    run = b.graphed_iteration(agent, handler, fused=True)
    losses_b = [float(run(x0, v0)) for x0, v0 in draws]
    assert losses_a == losses_b
    assert torch.equal(fa.pflat, b.fused_loop().pflat)
    assert torch.equal(fa.exp_avg_sq, b.fused_loop().exp_avg_sq)
