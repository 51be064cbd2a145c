"""torch-CPU autograd restatement of the reference rollout (gradient truth).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED by the
reference (no gradient golden exists; notebook cells 9/13 are stale-state).

Same op order as ``oracle/ball_sim_np.py`` (so the fp32 forward is bit-identical
to it — asserted in ``tests/test_oracle.py``), written with ``torch.where`` so
that the batched ``lax.cond`` of the reference (``ball_sim.py:163-168`` under
``vmap`` ``:332``) keeps its *select* semantics in reverse mode: both branches
run for every ball and the unselected one receives a zero cotangent (which can
still leak ``0*inf = NaN``).  ``dtype=torch.float64`` gives the
high-precision variant used to calibrate tolerances and for finite differences.

The gradient is taken w.r.t. ``velocity_init`` only (``value_and_grad(..., 4)``,
``ball_sim.py:334-338``); ``coordinate_init`` can optionally be differentiated
too (needed to check the x0 cotangent the CUDA backward also produces).
"""
from __future__ import annotations

import numpy as np
import torch

from .ball_sim_np import Consts, make_constants, G_FREE_FALL  # noqa: F401


class _ExactSqrt(torch.autograd.Function):
    """Correctly rounded sqrt.  ``torch.sqrt`` on CPU is NOT correctly rounded in this
    build (0.65 % of fp32 inputs differ from IEEE sqrt by 1 ulp — measured), whereas
    XLA:CPU, numpy and CUDA ``__fsqrt_rn`` are; so the value comes from numpy and only
    the derivative ``g / (2*sqrt(x))`` is torch arithmetic."""

    @staticmethod
    def forward(ctx, x):
        with np.errstate(invalid="ignore"):
            y = torch.from_numpy(np.sqrt(x.detach().numpy()))
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        return g / (y + y)


def _sqrt(x):
    return _ExactSqrt.apply(x)


def _pos_update(v, dt, x):
    """``fma(v, dt, x)``: get_new_state's coordinate update is single-rounding (see
    ``ball_sim_np.CONTRACT_POSITION_UPDATE``).  fp32 inputs: product exact in fp64, one fp64
    add, one rounding back (double rounding has probability ~2^-29 per op; the bit-exactness
    of this path against numpy is asserted on the test inputs)."""
    if v.dtype == torch.float32:
        return (v.double() * dt.double() + x.double()).float()
    return x + v * dt


def _proj(px, py, s0x, s0y, s1x, s1y):
    svx = s1x - s0x
    svy = s1y - s0y
    L = svx * svx + svy * svy
    pvx = px - s0x
    pvy = py - s0y
    t = (pvx * svx + pvy * svy) / L
    t = torch.minimum(torch.maximum(t, torch.zeros_like(t)), torch.ones_like(t))
    return s0x + t * svx, s0y + t * svy


def _friction(vx, vy, m, g, f, r):
    one = torch.ones_like(vx)
    cx = torch.minimum(torch.maximum(vx, -one), one)
    cy = torch.minimum(torch.maximum(vy, -one), one)
    return ((((-cx) * m) * g) * f) / r, ((((-cy) * m) * g) * f) / r


def rollout(
    sim_time,
    n_steps,
    segments,
    coordinate_init,
    velocity_init,
    attractor,
    constants,
    dtype=torch.float32,
    forced_idx=None,
    forced_hit=None,
    grad_wrt_x0=False,
    num_threads=None,
):
    """Differentiable rollout. Returns dict(x_T, v_T, x0, v0, idx [n,B], hit [n,B]).

    ``forced_idx/forced_hit`` replay a recorded discrete history (used by the
    fp64 variant so it differentiates the same piecewise-smooth branch as fp32).
    """
    if num_threads:
        torch.set_num_threads(num_threads)
    c = constants if isinstance(constants, Consts) else make_constants(constants)
    if dtype == torch.float32:
        T = lambda s: torch.tensor(np.float32(s), dtype=dtype)  # noqa: E731
        dt = T(np.float32(sim_time / n_steps))
    else:  # high precision variant keeps the f32-rounded constants (same physical system)
        T = lambda s: torch.tensor(float(np.float32(s)), dtype=dtype)  # noqa: E731
        dt = T(np.float32(sim_time / n_steps))
    r, m, f, e, ks, g = T(c.radius), T(c.mass), T(c.friction), T(c.elasticity), T(c.k_s), T(c.g)
    two, zero = T(2.0), T(0.0)

    seg = torch.as_tensor(np.ascontiguousarray(segments, dtype=np.float32)).to(dtype)
    per_env = seg.dim() == 4
    in_dt = np.float64 if dtype == torch.float64 else np.float32  # fp64 variant keeps fp64 inputs
    x0 = torch.as_tensor(np.ascontiguousarray(coordinate_init, dtype=in_dt)).to(dtype).reshape(-1, 2)
    v0 = torch.as_tensor(np.ascontiguousarray(velocity_init, dtype=in_dt)).to(dtype).reshape(-1, 2)
    x0 = x0.clone().requires_grad_(grad_wrt_x0)
    v0 = v0.clone().requires_grad_(True)
    A = torch.as_tensor(np.ascontiguousarray(attractor, dtype=np.float32)).to(dtype).reshape(2)
    Ax, Ay = A[0], A[1]
    B = x0.shape[0]
    rows = torch.arange(B)

    S0x, S0y, S1x, S1y = seg[..., 0, 0], seg[..., 0, 1], seg[..., 1, 0], seg[..., 1, 1]

    x_, y_, vx, vy = x0[:, 0], x0[:, 1], v0[:, 0], v0[:, 1]
    idx_log, hit_log = [], []
    for step in range(n_steps):
        # forces + Euler (ball_sim.py:192-202)
        px_, py_ = -(two * (x_ - Ax)), -(two * (y_ - Ay))
        frx, fry = _friction(vx, vy, m, g, f, r)
        ax, ay = (ks * px_ + frx) / m, (ks * py_ + fry) / m
        v1x, v1y = vx + ax * dt, vy + ay * dt
        x1, y1 = _pos_update(v1x, dt, x_), _pos_update(v1y, dt, y_)
        # sweep (no gradient flows through argmin / compare; jax_geometry.py:172-174)
        with torch.no_grad():
            if forced_idx is None:
                prx, pry = _proj(x1[:, None], y1[:, None], S0x, S0y, S1x, S1y)
                dx, dy = x1[:, None] - prx, y1[:, None] - pry
                d = _sqrt(dx * dx + dy * dy)
                # first minimal index, NaN counts as minimal (numpy argmin semantics)
                idx = torch.from_numpy(np.argmin(d.numpy(), axis=1))
                dist = d[rows, idx]
                hit = dist <= r
            else:
                idx = torch.as_tensor(forced_idx[step]).long()
                hit = torch.as_tensor(forced_hit[step]).bool()
        idx_log.append(idx.numpy().astype(np.int32))
        hit_log.append(hit.numpy().copy())
        if per_env:
            s0x, s0y, s1x, s1y = S0x[rows, idx], S0y[rows, idx], S1x[rows, idx], S1y[rows, idx]
        else:
            s0x, s0y, s1x, s1y = S0x[idx], S0y[idx], S1x[idx], S1y[idx]
        if bool(hit.any()):
            # resolve_collision (ball_sim.py:97-132), evaluated for every ball (select)
            c1x, c1y = _proj(x1, y1, s0x, s0y, s1x, s1y)
            ox, oy = c1x - x_, c1y - y_
            od = _sqrt(ox * ox + oy * oy)
            nx, ny = ox / od, oy / od
            toi = torch.abs(od - r) / torch.abs(nx * v1x + ny * v1y)
            toi = torch.minimum(torch.maximum(toi, zero), dt)
            vsx, vsy = vx + ax * toi, vy + ay * toi
            xsx, xsy = _pos_update(vsx, toi, x_), _pos_update(vsy, toi, y_)
            c2x, c2y = _proj(xsx, xsy, s0x, s0y, s1x, s1y)
            o2x, o2y = c2x - xsx, c2y - xsy
            o2d = _sqrt(o2x * o2x + o2y * o2y)
            n2x, n2y = o2x / o2d, o2y / o2d
            k = n2x * vsx + n2y * vsy
            vnx, vny = n2x * k, n2y * k
            vpx, vpy = vsx - vnx, vsy - vny
            vax, vay = vpx - e * vnx, vpy - e * vny
            pbx, pby = -(two * (xsx - Ax)), -(two * (xsy - Ay))
            fbx, fby = _friction(vax, vay, m, g, f, r)
            abx, aby = (pbx + fbx) / m, (pby + fby) / m
            tau = dt - toi
            vbx, vby = vax + abx * tau, vay + aby * tau
            xbx, xby = _pos_update(vbx, tau, xsx), _pos_update(vby, tau, xsy)
            x_, y_ = torch.where(hit, xbx, x1), torch.where(hit, xby, y1)
            vx, vy = torch.where(hit, vbx, v1x), torch.where(hit, vby, v1y)
        else:
            x_, y_, vx, vy = x1, y1, v1x, v1y
    return {
        "x_T": torch.stack([x_, y_], dim=1),
        "v_T": torch.stack([vx, vy], dim=1),
        "x0": x0,
        "v0": v0,
        "idx": np.stack(idx_log) if idx_log else np.zeros((0, B), np.int32),
        "hit": np.stack(hit_log) if hit_log else np.zeros((0, B), bool),
    }


def value_and_grad(
    sim_time, n_steps, segments, coordinate_init, velocity_init, target, attractor, constants,
    dtype=torch.float32, forced_idx=None, forced_hit=None, grad_wrt_x0=False, num_threads=None,
):
    """Restates ``vmapped_grad_and_value`` (ball_sim.py:334-338).

    Returns ((loss, (x_T, v_T)), dL/dv0[, dL/dx0]) as numpy arrays.
    """
    out = rollout(sim_time, n_steps, segments, coordinate_init, velocity_init, attractor,
                  constants, dtype, forced_idx, forced_hit, grad_wrt_x0, num_threads)
    tgt = torch.as_tensor(np.ascontiguousarray(target, dtype=np.float32)).to(dtype).reshape(1, 2)
    per_ball = torch.abs(out["x_T"] - tgt).sum(dim=1)  # ball_sim.py:286
    loss = per_ball.sum()  # :324
    if out["x_T"].requires_grad:
        loss.backward()
        gv = out["v0"].grad.numpy().copy()
        gx = out["x0"].grad.numpy().copy() if grad_wrt_x0 else None
    else:  # n_steps == 0
        gv = np.zeros_like(out["v0"].detach().numpy())
        gx = None
    res = ((loss.item(), (out["x_T"].detach().numpy(), out["v_T"].detach().numpy())), gv)
    aux = {"idx": out["idx"], "hit": out["hit"], "per_ball_loss": per_ball.detach().numpy(), "grad_x0": gx}
    return res, aux
