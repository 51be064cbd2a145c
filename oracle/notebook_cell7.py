"""Known-answer check of the oracle's Euler/force arithmetic against notebook cell 7.

TEST INFRASTRUCTURE ONLY.  The reference notebook (``scripts/notebooks/jax_ball.ipynb``
cells 2, 5, 6, 7) is an earlier, collision-free prototype whose physics differ from
``ball_sim.py`` in two constants-only ways: the potential is scaled by 20 and friction is
``-sign(v) * m * g * radius * f / radius``.  Its force -> acceleration -> semi-implicit
Euler arithmetic is the same code shape as ``ball_sim.py:61,75-76``, so reproducing its
printed float32 output to every digit pins: f32 weak-typed constants, left-to-right
products, ``(F_pot + F_fr)/m``, ``v' = v + a*dt``, ``x' = x + v'*dt``, and no FMA
contraction in those expressions.  It pins nothing about collisions.
"""
import numpy as np

from .ball_sim_np import acceleration, euler, Consts

f32 = np.float32


FP_POS, FP_VEL, FP_FORCE = 1, 2, 32  # bits of include/doper_fp_policy.h that this prototype exercises


def run_notebook_sim(physics: dict, policy: int | None = None):
    """``policy``: None = the oracle's default helpers (ball_sim_np.euler); else a contraction mask, evaluated
    here explicitly so that tests can show which masks reproduce the recorded output."""
    r = physics["radius"]
    volume = 4 * np.pi * (r ** 3) / 3  # cell 5: python-double arithmetic
    mass = volume * physics["ro"]
    c = Consts(radius=f32(r), mass=f32(mass), friction=f32(physics["f"]), elasticity=f32(0),
               k_s=f32(0), g=f32(physics["g"]))
    n = physics["n_steps"]
    dt = f32(physics["sim_time"] / n)
    x = np.array(physics["coordinate_init"], np.float32)
    v = np.array(physics["velocity_init"], np.float32)
    A = np.array(physics["attractor_arg"], np.float32)
    scale = f32(physics["potential_scale"])
    traj = []
    for _ in range(n):
        traj.append(x.copy())
        pot = -(scale * (f32(2.0) * (x - A)))  # -grad(20*sum((x-A)**2))
        fr = ((((-np.sign(v)) * c.mass) * c.g) * c.radius) * c.friction / c.radius  # cell 2
        if policy is None:
            a = acceleration(pot, fr, c)  # same helper as the rollout oracle
            x, v = euler(x, v, a, dt)  # same helper as the rollout oracle
            continue
        from .ball_sim_np import fma_f32
        y = f32(2.0) * (x - A)
        num = fma_f32(np.full_like(y, -scale), y, fr) if policy & FP_FORCE else pot + fr  # -(scale*y) + fr
        a = num / c.mass
        v1 = fma_f32(a, dt, v) if policy & FP_VEL else v + a * dt
        x = fma_f32(v1, dt, x) if policy & FP_POS else x + v1 * dt
        v = v1
    return x, v, np.stack(traj)
