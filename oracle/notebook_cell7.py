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


def run_notebook_sim(physics: dict):
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
        a = acceleration(pot, fr, c)  # same helper as the rollout oracle
        x, v = euler(x, v, a, dt)  # same helper as the rollout oracle
    return x, v, np.stack(traj)
