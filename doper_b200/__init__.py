"""doper_b200 — B200-native (sm_100a) differentiable ball-physics rollout.

A from-scratch drop-in for the one hot path of VGGVRobotics/doper: ``doper/sim`` (ball vs
line-segment collision sweep, inner check, time-of-impact reflection, semi-implicit Euler) and
its reverse-mode backward.  ``doper_b200.sim.ball_sim`` / ``doper_b200.sim.jax_geometry`` mirror
the reference modules of the same names; the compute lives in ``csrc/`` behind the C ABI declared
in ``include/doper_b200.h``.  There is no CPU fallback: importing ``doper_b200.sim`` without the
built library raises ``ImportError``; calling it without a CUDA device raises ``RuntimeError``.
"""
__version__ = "0.1.0"
