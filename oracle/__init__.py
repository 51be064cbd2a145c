"""CPU oracle for the doper/sim differentiable ball-rollout hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import or execute it, and there
only as the checker (or as the timed CPU baseline), never as a fallback for the
CUDA path.  The product (``doper_b200``) raises if its CUDA library is missing.

Pinning status:
  * Euler / force arithmetic (reference ``ball_sim.py:28-79``): pinned by the
    only recorded numeric output of the reference, notebook cell 7
    (``tests/golden/notebook_cell7.json``).
  * Collision sweep, argmin tie-breaking, ``resolve_collision``, inner check and
    every gradient: PARITY UNPINNED by the reference (it has no tests, no golden
    vectors, and JAX is not installable here) — pinned only by this restatement.
"""
