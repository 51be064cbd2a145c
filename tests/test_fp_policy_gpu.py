"""Switching the floating-point contraction policy is ONE number on both sides (include/doper_fp_policy.h):
the kernels compiled with -DDOPER_FP_POLICY=<mask> are bit-identical to the C oracle compiled with the same
mask, and differ from the default build.  Runs the alternative build in a subprocess (DOPER_B200_LIB)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu

CHILD = r"""
import json, sys
import numpy as np
sys.path.insert(0, %(root)r)
from doper_b200 import synthetic as syn, _abi
from doper_b200.sim import ball_sim
from oracle import ball_sim_np as onp
from oracle.cref import CRef
mask = _abi.lib.doper_fp_policy()
out = {"mask": int(mask)}
for n_balls, n, opts in ((700, 400, ball_sim.RolloutOptions(sequential_backward=True)),
                         (700, 400, ball_sim.RolloutOptions(sequential_backward=True, pipelined=False)),
                         (40000, 150, ball_sim.RolloutOptions(sequential_backward=True))):
    w = syn.make_workload("c2", n_balls=n_balls)
    consts = dict(w["constants"]); consts["attractor_strength"] = 300.0   # plenty of collisions
    c = onp.make_constants(consts)
    dt = np.float32((n / 1000) / n)
    res = {}
    for name, pol in (("same_policy", mask), ("default_policy", 1)):
        ref = CRef("portable", policy=pol).value_and_grad(n, dt, c, w["attractor"], w["target"], w["segments"], w["x0"], w["v0"])
        got = ball_sim.rollout_history(n / 1000, n, w["segments"], w["x0"], w["v0"], w["attractor"], consts, opts)
        (_, (xT, vT)), gv = ball_sim.vmapped_grad_and_value(n / 1000, n, w["segments"], w["x0"], w["v0"], w["target"], w["attractor"], consts, opts)
        bits = lambda a, b: int((np.asarray(a, np.float32).view(np.uint32) != np.asarray(b, np.float32).view(np.uint32)).sum())
        res[name] = {"collisions": int(ref["hit"].sum()), "hit_mismatch": int((got["hit"] != ref["hit"]).sum()),
                     "idx_mismatch": int((got["idx"][ref["hit"] & got["hit"]] != ref["idx"][ref["hit"] & got["hit"]]).sum()),
                     "xT_bits": bits(xT, ref["x_T"]), "vT_bits": bits(vT, ref["v_T"]), "grad_bits": bits(gv, ref["grad_v0"])}
    out[f"{n_balls}x{n}:{'pipe' if opts.pipelined is None else 'baked'}"] = res
print(json.dumps(out))
"""


def test_alternative_policy_build_matches_the_oracle_built_with_the_same_mask():
    from doper_b200 import build as b
    lib = b.lib_path(b.ALT_POLICY)
    if not b.is_current(b.ALT_POLICY):
        pytest.skip(f"{lib.name} is missing or older than the sources (python -m doper_b200.build --all, or __graft_entry__.build())")
    env = dict(os.environ, DOPER_B200_LIB=str(lib))
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": str(ROOT)}], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out.pop("mask") == b.ALT_POLICY
    for case, res in out.items():
        same, other = res["same_policy"], res["default_policy"]
        assert same["collisions"] > 500, case
        assert same["hit_mismatch"] == 0 and same["idx_mismatch"] == 0, (case, same)
        assert same["xT_bits"] == 0 and same["vT_bits"] == 0 and same["grad_bits"] == 0, (case, same)
        assert other["xT_bits"] > 0, (case, "the alternative mask must change the arithmetic", other)
