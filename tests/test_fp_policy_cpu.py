"""The floating-point contraction policy (include/doper_fp_policy.h): what the reference's recorded output
pins, that the C oracle follows the shared header, and how little of the collision path depends on the
choices nothing pins.  CPU only."""
import itertools
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import cref
from oracle.notebook_cell7 import FP_FORCE, FP_POS, FP_VEL, run_notebook_sim

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = Path(__file__).parent / "golden"


def _mismatches(policy):
    g = json.loads((GOLDEN / "notebook_cell7.json").read_text())
    out = run_notebook_sim(g["physics"], policy=policy)
    gold = [np.asarray(g[k], np.float32) for k in ("final_coordinate", "final_velocity", "trajectory")]
    return sum(int((a.view(np.int32) != b.view(np.int32)).sum()) for a, b in zip(out, gold))


def test_notebook_pins_exactly_one_policy_of_the_euler_step():
    """Of the 8 ways to contract {x + v dt, v + a dt, -(k y) + f} only 'coordinate update alone' reproduces all
    126 recorded float32 values: POS is pinned ON, VEL and FORCE are pinned OFF."""
    bad = {}
    for pos, vel, force in itertools.product((0, 1), repeat=3):
        bad[(pos, vel, force)] = _mismatches(pos * FP_POS | vel * FP_VEL | force * FP_FORCE)
    assert bad[(1, 0, 0)] == 0
    assert all(v > 0 for k, v in bad.items() if k != (1, 0, 0)), bad
    assert bad[(0, 0, 0)] == 36  # 'as written': 36 values off by one ulp (tests/test_oracle.py)


def test_default_mask_is_the_pinned_one_everywhere():
    hdr = (ROOT / "include" / "doper_fp_policy.h").read_text()
    assert "#define DOPER_FP_POLICY DOPER_FP_POS" in hdr
    assert cref.FP_DEFAULT == cref.FP_POS == 1
    assert cref.CRef("portable").policy == cref.FP_POS
    # both sides include the one header
    assert '#include "../include/doper_fp_policy.h"' in (ROOT / "oracle" / "ball_sim_ref.c").read_text()
    assert '#include "../../include/doper_fp_policy.h"' in (ROOT / "doper_b200" / "csrc" / "physics.cuh").read_text()
    from doper_b200 import _abi
    assert _abi.lib.doper_fp_policy() == cref.FP_POS  # host-callable: the mask the kernels were compiled with


def test_other_masks_build_and_differ():
    m = cref.FP_POS | cref.FP_LERP | cref.FP_DOT
    cr = cref.CRef("portable", policy=m)
    assert cr.policy == m
    rng = np.random.default_rng(0)
    pts = rng.uniform(-5, 5, size=(20000, 2)).astype(np.float32)
    segs = rng.uniform(-5, 5, size=(60, 2, 2)).astype(np.float32)
    i0, d0 = cref.CRef("portable").closest_segment(pts, segs)
    i1, d1 = cr.closest_segment(pts, segs)
    assert (d0.view(np.uint32) != d1.view(np.uint32)).any()      # the arithmetic did change ...
    assert (np.abs(d0 - d1) <= 4e-6 * np.maximum(d0, 1.0)).all()  # ... by roundings only
    assert (i0 != i1).mean() < 1e-3


def test_collision_decisions_do_not_depend_on_the_unpinned_choices():
    """One-step sensitivity on a BASELINE configs[1] sub-batch (1,024 balls x 500 steps = 512,000 decisions taken
    from the SAME start states): no hit flag and no colliding segment changes under any contraction policy; only
    the closest segment of steps that do NOT collide (not an output of the reference) and last bits of the next
    state move.  The full batch is in profiles/r2_fp_policy_study_c2.json (tools/fp_policy_study.py)."""
    import sys
    sys.path.insert(0, str(ROOT / "tools"))
    import fp_policy_study as st
    res = st.study("c2", 1024)
    assert res["collisions_default_policy"] > 500
    for label, r in res["policies"].items():
        one = r["one_step"]
        assert one["hit_flag_flips"] == 0, (label, one)
        assert one["colliding_segment_changes"] == 0, (label, one)
        if r["mask"] != cref.FP_DEFAULT:
            assert one["start_states_whose_next_state_differs_in_a_bit"] > 0, label  # the masks are not no-ops
        # chaotic divergence after the first flipped bit stays a handful of balls
        assert r["rollout"]["balls_with_a_different_hit_sequence"] <= 8, (label, r["rollout"])


def test_committed_study_of_the_full_c2_batch():
    d = json.loads((ROOT / "profiles" / "r2_fp_policy_study_c2.json").read_text())
    assert d["balls"] == 4096 and d["n_steps"] == 500 and d["collisions_default_policy"] == 3032
    for label, r in d["policies"].items():
        assert r["one_step"]["hit_flag_flips"] == 0 and r["one_step"]["colliding_segment_changes"] == 0, label
