"""How much of the collision path depends on the UNPINNED floating-point contraction choices?

The reference is a JAX program; whether XLA:CPU fused a given ``a*b + c`` is not visible in its source and
only three expressions are pinned by its recorded output (include/doper_fp_policy.h).  This CPU-only study
rebuilds the C oracle with each plausible contraction policy and counts, on a BASELINE workload,

  one_step   for every step of the DEFAULT policy's trajectory, the decision (hit flag, colliding segment,
             closest segment) the other policy takes from the SAME start state: isolates the decision
             sensitivity from the chaotic divergence that follows the first flipped bit;
  rollout    the policy's own rollout from the same inputs: balls whose hit sequence differs anywhere,
             final-state differences.

    python tools/fp_policy_study.py [workload=c2] [balls=0] > profiles/r2_fp_policy_study_c2.json
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from doper_b200 import synthetic as syn  # noqa: E402
from oracle import ball_sim_np as onp  # noqa: E402
from oracle import cref  # noqa: E402

POLICIES = {
    "default (POS)": cref.FP_POS,
    "nothing contracted": 0,
    "POS+LERP": cref.FP_POS | cref.FP_LERP,
    "POS+DOT": cref.FP_POS | cref.FP_DOT,
    "POS+DOT(rev)": cref.FP_POS | cref.FP_DOT | cref.FP_DOT_REV,
    "POS+LERP+DOT": cref.FP_POS | cref.FP_LERP | cref.FP_DOT,
    "POS+LERP+DOT(rev)": cref.FP_POS | cref.FP_LERP | cref.FP_DOT | cref.FP_DOT_REV,
    "POS+REFLECT/FORCE": cref.FP_POS | cref.FP_FORCE,
    "POS+VEL (refuted by the notebook)": cref.FP_POS | cref.FP_VEL,
    "everything contracted": cref.FP_POS | cref.FP_VEL | cref.FP_LERP | cref.FP_DOT | cref.FP_FORCE,
}


def study(name="c2", balls=0, policies=None):
    w = syn.make_workload(name, n_balls=balls or None) if balls else syn.make_workload(name)
    n, B = w["n_steps"], w["x0"].shape[0]
    c = onp.make_constants(w["constants"])
    dt = np.float32(w["sim_time"] / n)
    args = (c, w["attractor"], w["segments"])
    base = cref.CRef("portable").rollout_fwd(n, dt, *args, w["x0"], w["v0"], want_traj=True)
    # start state of every step of the default trajectory, flattened to n*B one-step "balls"
    xs = np.concatenate([w["x0"][None], base["traj_x"][:-1]], axis=0).reshape(-1, 2)
    vs = np.concatenate([w["v0"][None], base["traj_v"][:-1]], axis=0).reshape(-1, 2)
    hit0, idx0 = base["hit"].reshape(-1), base["idx"].reshape(-1)
    out = {"workload": name, "balls": int(B), "n_steps": int(n), "segments": int(w["segments"].shape[-3]),
           "ball_steps": int(B * n), "collisions_default_policy": int(hit0.sum()), "policies": {}}
    for label, mask in (policies or POLICIES).items():
        cr = cref.CRef("portable", policy=mask)
        assert cr.policy == mask
        one = cr.rollout_fwd(1, dt, *args, xs, vs)
        h1, i1 = one["hit"].reshape(-1), one["idx"].reshape(-1)
        both = hit0 & h1
        full = cr.rollout_fwd(n, dt, *args, w["x0"], w["v0"])
        differs = (full["hit"] != base["hit"]).any(axis=0) | ((full["idx"] != base["idx"]) & base["hit"]).any(axis=0)
        moved = (full["x_T"].view(np.uint32) != base["x_T"].view(np.uint32)).any(axis=1)
        dx = np.abs(full["x_T"].astype(np.float64) - base["x_T"].astype(np.float64)).max(axis=1)
        out["policies"][label] = {
            "mask": int(mask),
            "one_step": {
                "hit_flag_flips": int((hit0 != h1).sum()),
                "colliding_segment_changes": int((i1[both] != idx0[both]).sum()),
                "closest_segment_changes_all_steps": int((i1 != idx0).sum()),
                "start_states_whose_next_state_differs_in_a_bit": int(
                    ((one["x_T"].view(np.uint32) != base["traj_x"].reshape(-1, 2).view(np.uint32)).any(axis=1) |
                     (one["v_T"].view(np.uint32) != base["traj_v"].reshape(-1, 2).view(np.uint32)).any(axis=1)).sum()),
            },
            "rollout": {
                "balls_with_a_different_hit_sequence": int(differs.sum()),
                "balls_with_a_different_final_position_bit": int(moved.sum()),
                "final_position_abs_diff_m": {"median": float(np.median(dx)), "q99": float(np.quantile(dx, 0.99)),
                                              "max": float(dx.max())},
                "collisions": int(full["hit"].sum()),
            },
        }
    return out


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    balls = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    print(json.dumps(study(name, balls), indent=1))
