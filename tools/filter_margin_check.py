"""CPU emulation of the sweep's candidate filter: measures |sqrt(a_i) - d_ref_i| in units of
u*M (u = 2^-24, M = max |coordinate|) over adversarial inputs, to validate the margin
delta = 128*u*M used by sweep.cuh.  Pure numpy (float32 + exact fma)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle.ball_sim_np import fma_f32, distances_to_segments  # noqa: E402

f32 = np.float32


def approx_d2(p, segs):
    px, py = p[:, 0:1], p[:, 1:2]
    s0x, s0y = segs[:, 0, 0][None], segs[:, 0, 1][None]
    svx = segs[:, 1, 0][None] - s0x
    svy = segs[:, 1, 1][None] - s0y
    L = svx * svx + svy * svy
    rl = f32(1.0) / L
    pvx, pvy = px - s0x, py - s0y
    dot = fma_f32(pvx, svx, pvy * svy)
    t = np.clip(dot * rl, f32(0), f32(1))
    ex = fma_f32(-t, svx, pvx)
    ey = fma_f32(-t, svy, pvy)
    return fma_f32(ex, ex, ey * ey)


def run(name, p, segs):
    M = max(np.abs(p).max(), np.abs(segs).max())
    d = distances_to_segments(p, segs).astype(np.float64)
    a = np.sqrt(approx_d2(p, segs).astype(np.float64))
    err = np.abs(a - d) / (2.0 ** -24 * M)
    print(f"{name:40s} M={M:9.3g} max err = {err.max():8.2f} u*M   (99.99% {np.quantile(err, 0.9999):6.2f})")
    return err.max()


def main():
    rng = np.random.default_rng(0)
    worst = 0
    for trial in range(3):
        # generic: triangle soup scale
        segs = rng.uniform(-12, 20, (400, 2, 2)).astype(np.float32)
        p = rng.uniform(-14, 22, (4000, 2)).astype(np.float32)
        worst = max(worst, run("generic", p, segs))
        # short segments, far points
        c = rng.uniform(-85, 85, (400, 1, 2))
        segs = (c + rng.normal(0, 1e-4, (400, 2, 2))).astype(np.float32)
        p = rng.uniform(-85, 85, (4000, 2)).astype(np.float32)
        worst = max(worst, run("tiny segments", p, segs))
        # points almost on the segments (cancellation)
        segs = rng.uniform(-85, 85, (400, 2, 2)).astype(np.float32)
        t = rng.uniform(-0.2, 1.2, (4000, 1))
        k = rng.integers(0, 400, 4000)
        on = segs[k, 0] + t * (segs[k, 1] - segs[k, 0])
        p = (on + rng.normal(0, 1e-5, (4000, 2))).astype(np.float32)
        worst = max(worst, run("points on segments", p, segs))
        # long segments, off-centre domain
        segs = (rng.uniform(-1, 1, (400, 2, 2)) * 300 + 500).astype(np.float32)
        p = (rng.uniform(-1, 1, (4000, 2)) * 300 + 500).astype(np.float32)
        worst = max(worst, run("offset domain, long segments", p, segs))
        # nearly perpendicular foot close to the ends (clip boundaries)
        segs = rng.uniform(-12, 20, (400, 2, 2)).astype(np.float32)
        k = rng.integers(0, 400, 4000)
        sv = segs[k, 1] - segs[k, 0]
        nrm = np.stack([-sv[:, 1], sv[:, 0]], 1)
        end = np.where(rng.random((4000, 1)) < 0.5, segs[k, 0], segs[k, 1])
        p = (end + nrm * rng.uniform(-2, 2, (4000, 1))).astype(np.float32)
        worst = max(worst, run("feet at segment ends", p, segs))
    print("worst:", worst, "-> margin 128 u*M has slack x", 128 / worst)
    assert worst < 64


if __name__ == "__main__":
    main()
