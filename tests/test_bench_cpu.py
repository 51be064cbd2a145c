"""bench.py's contract that can be checked without a GPU: the reference arm (the oracle's C port on the
host cores) prints ONE JSON line with the keys the driver reads, and the product arm refuses to run
without CUDA instead of falling back to anything."""
import json
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]


def _run(*args):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    out = _run("--impl", "reference", "--workload", "c1", "--steps", "2", "--warmup", "1")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ball_steps_per_sec_fwd_bwd" and d["unit"] == "ball-steps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32" and d["n_gpus"] == 1
    assert d["steps"] == 2 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("c1:")
    cb, e2e = d["cpu_baseline"], d["e2e"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert e2e == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        return  # on a GPU box the product arm is exercised by the gpu tests and by the bench itself
    out = _run("--steps", "1", "--warmup", "1", "--no-cpu")
    assert out.returncode != 0
    assert not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]  # no number from a fallback


def test_clock_sampler_counts_only_samples_inside_the_timed_region():
    """bench.py's `clocks`: the median SM clock and the throttle reasons come from the samples taken inside the timed
    region (idle clocks before it do not count), and a region without samples reports None, never an idle clock."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    s = bench.ClockSampler(0)
    s.max_mhz = 1965
    s.samples = [(0.5, 600, 0x1), (1.1, 1965, 0x0), (1.2, 1950, 0x4), (1.3, 1965, 0x0), (2.5, 700, 0x8)]
    s.window = (1.0, 2.0)
    out = s.summary()
    assert out["samples"] == 3 and out["sm_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"] and out["sm_max_mhz"] == 1965
    s.window = (3.0, 4.0)
    assert s.summary()["sm_mhz"] is None and s.summary()["samples"] == 0
