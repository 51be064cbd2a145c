"""Extract the only recorded numeric output of the reference into a golden fixture.

Reads ``/root/reference/scripts/notebooks/jax_ball.ipynb`` (cell 7: the printed result of
``jax.vmap(run_sim)(...)``, 3 balls x 20 Euler steps, no collisions) together with the
inputs of cells 2 and 5, and writes ``tests/golden/notebook_cell7.json``.  Run once, in
the build container (the reference tree does not exist on the GPU box); the JSON is
committed.  The numbers are parsed from the notebook's stored text output — nothing is
recomputed here.
"""
import json
import re
import sys
from pathlib import Path

NB = Path("/root/reference/scripts/notebooks/jax_ball.ipynb")
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden" / "notebook_cell7.json"


def main():
    nb = json.loads(NB.read_text())
    cell = nb["cells"][7]
    assert "".join(cell["source"]).strip() == "test"
    text = "".join(cell["outputs"][0]["data"]["text/plain"])
    arrays = re.findall(r"DeviceArray\(\[(.*?)\], dtype=float32\)", text, flags=re.S)
    parsed = []
    for a in arrays:
        nums = [float(x) for x in re.findall(r"-?\d+\.\d*(?:e-?\d+)?", a)]
        assert len(nums) == 6, nums
        parsed.append([nums[0:2], nums[2:4], nums[4:6]])
    assert len(parsed) == 22, len(parsed)
    golden = {
        "source": "reference scripts/notebooks/jax_ball.ipynb cell 7 (text/plain output)",
        "physics": {  # notebook cell 2 / cell 5 (notebook-variant physics, see oracle/notebook_cell7.py)
            "potential_scale": 20, "g": 9.8, "sim_time": 0.2, "n_steps": 20,
            "radius": 0.05, "ro": 1000.0, "f": 0.007,
            "coordinate_init": [[0.2, 0.4]] * 3,
            "velocity_init": [[1.0, 0.0], [1.0, 0.1], [1.0, 0.0]],
            "attractor_arg": [[0.5, 0.5]] * 3,  # passed as `target_coordinate` to run_sim (cell 6)
        },
        "final_coordinate": parsed[0],
        "final_velocity": parsed[1],
        "trajectory": parsed[2:],  # 20 x 3 x 2, position at the START of each step
        "printed_digits": "numpy repr of float32 (shortest round-trip, <= 8 significant digits)",
    }
    OUT.parent.mkdir(parents=True, exist_ok=True)
    OUT.write_text(json.dumps(golden, indent=1))
    print("wrote", OUT, "arrays:", len(parsed))


if __name__ == "__main__":
    sys.exit(main())
