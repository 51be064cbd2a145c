"""Fill profiles/roofline_traffic.json (what bench.py's `roofline.traffic` / `roofline.executed` read) from an
`ncu --set full` report of the timed kernels on a named workload:
    python tools/roofline_from_ncu.py report.ncu-rep c2 4096 500 profiles/r2_ncu_fwd_pipe16_c2.txt"""
import csv
import json
import subprocess
import sys
from pathlib import Path

rep, workload, B, n_steps, summary = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
H = rows[0]
path = Path(__file__).resolve().parents[1] / "profiles" / "roofline_traffic.json"
doc = json.loads(path.read_text())


def num(v):
    return float(v.replace(",", "")) if v not in ("", "n/a") else 0.0


def unit_bytes(row, name):
    i = H.index(name)
    v, u = num(row[i]), rows[1][i]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


seen = set()
for row in rows[2:]:
    name = row[H.index("Kernel Name")].split("(")[0].replace("void ", "").strip()
    short = name.split("::")[-1]
    short = short.replace(" ", "")
    import re
    short = re.sub(r"pipe_kernel<(\d+),0(?:,\d+)?>", r"pipe_kernel<\1>", short)  # the name doper_rollout_plan_name reports
    if short in seen:
        continue
    seen.add(short)
    g = lambda k: num(row[H.index(k)])
    wi = g("smsp__inst_executed.sum")
    lanes = g("smsp__thread_inst_executed_per_inst_executed.ratio")
    # warp instructions on the FMA pipe (FADD / FMUL / FFMA issue there) x active lanes ~ FP32 lane-instructions
    fp = (g("sm__inst_executed_pipe_fma.sum") if "sm__inst_executed_pipe_fma.sum" in H else 0.0) * lanes
    ffma = 0.0
    doc.setdefault("kernels", {})[short] = {
        "workload": workload, "report": summary,
        "fma_pipe_pct": round(g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), 2),
        "issue_active_pct": round(g("smsp__issue_active.avg.pct_of_peak_sustained_active"), 2),
        "warp_instructions_per_launch": wi, "active_lanes_per_instruction": round(lanes, 2),
        "fp32_lane_ops_per_ball_step": None,  # this ncu version's `--set full` has no per-opcode thread counters in the raw page
        "warp_instructions_per_ball_step": round(wi / (B * n_steps), 2),
        "duration_us_under_ncu": round(g("gpu__time_duration.sum"), 2),
        "dram_bytes_per_launch": unit_bytes(row, "dram__bytes_read.sum") + unit_bytes(row, "dram__bytes_write.sum"),
    }
    if "rollout_fwd" in short:
        doc.setdefault("traffic", {})[workload] = doc["kernels"][short]["dram_bytes_per_launch"]
doc["_comment_r2"] = ("kernels.*: read from `ncu --set full` of the kernels bench.py times, per launch; bench.py looks the forward "
                      "kernel up by the name doper_rollout_plan_name reports and uses the entry only when `workload` matches.")
path.write_text(json.dumps(doc, indent=1))
print(json.dumps({k: v for k, v in doc["kernels"].items() if v.get("workload") == workload}, indent=1))
