"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).
usage: python tools/launch_list.py launches.csv "<command that was profiled>" > profiles/...md"""
import csv
import sys
from collections import OrderedDict

lines = [ln for ln in open(sys.argv[1]) if ln.startswith('"')]
rows = list(csv.DictReader(lines))
agg = OrderedDict()
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    if r.get("Metric Unit") in ("us", "usecond"):
        v *= 1e3
    a = agg.setdefault(r["Kernel Name"].split("(")[0], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
step = {k: a for k, a in agg.items() if any(t in k for t in ("rollout_", "l1_loss", "loss_jacobian", "allreduce_oneshot"))}
stot = sum(a[1] for a in step.values())
print(f"# ncu launch list — `{sys.argv[2]}`, gpu__time_duration.sum (ns), --clock-control none")
print("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
print("| kernel | launches | mean ns | total ns | share of all | share of the step kernels |")
print("|---|---:|---:|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    s2 = f"{100 * t / stot:.1f}%" if k in step else "—"
    print(f"| `{k}` | {n} | {t / n:.0f} | {t:.0f} | {100 * t / tot:.1f}% | {s2} |")
