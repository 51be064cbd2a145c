"""Instructions executed and stall samples per CUDA source line, from
`ncu -i X.ncu-rep --kernel-name regex:K --page source --csv --print-source cuda,sass > f.csv`.
usage: python tools/ncu_lines.py f.csv <units, e.g. ball-steps> [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
cur, agg = None, []
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 8 or r[0] in ("Line No", "Function Name") or not r[0]:
        continue
    try:
        ex, smp = int(r[7]), int(r[6] or 0)
    except ValueError:
        continue
    agg.append((ex, smp, cur, r[0], r[1].strip()[:110]))
tot, S = sum(a[0] for a in agg), sum(a[1] for a in agg)
print(f"instructions {tot} = {tot / units:.2f} per unit; stall samples {S}")
for ex, smp, f, l, src in sorted(agg, reverse=True)[:top]:
    print(f"{ex / units:6.2f} {100 * smp / max(S, 1):5.1f}%  {f}:{l}  {src}")
