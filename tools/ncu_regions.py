"""Summarise an `ncu --page source --csv` export: instructions executed per warp-step and stall
samples, grouped into runs of consecutive SASS instructions with equal execution counts."""
import csv
import sys

path, warps_steps = sys.argv[1], float(sys.argv[2])
rows = list(csv.reader(open(path)))
hdr, data = rows[1], rows[2:]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
out = []
for r in data:
    try:
        ex = int(r[iex])
    except Exception:
        continue
    out.append((r[ia][-5:], r[isrc].strip()[:70], ex / warps_steps, int(r[ismp] or 0)))
tot = sum(o[2] for o in out)
S = sum(o[3] for o in out)
print(f"instructions per warp-step: {tot:.1f}; stall samples: {S}")
grp, cur = [], None
for a, s, e, smp in out:
    if cur and abs(cur["e"] - e) < 0.02 * max(cur["e"], 0.05) + 0.005:
        cur["n"] += 1; cur["smp"] += smp; cur["last"] = a; cur["tot"] += e
    else:
        cur = {"first": a, "last": a, "e": e, "n": 1, "smp": smp, "tot": e, "s": s}
        grp.append(cur)
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
for g in grp:
    if g["tot"] > thr or g["smp"] > 0.01 * S:
        print(f"{g['first']}-{g['last']} n={g['n']:4d} x{g['e']:6.2f} = {g['tot']:6.1f}/step  samples {100 * g['smp'] / S:5.1f}%  {g['s']}")
