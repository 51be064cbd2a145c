"""Per-source-line hot spots of one kernel of an .ncu-rep captured with --import-source on (-lineinfo build):
    python tools/ncu_source_hot.py report.ncu-rep <kernel regex> [launch index] [top N]
Prints, per file:line, the warp-stall samples and warp instructions executed (sorted by samples)."""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{rx}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and r[0].strip().isdigit():
        d = dict(zip(hdr, r))
        try:
            out.append((int(d["# Samples"]), int(d["Instructions Executed"]), cur, int(r[0]), r[1].strip()[:110]))
        except (ValueError, KeyError):
            pass
ts, ti = sum(o[0] for o in out), sum(o[1] for o in out)
print(f"# {rep} kernel ~ {rx}: {ts} stall samples, {ti} warp instructions attributed to source lines")
byfile = {}
for s, i, f, _, _ in out:
    a = byfile.setdefault(f, [0, 0]); a[0] += s; a[1] += i
for f, (s, i) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"#   {f:18s} samples {100 * s / max(ts, 1):5.1f}%  instructions {100 * i / max(ti, 1):5.1f}%")
print("samples%  instr%  file:line  source")
for s, i, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{100 * s / max(ts, 1):6.2f}  {100 * i / max(ti, 1):6.2f}  {f}:{ln}  {src}")
