"""Per-kernel totals of the last whole frame in an ncu launch list (gpu__time_duration.sum, --csv).
usage: python tools/launch_summary.py launches.csv [--all]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi = H.index("Kernel Name"), H.index("Metric Value")
si = H.index("Stream") if "Stream" in H else None
launches = [(r[ki].split("(")[0].split("::")[-1], float(r[vi].replace(",", "")) / 1000, r[si] if si is not None else "") for r in rows[hdr + 2:] if len(r) > vi]
starts = [i for i, (n, _, _) in enumerate(launches) if n.startswith("reset_kernel")]
frame = launches[starts[-2]:starts[-1]] if len(starts) >= 2 else launches
tot = collections.OrderedDict()
for n, t, s in frame:
    e = tot.setdefault(n, [0, 0.0, set()])
    e[0] += 1
    e[1] += t
    e[2].add(s)
whole = sum(t for _, t, _ in frame)
print(f"{len(frame)} launches in the frame, {whole:.1f} us serialised")
for n, (c, t, s) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"  {n:34s} x{c:3d} {t:9.1f} us  {t / whole:6.1%}  streams {sorted(s)}")
if "--all" in sys.argv:
    for n, t, s in frame:
        print(f"    {s:>4s} {n:34s} {t:8.1f}")
