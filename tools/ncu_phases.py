"""Buckets an `ncu --page source --print-source sass --csv` listing of raster_meshlets_kernel into the kernel's phases,
using `nvdisasm -g -c` line info of the same cubin.  Helper code inlined from lh_internal.cuh is charged to the phase
of the last kernel-body line seen in address order.
usage: ncu_phases.py <sass.csv> <nvdisasm output> <kernel substring>"""
import csv, re, sys, collections
csvp, sassp, kern = sys.argv[1:4]
rows = list(csv.reader(open(csvp)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ci, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
inst = [(r[1].strip(), int(r[ci] or 0), int(r[si] or 0), int(r[ti] or 0)) for r in rows[hi + 1:] if len(r) > ci]
lines = open(sassp).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l][0]
seq, cur = [], None
for l in lines[start + 1:]:
    if l.startswith(".text."):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((cur, m.group(2)))
print(f"# {len(inst)} profiled, {len(seq)} disassembled", file=sys.stderr)
PHASES = [(642, "0 setup/ticket/prefetch"), (700, "1 re-test vs stored keys"), (721, "2 header (pointer chase)"), (731, "3 vertex phase"),
          (765, "4 exact-bounds early-outs"), (796, "5 triangle cull phase"), (866, "6 raster phase"), (927, "7 epilogue")]
HELPERS = [(65, 83, "3 vertex phase"), (91, 105, "6 raster phase (emit)"), (113, 113, None), (260, 289, "2 header (pointer chase)"),
           (311, 313, None), (326, 330, "2 header (pointer chase)"), (332, 340, "5 triangle cull phase (push_deferred)")]
def phase_of(line):
    p = PHASES[0][1]
    for lo, name in PHASES:
        if line >= lo: p = name
    return p
per = collections.defaultdict(lambda: [0, 0, 0, 0])
ops = collections.defaultdict(lambda: collections.Counter())
last = PHASES[0][1]
for k in range(min(len(inst), len(seq))):
    key, op = seq[k]
    ph = last
    if key and key[0] == "raster.cu":
        if key[1] >= 642:
            ph = last = phase_of(key[1])
        else:
            for lo, hi_, name in HELPERS:
                if lo <= key[1] <= hi_ and name: ph = name
    p = per[ph]
    p[0] += inst[k][1]; p[1] += inst[k][2]; p[2] += inst[k][3]; p[3] += 1
    ops[ph][op.split()[0].split(".")[0] if not op.startswith("@") else op.split()[1].split(".")[0]] += inst[k][1]
ti_, ts_ = sum(p[0] for p in per.values()), sum(p[1] for p in per.values())
print(f"total warp-instr {ti_}  samples {ts_}")
for ph in sorted(per):
    p = per[ph]
    top = ", ".join(f"{o} {100*c/max(1,p[0]):.0f}%" for o, c in ops[ph].most_common(6))
    print(f"{ph:40s} sass {p[3]:4d} inst {p[0]:11d} ({100*p[0]/ti_:5.1f}%) samples {100*p[1]/max(1,ts_):5.1f}% thr/inst {p[2]/max(1,p[0]):5.1f} | {top}")
