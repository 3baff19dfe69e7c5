"""Per-source-line attribution of an `ncu --page source --print-source sass --csv` listing, using the `nvdisasm -g -c`
line info of the same cubin ("//## File ..., line N" markers, incl. inlined-at chains: the OUTERMOST raster.cu/cull.cu
line in the kernel body is charged, so helper code counts towards the kernel line that called it).
usage: ncu_by_line.py <sass.csv> <nvdisasm output> <kernel substring> [source file to print beside]"""
import csv, re, sys, collections
csvp, sassp, kern = sys.argv[1:4]
srcp = sys.argv[4] if len(sys.argv) > 4 else None
rows = list(csv.reader(open(csvp)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ci, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
inst = []
for r in rows[hi + 1:]:
    if len(r) <= ci: continue
    inst.append((r[1].strip(), int(r[ci] or 0), int(r[si] or 0), int(r[ti] or 0), {h: int(r[i] or 0) for i, h in stall_cols}))
lines = open(sassp).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l][0]
seq, cur = [], None
fname = (srcp or "raster.cu").split("/")[-1]
for l in lines[start + 1:]:
    if l.startswith(".text."): break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        # "inlined at" chains: take the last (outermost) location in the kernel's own file
        locs = [(m.group(1).split("/")[-1], int(m.group(2)))] + [(a.split("/")[-1], int(b)) for a, b in re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))]
        own = [x for x in locs if x[0] == fname]
        cur = own[-1] if own else locs[-1]
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: seq.append((cur, m.group(2)))
print(f"# {len(inst)} profiled, {len(seq)} disassembled", file=sys.stderr)
per = collections.defaultdict(lambda: [0, 0, 0, collections.Counter(), collections.Counter()])
for k in range(min(len(inst), len(seq))):
    key, op = seq[k]
    p = per[key]
    p[0] += inst[k][1]; p[1] += inst[k][2]; p[2] += inst[k][3]
    for h, v in inst[k][4].items(): p[3][h] += v
    p[4][(op.split()[1] if op.startswith("@") else op.split()[0]).split(".")[0]] += inst[k][1]
ti_, ts_ = sum(p[0] for p in per.values()), sum(p[1] for p in per.values())
src = open(srcp).read().split("\n") if srcp else None
print(f"total warp-instr {ti_}  samples {ts_}")
for key in sorted(per, key=lambda k: (k is None, k)):
    p = per[key]
    if p[0] * 400 < ti_ and p[1] * 400 < ts_: continue
    stalls = ", ".join(f"{h[6:]} {100*v/max(1,p[1]):.0f}%" for h, v in p[3].most_common(3))
    ops = ", ".join(f"{o} {100*c/max(1,p[0]):.0f}%" for o, c in p[4].most_common(4))
    text = src[key[1] - 1].strip()[:70] if src and key and key[0] == fname and key[1] <= len(src) else ""
    print(f"{str(key):28s} inst {100*p[0]/ti_:5.1f}% smp {100*p[1]/max(1,ts_):5.1f}% thr {p[2]/max(1,p[0]):4.1f} | {stalls:42s} | {ops:40s} | {text}")
