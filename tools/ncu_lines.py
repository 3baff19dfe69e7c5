"""Joins an `ncu --page source --csv` SASS listing with nvdisasm -g line info: per source line, instructions executed
and stall samples.  usage: ncu_lines.py <sass.csv> <nvdisasm -g -c output> <kernel substring> [top]"""
import csv, re, sys, collections
csvp, sassp, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(csvp)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ci, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
inst = [(r[1].strip(), int(r[ci] or 0), int(r[si] or 0), int(r[ti] or 0)) for r in rows[hi + 1:] if len(r) > ci]
# parse nvdisasm: find function, collect (line, opcode) per instruction
lines = open(sassp).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l]
# pick the section header of the form ".section .text.<mangled>" is absent in -c mode; use '.text.' label
assert start, "kernel not found"
cur = None
seq = []
i = start[0]
for l in lines[i + 1:]:
    if l.startswith(".text.") or l.startswith("//--------------------- .text") and seq:
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        inl = "inlined" in m.group(3)
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((cur, m.group(2)))
print(f"# {len(inst)} profiled instrs, {len(seq)} disassembled instrs", file=sys.stderr)
n = min(len(inst), len(seq))
per = collections.defaultdict(lambda: [0, 0, 0, 0])
for k in range(n):
    key = seq[k][0]
    p = per[key]
    p[0] += inst[k][1]; p[1] += inst[k][2]; p[2] += inst[k][3]; p[3] += 1
tot_i = sum(p[0] for p in per.values()); tot_s = sum(p[1] for p in per.values())
print(f"total warp-instr {tot_i}  samples {tot_s}")
for key, p in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{str(key):36s} sass {p[3]:4d}  inst {p[0]:12d} ({100*p[0]/tot_i:5.1f}%)  samples {p[1]:7d} ({100*p[1]/max(1,tot_s):5.1f}%)  thr/inst {p[2]/max(1,p[0]):5.1f}")
