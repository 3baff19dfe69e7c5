"""Prints the last frame of an ncu launch list, aggregated by kernel name (count, total us)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; data = rows[hi + 1:]
ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
last = [i for i, r in enumerate(data) if 'reset_kernel' in r[ki]][-1]
agg = collections.OrderedDict()
verbose = len(sys.argv) > 2
for r in data[last:]:
    n = r[ki].replace('unnamed>::', '').replace('void ', '').split('(')[0]
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += float(r[vi]) / 1e3
    if verbose: print(n, round(float(r[vi]) / 1e3, 1))
tot = sum(a[1] for a in agg.values())
for n, a in agg.items(): print(f"{n:40s} x{a[0]:3d} {a[1]:9.1f} us {100*a[1]/tot:5.1f}%")
print("total", round(tot, 1), "us")
