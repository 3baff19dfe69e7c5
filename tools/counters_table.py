"""Tabulates an `ncu --metrics ... --csv` listing (long format) per launch.  usage: counters_table.py <csv> [last_n_launches]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
per = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < 15: continue
    k = int(r[0])
    d = per.setdefault(k, {'name': re.sub(r'void |unnamed>::|\(.*', '', r[4]), 'grid': r[8], 'stream': r[6]})
    try: d[r[12]] = float(r[14].replace(',', ''))
    except ValueError: d[r[12]] = r[14]
ids = list(per)
n = int(sys.argv[2]) if len(sys.argv) > 2 else len(ids)
ids = ids[-n:]
print(f"{'kernel':34s} {'strm':>4s} {'us':>8s} {'Minst':>8s} {'thr/in':>6s} {'iss%':>5s} {'wrp%':>5s} {'red M':>8s} {'atom M':>7s} {'ltsM':>8s} {'dramMB':>7s}")
tot = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for k in ids:
    d = per[k]
    us = d.get('gpu__time_duration.sum', 0) / 1e3
    inst = d.get('smsp__inst_executed.sum', 0)
    thr = d.get('smsp__thread_inst_executed.sum', 0)
    red = d.get('lts__t_sectors_op_red.sum', 0); atom = d.get('lts__t_sectors_op_atom.sum', 0)
    dram = (d.get('dram__bytes_read.sum', 0) + d.get('dram__bytes_write.sum', 0)) / 1e6
    print(f"{d['name'][:34]:34s} {d['stream']:>4s} {us:8.1f} {inst/1e6:8.2f} {thr/max(inst,1):6.1f} {d.get('smsp__issue_active.avg.pct_of_peak_sustained_active',0):5.1f} "
          f"{d.get('sm__warps_active.avg.pct_of_peak_sustained_active',0):5.1f} {red/1e6:8.3f} {atom/1e6:7.3f} {d.get('lts__t_sectors.sum',0)/1e6:8.2f} {dram:7.2f}")
    t = tot[d['name']]; t[0] += 1; t[1] += us; t[2] += inst; t[3] += red + atom
print()
for name, t in tot.items():
    print(f"{name[:40]:40s} x{t[0]:3d} {t[1]:9.1f} us {t[2]/1e6:9.2f} Minst {t[3]/1e6:8.3f} M red+atom sectors")
