"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share.
usage: launch_summary.py launches.csv [out.csv]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split('(')[0][:70]
    v = float(r[vi].replace(',', ''))
    v = v / 1000 if r[ui] == 'ns' else (v * 1000 if r[ui] == 'ms' else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
out = [['kernel', 'launches', 'total_us', 'avg_us', 'share_pct']]
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append([k, n, round(t, 1), round(t / n, 1), round(100 * t / tot, 2)])
for r in out:
    print(f'{str(r[0]):72s} {str(r[1]):>8s} {str(r[2]):>12s} {str(r[3]):>10s} {str(r[4]):>8s}')
if len(sys.argv) > 2:
    csv.writer(open(sys.argv[2], 'w')).writerows(out)
