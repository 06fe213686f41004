"""Per-launch table from an `ncu --csv --metrics ...` log: id, kernel, microseconds and the other metrics asked for.
usage: launch_times.py launches.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, mi, vi, idi = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID"))
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[idi], r[ki]), {})[r[mi]] = r[vi].replace(",", "")
for (i, k), m in d.items():
    rest = " ".join(f"{n.split('.')[0].replace('__', '_')}={float(v):.3g}" for n, v in m.items() if n != "gpu__time_duration.sum")
    print(f"{i:>4s} {k.split('(')[0][:44]:44s} {float(m['gpu__time_duration.sum']) / 1e3:9.1f} us  {rest}")
