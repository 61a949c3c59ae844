"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name.  python tools/launch_agg.py file.csv [top]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v, u = float(r[vi].replace(",", "")), r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
    a = agg.setdefault(r[ki][:80], [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
for k, (c, ms) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print(f"{ms:9.3f} ms  {c:5d}x  {100 * ms / tot:5.1f}%  {k}")
print(f"total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches")
