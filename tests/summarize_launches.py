"""ncu --csv launch list -> per-kernel summary (count, total us, share, avg us).  usage: summarize_launches.py in.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    us = v / 1e3 if r[ui] in ("ns", "nsecond") else (v if r[ui] in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"goat::\(anonymous namespace\)::", "", name)
    agg[name[:110]][0] += 1
    agg[name[:110]][1] += us
tot = sum(v[1] for v in agg.values())
print("kernel,count,total_us,share,avg_us")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%s,%d,%.1f,%.1f%%,%.1f" % (k, n, us, 100.0 * us / tot, us / n))
print("total,%d,%.1f,100%%," % (sum(v[0] for v in agg.values()), tot))
