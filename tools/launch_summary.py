"""Reduce an `ncu --metrics gpu__time_duration.sum --csv` launch list to per-kernel totals:
   python tools/launch_summary.py gpurun_out/<tag>_launches.csv > profiles/<name>.csv"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik]).replace("b200::<", "").strip()
    v = float(r[iv].replace(",", ""))
    v = v / 1e3 if r[iu] in ("ns", "nsecond") else v
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
total = sum(a[1] for a in agg.values())
n = sum(a[0] for a in agg.values())
print(f"# launches {n}, total {total:.1f} us")
print("kernel,launches,total_us,share,avg_us")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"\"{k}\",{c},{t:.1f},{t / total:.4f},{t / c:.2f}")
