"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: share of GPU time, total us, launches.
usage: python tools/summarize_launches.py <csv> [first_id last_id]"""
import csv, re, sys
from collections import OrderedDict
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = [r for r in csv.DictReader(lines) if r["Metric Name"] == "gpu__time_duration.sum"]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
agg = OrderedDict()
for r in rows:
    i = int(r["ID"])
    if not (lo <= i <= hi):
        continue
    us = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[r["Metric Unit"]]
    name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
    if "csvit" not in name:
        name = "(torch: " + name[:70] + ")"
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
ours = sum(a[1] for k, a in agg.items() if not k.startswith("(torch"))
print(f"# {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.2f} ms; kernels of this library: {100 * ours / tot:.1f}% of the GPU time")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(35)]:
    print(f"{100 * a[1] / tot:6.2f}%  {a[1]:9.1f} us x{a[0]:4d}  {k}")
