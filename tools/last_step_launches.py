"""Per-launch durations of the LAST step in an `ncu --metrics gpu__time_duration.sum --csv` launch list of
`bench.py --steps 1 --warmup W --no-graph --no-extras` (the step is delimited by the patch_im2col launches).
usage: python tools/last_step_launches.py <csv>"""
import csv, re, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = [r for r in csv.DictReader(lines) if r["Metric Name"] == "gpu__time_duration.sum"]
names = [re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("csvit::", "").strip() for r in rows]
us = [float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}[r["Metric Unit"]] for r in rows]
starts = [i for i, n in enumerate(names) if n.startswith("patch_im2col")]
lo = starts[-1]
tot = sum(us[lo:])
print(f"# last step: launches {lo}..{len(rows) - 1} ({len(rows) - lo}), {tot / 1e3:.3f} ms summed (cold-cache, serialised under ncu)")
run, prev = [], None
for n, t in zip(names[lo:], us[lo:]):
    print(f"{t:9.1f} us  {n[:110]}")
