"""Per-kernel DRAM traffic of the LAST step from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv`
on `bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-graph`.  Writes profiles/r2_dram_traffic.json: for the GEMM engine
(gemm_pair_kernel + gemm_tc_kernel), swin_attn_core_kernel and swin_attn_fused_kernel the launches per step and the DRAM bytes per launch
(bench.py reads them for `roofline*.traffic`).   usage: python tools/dram_traffic.py <csv> <out.json>"""
import csv, json, re, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
by_id = {}
for r in rows:
    d = by_id.setdefault(int(r["ID"]), {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("csvit::", "").strip()})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
    d[r["Metric Name"]] = v * scale
ids = sorted(by_id)
names = [by_id[i]["name"] for i in ids]
# the capture is filtered to these kernels: a step starts with the patch-embedding GEMM (gemm_tc_kernel<128, 0, 1> / <128, 1, 1>) - find the
# last repetition of the per-step launch pattern by its length
groups = {"gemm": ("gemm_pair_kernel", "gemm_tc_kernel"), "gemm_pair": ("gemm_pair_kernel",), "attn_core": ("swin_attn_core_kernel",),
          "attn_fused": ("swin_attn_fused_kernel",)}
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_|swin_attn python bench.py "
                 "--steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-graph (last of the three passes of the capture)", "workload": "swin_b spatial predict_batch, batch 256, fp16"}
for key, pats in groups.items():
    sel = [i for i, n in zip(ids, names) if any(n.startswith(p) for p in pats)]
    per_step = len(sel) // steps
    last = sel[-per_step:]
    rd = sum(by_id[i].get("dram__bytes_read.sum", 0.0) for i in last)
    wr = sum(by_id[i].get("dram__bytes_write.sum", 0.0) for i in last)
    us = sum(by_id[i].get("gpu__time_duration.sum", 0.0) for i in last)
    out[key] = {"launches_per_step": per_step, "dram_read_bytes_per_step": rd, "dram_write_bytes_per_step": wr,
                "traffic_bytes_per_launch": (rd + wr) / max(per_step, 1), "us_per_step_under_ncu": us}
try:      # keep the tensor-pipe readings of the `ncu --set full` captures (profiles/r2_ncu_attention_key_metrics.txt) that live in the same file
    old = json.load(open(sys.argv[2]))
    for key in groups:
        for k, v in old.get(key, {}).items():
            if k.startswith("tensor_pipe"):
                out[key][k] = v
except Exception:
    pass
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
