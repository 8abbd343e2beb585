#!/bin/bash
# Final round-2 evidence run on the final code (1 GPU): full default bench, reference arm, launch list + DRAM traffic under ncu, attention
# half-block microbench, SwinV2-B bench, smoke.
set -x
timeout 900 python bench.py > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/z_ref.json 2> gpurun_out/z_ref.err
timeout 300 python __graft_entry__.py smoke > gpurun_out/z_smoke.log 2>&1; tail -1 gpurun_out/z_smoke.log | cut -c1-200
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 1 --warmup 2 --no-extras --no-cpu-baseline --no-graph > gpurun_out/z_ncu.log 2>&1
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'gemm_|swin_attn' --csv --log-file gpurun_out/z_dram.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-graph > gpurun_out/z_ncu_dram.log 2>&1
PREC=fp16 timeout 400 python tools/bench_attn_halfblock.py > gpurun_out/z_halfblock.jsonl 2> gpurun_out/z_halfblock.err
timeout 300 python bench.py --variant swinv2_b --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/z_v2b.json 2> gpurun_out/z_v2b.err
timeout 200 python tools/bench_gemm_vs_cublas.py > gpurun_out/z_gemm_vs_cublas.txt 2>&1
python - <<'P'
import json
d=[json.loads(l) for l in open('gpurun_out/z_bench.json') if l.startswith('{')][0]
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'], 'bf16', d.get('bf16'))
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], 'all', d['roofline_all_linear']['achieved'], d['roofline_all_linear']['frac'])
print('check', d['check'], 'cpu', d.get('cpu_baseline',{}).get('value'), 'ft', d.get('finetune',{}).get('value'))
P
cut -c1-250 gpurun_out/z_ref.json gpurun_out/z_v2b.json; wc -l gpurun_out/z_launches.csv gpurun_out/z_dram.csv
