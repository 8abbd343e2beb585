#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c14_gpu_tests.log 2>&1; tail -3 gpurun_out/c14_gpu_tests.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke > gpurun_out/c14_smoke.log 2>&1; tail -1 gpurun_out/c14_smoke.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench.err; python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/c14_bench.json') if l.startswith('{')][0]
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['clocks'], 'bf16', d.get('bf16'))
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['launches_per_step'], 'all', d['roofline_all_linear']['achieved'], d['roofline_all_linear']['frac'])
print('attn', d['roofline_attention']['achieved'], d['roofline_attention']['frac'], 'fused', d['roofline_attention_fused']['achieved'], d['roofline_attention_fused']['tflops'])
print('check', d['check'], 'cpu', d.get('cpu_baseline',{}).get('value'), 'ft', d.get('finetune',{}).get('value'))
"; tail -2 gpurun_out/c14_bench.err | cut -c1-200
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c14_ref.json 2>/dev/null; cut -c1-300 gpurun_out/c14_ref.json
PREC=fp16 timeout 600 python tools/bench_attn_halfblock.py > gpurun_out/c14_halfblock.jsonl 2>/dev/null; cut -c1-200 gpurun_out/c14_halfblock.jsonl
