#!/bin/bash
# Final round-1 evidence run (1 GPU): bench (both arms), launch lists (inference + finetune + SwinV2), step breakdowns, microbenchmarks.
set -x
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
timeout 200 python tools/step_breakdown.py > gpurun_out/final_breakdown.txt 2>&1
timeout 300 python bench.py --workload temporal --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/final_temporal.json 2> gpurun_out/final_temporal.err
timeout 300 python bench.py --workload finetune --batch 32 --steps 10 --warmup 3 > gpurun_out/final_ft_b32.json 2> gpurun_out/final_ft_b32.err
timeout 300 python bench.py --variant swinv2_b --steps 10 --warmup 3 --cpu-sample 8 > gpurun_out/final_v2b.json 2> gpurun_out/final_v2b.err
VARIANT=swinv2_b timeout 200 python tools/step_breakdown.py > gpurun_out/final_breakdown_v2b.txt 2>&1
timeout 300 python tools/bench_attn_halfblock.py > gpurun_out/final_halfblock.jsonl 2> gpurun_out/final_halfblock.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 2 --no-extras --no-cpu-baseline --no-graph > gpurun_out/final_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/final_launches_v2b.csv python bench.py --variant swinv2_b --steps 1 --warmup 2 --no-extras --no-cpu-baseline --no-graph > gpurun_out/final_ncu_v2b.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/final_launches_ft.csv python bench.py --workload finetune --steps 1 --warmup 2 --no-graph > gpurun_out/final_ncu_ft.log 2>&1
cat gpurun_out/final_bench.json; cat gpurun_out/final_ref.json; tail -12 gpurun_out/final_breakdown.txt; cat gpurun_out/final_v2b.json; cut -c1-200 gpurun_out/final_ft_b32.json gpurun_out/final_temporal.json
wc -l gpurun_out/final_launches.csv gpurun_out/final_launches_ft.csv gpurun_out/final_launches_v2b.csv
