#!/bin/bash
# Final round-1 evidence run (1 GPU): bench (both arms), launch lists (inference + finetune), step breakdown, microbench.
set -x
timeout 600 python bench.py --steps 30 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
timeout 200 python tools/step_breakdown.py > gpurun_out/final_breakdown.txt 2>&1
timeout 300 python tools/bench_attn_halfblock.py > gpurun_out/final_halfblock.jsonl 2> gpurun_out/final_halfblock.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 2 --no-extras --no-cpu-baseline --no-graph > gpurun_out/final_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/final_launches_ft.csv python bench.py --workload finetune --steps 1 --warmup 2 --no-graph > gpurun_out/final_ncu_ft.log 2>&1
cat gpurun_out/final_bench.json; cat gpurun_out/final_ref.json; tail -12 gpurun_out/final_breakdown.txt; wc -l gpurun_out/final_launches.csv gpurun_out/final_launches_ft.csv
