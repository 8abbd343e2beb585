#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -rf > gpurun_out/c9_gpu_tests.log 2>&1; tail -6 gpurun_out/c9_gpu_tests.log | cut -c1-250
FUSE_ATTN=1 PREC=fp16 timeout 600 python tools/step_breakdown.py > gpurun_out/c9_breakdown.log 2>&1; cat gpurun_out/c9_breakdown.log
timeout 600 python bench.py --no-cpu-baseline --no-finetune-record > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err; cut -c1-400 gpurun_out/c9_bench.json; tail -2 gpurun_out/c9_bench.err
