#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -rxf > gpurun_out/c5_gpu_tests.log 2>&1; tail -6 gpurun_out/c5_gpu_tests.log | cut -c1-250
timeout 300 python __graft_entry__.py smoke > gpurun_out/c5_smoke.log 2>&1; tail -2 gpurun_out/c5_smoke.log
timeout 600 python bench.py > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; cat gpurun_out/c5_bench.json; tail -3 gpurun_out/c5_bench.err
PREC=fp16 timeout 600 python tools/bench_attn_halfblock.py > gpurun_out/c5_halfblock.jsonl 2>&1; cat gpurun_out/c5_halfblock.jsonl | cut -c1-330
