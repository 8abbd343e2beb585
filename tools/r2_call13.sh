#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "regex:gemm_|swin_attn" -c 2000 --csv --log-file gpurun_out/c13_dram.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline --no-graph > gpurun_out/c13_ncu.log 2>&1; tail -2 gpurun_out/c13_ncu.log
python tools/dram_traffic.py gpurun_out/c13_dram.csv gpurun_out/c13_dram_traffic.json 3 | head -40
python tools/ncu_attn.py && timeout 600 ncu --set full --clock-control none --import-source on -k regex:swin_attn -c 2 -f -o gpurun_out/c13_attn python tools/ncu_attn.py > gpurun_out/c13_ncu_attn.log 2>&1; tail -2 gpurun_out/c13_ncu_attn.log
timeout 300 python tools/bench_attn_kernels.py 2>&1 | tee gpurun_out/c13_attn_bench.log | tail -20
