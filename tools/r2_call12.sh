#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "mlp" 2>&1 | tail -4
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x -k "fp16 or bf16" 2>&1 | tail -3
PREC=fp16 timeout 600 python tools/step_breakdown.py > gpurun_out/c12_breakdown.log 2>&1; cat gpurun_out/c12_breakdown.log | head -8
FUSE_WIDTHS=128 PREC=fp16 timeout 600 python tools/step_breakdown.py > gpurun_out/c12_breakdown_w128.log 2>&1; head -7 gpurun_out/c12_breakdown_w128.log
