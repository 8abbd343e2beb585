#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "swin_attn_core" -x > gpurun_out/c2_core_tests.log 2>&1; tail -15 gpurun_out/c2_core_tests.log | cut -c1-250
if ! grep -q " passed" gpurun_out/c2_core_tests.log || grep -q "failed" gpurun_out/c2_core_tests.log; then exit 1; fi
timeout 900 python -m pytest tests/test_parity_gpu.py -q -rxf -k "fp16 or bf16" > gpurun_out/c2_parity.log 2>&1; tail -5 gpurun_out/c2_parity.log | cut -c1-300
FUSE_ATTN=1 PREC=fp16 timeout 600 python tools/step_breakdown.py > gpurun_out/c2_breakdown_fused.log 2>&1; cat gpurun_out/c2_breakdown_fused.log
FUSE_ATTN=0 PREC=fp16 timeout 600 python tools/step_breakdown.py > gpurun_out/c2_breakdown_core_only.log 2>&1; cat gpurun_out/c2_breakdown_core_only.log
