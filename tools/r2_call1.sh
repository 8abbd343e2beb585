#!/bin/bash
# round-2 GPU call 1: fused attention wired in, new parity tests, precision diag, step breakdown
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -q -k "fused" -x > gpurun_out/c1_fused_tests.log 2>&1; tail -3 gpurun_out/c1_fused_tests.log
python -m pytest tests/test_parity_gpu.py -q -rxXf > gpurun_out/c1_parity.log 2>&1; tail -40 gpurun_out/c1_parity.log
python tools/diag_precision.py > gpurun_out/c1_diag.log 2>&1; tail -30 gpurun_out/c1_diag.log
FUSE_ATTN=0 PREC=fp16 python tools/step_breakdown.py > gpurun_out/c1_breakdown_unfused.log 2>&1; cat gpurun_out/c1_breakdown_unfused.log
FUSE_ATTN=1 PREC=fp16 python tools/step_breakdown.py > gpurun_out/c1_breakdown_fused.log 2>&1; cat gpurun_out/c1_breakdown_fused.log
