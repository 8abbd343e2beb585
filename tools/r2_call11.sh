#!/bin/bash
# 8-GPU confirmation: allreduce kernel check, headline bench with the finetune sub-record, finetune with NCCL for comparison
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29721 tools/multi_gpu_check.py > gpurun_out/c11_check8.log 2>&1; echo "check rc=$?"; grep "^{" gpurun_out/c11_check8.log | grep -v '"values"'; grep '"values"' gpurun_out/c11_check8.log | grep -c '"ok": true'
timeout 600 $TR --master-port 29722 bench.py --gpus 8 --steps 20 --no-cpu-baseline > gpurun_out/c11_bench_8gpu.json 2> gpurun_out/c11_bench_8gpu.err; echo "bench8 rc=$?"; grep "^{" gpurun_out/c11_bench_8gpu.json | cut -c1-200; tail -2 gpurun_out/c11_bench_8gpu.err | cut -c1-300
timeout 400 $TR --master-port 29723 bench.py --gpus 8 --workload finetune --steps 10 --warmup 2 --comm nccl > gpurun_out/c11_ft8_nccl.json 2> gpurun_out/c11_ft8_nccl.err; echo "ft nccl rc=$?"; grep "^{" gpurun_out/c11_ft8_nccl.json | cut -c1-120
