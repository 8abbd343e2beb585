#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29701 tools/multi_gpu_check.py > gpurun_out/c6_check.log 2>&1; echo "check rc=$?"; grep "^{" gpurun_out/c6_check.log; tail -5 gpurun_out/c6_check.log | grep -v "^{" | cut -c1-300
timeout 400 $TR --master-port 29702 bench.py --gpus 2 --workload finetune --steps 10 --warmup 2 --comm symm > gpurun_out/c6_ft_symm.json 2> gpurun_out/c6_ft_symm.err; echo "symm rc=$?"; cat gpurun_out/c6_ft_symm.json | cut -c1-1500; tail -3 gpurun_out/c6_ft_symm.err | cut -c1-300
timeout 400 $TR --master-port 29703 bench.py --gpus 2 --workload finetune --steps 10 --warmup 2 --comm nccl > gpurun_out/c6_ft_nccl.json 2> gpurun_out/c6_ft_nccl.err; echo "nccl rc=$?"; cat gpurun_out/c6_ft_nccl.json | cut -c1-1500; tail -3 gpurun_out/c6_ft_nccl.err | cut -c1-300
timeout 300 python bench.py --workload finetune --steps 10 --warmup 2 > gpurun_out/c6_ft_1gpu.json 2> gpurun_out/c6_ft_1gpu.err; echo "1gpu rc=$?"; cat gpurun_out/c6_ft_1gpu.json | cut -c1-1200
