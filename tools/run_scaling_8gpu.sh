#!/bin/bash
# 8-GPU evidence for BASELINE configs[1..3]: spatial forward, temporal forward, finetune step (one box, torchrun, NCCL).
set -x
N=${1:-8}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29601 bench.py --gpus $N --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/scale_spatial_${N}gpu.json 2> gpurun_out/scale_spatial_${N}gpu.err
timeout 300 $RUN --master-port 29602 bench.py --gpus $N --steps 20 --warmup 3 --no-extras --no-cpu-baseline --workload temporal > gpurun_out/scale_temporal_${N}gpu.json 2> gpurun_out/scale_temporal_${N}gpu.err
timeout 300 $RUN --master-port 29603 bench.py --gpus $N --steps 10 --warmup 3 --workload finetune > gpurun_out/scale_finetune_${N}gpu.json 2> gpurun_out/scale_finetune_${N}gpu.err
cat gpurun_out/scale_*_${N}gpu.json
