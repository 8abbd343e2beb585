#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/c10_bench_1gpu.json 2> gpurun_out/c10_bench_1gpu.err; echo "1gpu rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c10_bench_1gpu.json')); print(d['value'], d.get('finetune'))"; tail -3 gpurun_out/c10_bench_1gpu.err | cut -c1-300
timeout 600 $TR --master-port 29711 bench.py --gpus 2 --steps 10 --no-cpu-baseline > gpurun_out/c10_bench_2gpu.json 2> gpurun_out/c10_bench_2gpu.err; echo "2gpu rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c10_bench_2gpu.json')); print(d['value'], d['e2e']['value'], d.get('finetune'))"; tail -3 gpurun_out/c10_bench_2gpu.err | cut -c1-300
