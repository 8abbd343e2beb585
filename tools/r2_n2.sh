set -x
python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -x -q -m gpu > gpurun_out/n2_tests.log 2>&1
tail -3 gpurun_out/n2_tests.log
for m in 1 2 1 2; do
CSVIT_RED_ADD=$m python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>> gpurun_out/n2_bench.err | cut -c1-200 | tee -a gpurun_out/n2_bench_mode$m.json
done
python tools/bench_mlp.py > gpurun_out/n2_mlp_red1.txt 2>&1; CSVIT_RED_ADD=2 python tools/bench_mlp.py > gpurun_out/n2_mlp_red2.txt 2>&1
tail -8 gpurun_out/n2_mlp_red1.txt gpurun_out/n2_mlp_red2.txt
