set -x
python tools/diag_epilogue.py > gpurun_out/n1_epi_red1.txt 2>&1
CSVIT_RED_ADD=0 python tools/diag_epilogue.py > gpurun_out/n1_epi_red0.txt 2>&1
python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -x -q -m gpu > gpurun_out/n1_tests.log 2>&1
tail -3 gpurun_out/n1_tests.log
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/n1_bench_red1.json 2> gpurun_out/n1_bench.err
CSVIT_RED_ADD=0 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/n1_bench_red0.json 2>> gpurun_out/n1_bench.err
cat gpurun_out/n1_epi_red1.txt gpurun_out/n1_epi_red0.txt
cat gpurun_out/n1_bench_red1.json gpurun_out/n1_bench_red0.json | cut -c1-300
