#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_reference_scripts.py -m gpu -q -x 2>&1 | tail -3
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/c8_launches.csv python bench.py --steps 1 --warmup 2 --no-extras --no-cpu-baseline --no-graph > gpurun_out/c8_ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/c8_launches.csv > gpurun_out/c8_launches_summary.txt; head -30 gpurun_out/c8_launches_summary.txt
python tools/last_step_launches.py gpurun_out/c8_launches.csv > gpurun_out/c8_last_step.txt; wc -l gpurun_out/c8_last_step.txt
