#!/bin/bash
# round-2 checkpoint capture: tests, bench lines (A1/A0), ncu launch list
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2_tests.txt
timeout 300 python bench.py > gpurun_out/r2_A1.json 2> gpurun_out/r2_A1.err
timeout 300 python bench.py --arch A0 --no-cpu-baseline > gpurun_out/r2_A0.json 2> gpurun_out/r2_A0.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_b.log 2>&1
cat gpurun_out/r2_tests.txt; head -c 3000 gpurun_out/r2_A1.json; echo; tail -5 gpurun_out/r2_A1.err; head -c 1500 gpurun_out/r2_A0.json
