#!/bin/bash
# final-state capture for profiles/: tests, bench lines, ncu launch list, ncu --set full of the four MLP kernels
cd /root/repo; mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/final_tests.txt
timeout 300 python bench.py > gpurun_out/final_A1.json 2> gpurun_out/final_A1.err
timeout 300 python bench.py --arch A0 > gpurun_out/final_A0.json 2> gpurun_out/final_A0.err
timeout 300 python bench.py --arch A2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_A2.json 2> gpurun_out/final_A2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_ncu_b.log 2>&1
IMPL=1 STASH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_ -f -o gpurun_out/final_full python tools/ncu_fwd.py > gpurun_out/final_ncu_full.log 2>&1
cat gpurun_out/final_tests.txt; head -c 600 gpurun_out/final_A1.json; echo; ls -la gpurun_out | tail -12
