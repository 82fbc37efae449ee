#!/bin/bash
# ncu: launch list of the bench command + --set full (with source) of the three tcgen05 MLP kernels at the fine-pass size
cd /root/repo; mkdir -p gpurun_out
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_ncu_b.log 2>&1
IMPL=1 STASH=0 timeout -k 5 400 ncu --set full --clock-control none --import-source on -k regex:mlp_ -f -o gpurun_out/r2_final_full python tools/ncu_fwd.py > gpurun_out/r2_final_ncu_full.log 2>&1
tail -2 gpurun_out/r2_final_ncu_full.log; ls -la gpurun_out/r2_final_full.ncu-rep gpurun_out/r2_final_launches.csv
