#!/bin/bash
# ncu --set full (with source) of the three tcgen05 MLP kernels at the fine-pass size
cd /root/repo; mkdir -p gpurun_out
IMPL=1 STASH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_ -f -o gpurun_out/r2_full python tools/ncu_fwd.py > gpurun_out/r2_ncu_full.log 2>&1
tail -3 gpurun_out/r2_ncu_full.log; ls -la gpurun_out/r2_full.ncu-rep
