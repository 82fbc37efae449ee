#!/bin/bash
# dev: backward cycle counters (PROF build on the box only; the shipped .so is not touched here)
cd /root/repo; mkdir -p gpurun_out
cd /root/repo/nerf_pytorch_b200/csrc && rm -f mlp_tc_bwd.o mlp_tc.o && make EXTRA="-DNERFB200_PROF" -j8 > /dev/null 2>&1
cd /root/repo
timeout -k 5 120 python tools/bwd_prof.py > gpurun_out/r2_bwd_prof_A1.txt 2>&1
cat gpurun_out/r2_bwd_prof_A1.txt
