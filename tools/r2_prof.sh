#!/bin/bash
# dev: cycle counters of the fused backward (PROF build on the box only; the shipped .so is not touched here)
cd /root/repo/nerf_pytorch_b200/csrc && rm -f mlp_tc_bwd.o && make EXTRA=-DNERFB200_PROF -j8 > /dev/null 2>&1
cd /root/repo; mkdir -p gpurun_out
python tools/bwd_prof.py > gpurun_out/r2_bwd_prof_A1.txt 2>&1
ARCH=A0 python tools/bwd_prof.py > gpurun_out/r2_bwd_prof_A0.txt 2>&1
cat gpurun_out/r2_bwd_prof_A1.txt gpurun_out/r2_bwd_prof_A0.txt
