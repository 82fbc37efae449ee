#!/bin/bash
# dev: cycle counters of the fused backward (PROF build on the box only; the shipped .so is not touched here) + MMA rates
cd /root/repo; mkdir -p gpurun_out tools/bin
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I nerf_pytorch_b200/csrc -o tools/bin/mma_rate tools/mma_rate.cu && ./tools/bin/mma_rate > gpurun_out/r2_mma_rate.txt 2>&1
cat gpurun_out/r2_mma_rate.txt
cd /root/repo/nerf_pytorch_b200/csrc && rm -f mlp_tc_bwd.o && make EXTRA=-DNERFB200_PROF -j8 > /dev/null 2>&1
cd /root/repo
python tools/bwd_prof.py > gpurun_out/r2_bwd_prof_A1.txt 2>&1
cat gpurun_out/r2_bwd_prof_A1.txt
