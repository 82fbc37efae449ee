#!/bin/bash
# dev: forward cycle counters for the normal kernel and two timing experiments (PROF builds on the box only)
cd /root/repo; mkdir -p gpurun_out
for e in 0 1 2; do
  cd /root/repo/nerf_pytorch_b200/csrc && rm -f mlp_tc.o && make EXTRA="-DNERFB200_PROF -DNERFB200_EXP=$e" -j8 > /dev/null 2>&1
  cd /root/repo
  echo "=== EXP $e (0 normal, 1 epilogue without arithmetic, 2 no MMAs)" >> gpurun_out/r2_fwd_prof_exp.txt
  timeout -k 5 120 python tools/fwd_prof.py >> gpurun_out/r2_fwd_prof_exp.txt 2>&1
done
cat gpurun_out/r2_fwd_prof_exp.txt
