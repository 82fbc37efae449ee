#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
PSNR_ITERS=5000 PSNR_BATCH=1024 PSNR_SEEDS=3 PSNR_REF_BUDGET_S=400 timeout -k 5 900 python tools/psnr_scale.py > gpurun_out/r2_psnr_scale.log 2>&1
tail -12 gpurun_out/r2_psnr_scale.log
