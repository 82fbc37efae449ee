#!/bin/bash
# 2-GPU evidence: the NCCL gradient-equality test and a 2-rank bench line
cd /root/repo; mkdir -p gpurun_out
timeout -k 5 200 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -x -rs 2>&1 | tail -8 > gpurun_out/r2_2gpu_tests.txt; cat gpurun_out/r2_2gpu_tests.txt
timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_A1_2gpu.json 2> gpurun_out/r2_A1_2gpu.err
head -c 700 gpurun_out/r2_A1_2gpu.json; echo; tail -3 gpurun_out/r2_A1_2gpu.err
timeout -k 5 200 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2_A1_1gpu_samebox.json 2>/dev/null; head -c 400 gpurun_out/r2_A1_1gpu_samebox.json
