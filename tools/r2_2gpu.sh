#!/bin/bash
# 2-GPU evidence: the NCCL gradient-equality test and a 2-rank bench line next to a 1-GPU line on the same box
cd /root/repo; mkdir -p gpurun_out
timeout -k 5 120 python -m pytest tests/test_multigpu_gpu.py -q -m gpu -x 2>&1 | tail -3 > gpurun_out/r2_2gpu_tests.txt; cat gpurun_out/r2_2gpu_tests.txt
timeout -k 5 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_A1_2gpu.json 2> gpurun_out/r2_A1_2gpu.err
timeout -k 5 120 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2_A1_1gpu_samebox.json 2>/dev/null
python - <<'PY'
import json
a=json.load(open('gpurun_out/r2_A1_2gpu.json')); b=json.load(open('gpurun_out/r2_A1_1gpu_samebox.json'))
print('2gpu', a['value'], a['ms_per_step'], '1gpu', b['value'], b['ms_per_step'], 'eff', a['value']/(2*b['value']))
PY
