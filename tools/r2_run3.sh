#!/bin/bash
# hidden-256 tcgen05 forward: its tests first (hard-bounded), then all tcgen05 tests, the A1 / A2 bench lines, then everything
cd /root/repo; mkdir -p gpurun_out
timeout -k 5 150 python -m pytest tests/test_tc_gpu.py -q -m gpu -x -k "hidden_256 or unsupported" 2>&1 | tail -15 > gpurun_out/r2_tc256_tests.txt
cat gpurun_out/r2_tc256_tests.txt
if ! grep -q " passed" gpurun_out/r2_tc256_tests.txt || grep -q "failed\|error\|Error" gpurun_out/r2_tc256_tests.txt; then echo "TC256 TESTS FAILED"; exit 1; fi
timeout -k 5 200 python bench.py --arch A2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_A2.json 2> gpurun_out/r2b_A2.err
timeout -k 5 200 python bench.py --no-cpu-baseline > gpurun_out/r2b_A1.json 2> gpurun_out/r2b_A1.err
python - <<'PY'
import json
for f in ('A2','A1'):
    try:
        d=json.load(open(f'gpurun_out/r2b_{f}.json'))
        print(f, {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['ms_per_step'], 'fwd_only', d['fwd_only'], d['roofline_fwd']['ms'], d['roofline_fwd']['ms_training_forward'], d['roofline_bwd']['ms'])
    except Exception as e:
        print(f, 'FAILED', e); print(open(f'gpurun_out/r2b_{f}.err').read()[-800:])
PY
timeout -k 5 400 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r2_tests.txt; cat gpurun_out/r2_tests.txt
