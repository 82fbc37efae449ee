#!/bin/bash
# tests of the tcgen05 kernels first (hard-bounded: a deadlocked kernel ignores SIGTERM), then the bench line, then all tests
cd /root/repo; mkdir -p gpurun_out
timeout -k 5 150 python -m pytest tests/test_tc_gpu.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2_tc_tests.txt
cat gpurun_out/r2_tc_tests.txt
if ! grep -q " passed" gpurun_out/r2_tc_tests.txt || grep -q "failed\|error\|Error" gpurun_out/r2_tc_tests.txt; then echo "TC TESTS FAILED"; exit 1; fi
timeout -k 5 200 python bench.py --no-cpu-baseline > gpurun_out/r2b_A1.json 2> gpurun_out/r2b_A1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_A1.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['ms_per_step'], d['fwd_only']['ms_per_step'], d['roofline_fwd']['ms'], d['roofline_fwd']['ms_training_forward'], d['roofline_bwd']['ms'], d['loss'])
PY
tail -3 gpurun_out/r2b_A1.err
timeout -k 5 400 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r2_tests.txt; cat gpurun_out/r2_tests.txt
