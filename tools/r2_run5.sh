#!/bin/bash
# what the driver runs at round end, in short: reference arm (CPU), default bench line
cd /root/repo; mkdir -p gpurun_out
timeout -k 5 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_ref_arm.json 2> gpurun_out/r2_ref_arm.err; head -c 900 gpurun_out/r2_ref_arm.json; echo; tail -2 gpurun_out/r2_ref_arm.err
timeout -k 5 300 python bench.py > gpurun_out/r2_last_A1.json 2> gpurun_out/r2_last_A1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_last_A1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['cpu_baseline']['value'], d['torch_eager_gpu'].get('ms_per_step'), d['clocks'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline_fwd']['frac'], d['loss'])
PY
