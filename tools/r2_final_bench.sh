#!/bin/bash
# final bench lines of the round: A1 (headline, with CPU baseline + GPU-eager yardstick), A0, configs 3 and 4, A2
cd /root/repo; mkdir -p gpurun_out
timeout -k 5 400 python bench.py > gpurun_out/r2_final_A1.json 2> gpurun_out/r2_final_A1.err
timeout -k 5 200 python bench.py --arch A0 --no-cpu-baseline > gpurun_out/r2_final_A0.json 2> gpurun_out/r2_final_A0.err
timeout -k 5 200 python bench.py --config 3 --no-cpu-baseline > gpurun_out/r2_final_cfg3.json 2> gpurun_out/r2_final_cfg3.err
timeout -k 5 200 python bench.py --config 4 --no-cpu-baseline > gpurun_out/r2_final_cfg4.json 2> gpurun_out/r2_final_cfg4.err
timeout -k 5 200 python bench.py --arch A2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_A2.json 2> gpurun_out/r2_final_A2.err
for f in A1 A0 cfg3 cfg4 A2; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/r2_final_{f}.json'))
    print(f, round(d['value']), d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'fwd', d['fwd_only']['ms_per_step'], 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'eager', d.get('torch_eager_gpu'), d['clocks'])
except Exception as e:
    print(f, 'FAILED', e); print(open(f'gpurun_out/r2_final_{f}.err').read()[-600:])
PY
done
