#!/bin/bash
# round 2: sparse-query entry point, device-side auto decision, new tests; bench smoke + real bench
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py --rows 400000 --batch 64 --extras-rows 300000 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -2 > gpurun_out/r2d_bench_smoke.json; tail -c 600 gpurun_out/r2d_bench_smoke.json
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 300 gpurun_out/r2d_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2d_bench.json').read().strip().splitlines()[-1])
print('scan', l['value'], 'e2e', l['e2e']['value'], 'frac', l['roofline']['frac'], l['clocks'])
a=l['auto_mode']; print('auto', a['value'], 'e2e', a['e2e']['value'], a['e2e']['h2d_bytes_per_step'])
for k in ('cfg3_b1','cfg3_b256'):
    for m in ('auto','scan'):
        x=l[k][m]; print(k, m, x['mode_used'], 'q/s', round(x['value'],1), 'ms', round(x['ms_per_step'],3), 'kernel ms', round(x['kernel_ms_per_launch'],3), 'e2e', round(x['e2e']['value'],1))
d=l['cfg4_dense']; print('cfg4', d['value'], d['roofline']['achieved'], d['clocks'])
PY
