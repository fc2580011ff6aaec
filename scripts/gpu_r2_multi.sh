#!/bin/bash
# round 2: N-GPU bench line (sharded parity self-check, cfg2 scan + auto, cfg3, cfg4) under torchrun, one rank per GPU
N=${1:-8}
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/nccl_parity.py 2>&1 | grep -v Warning | tail -8
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/r2g_bench_${N}gpu.json 2> gpurun_out/r2g_bench_${N}gpu.err
tail -c 400 gpurun_out/r2g_bench_${N}gpu.err
python - <<PY
import json
l=json.loads([x for x in open('gpurun_out/r2g_bench_${N}gpu.json').read().strip().splitlines() if x.startswith('{')][-1])
print('parity', l.get('sharded_parity'), 'scan', round(l['value'],1), 'e2e', round(l['e2e']['value'],1), 'frac', round(l['roofline']['frac'],3), l['clocks'])
a=l['auto_mode']; print('auto', round(a['value'],1), 'e2e', round(a['e2e']['value'],1))
for k in ('cfg3_b1','cfg3_b256'):
    if k in l:
        for m in ('auto','scan'):
            x=l[k][m]; print(k, m, x['mode_used'], 'q/s', round(x['value'],1), 'ms', round(x['ms_per_step'],3), 'e2e', round(x['e2e']['value'],1))
if 'cfg4_dense' in l:
    d=l['cfg4_dense']; print('cfg4', round(d['value'],1), 'TF/GPU', round(d['roofline']['achieved'],1), 'call TF', round(d['roofline']['whole_call_tflops_per_gpu'],1), d['clocks'], d.get('without_threshold_sharing'))
print(l.get('errors'))
PY
