#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_search.py tests/test_gpu_retriever.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python scripts/bench_configs.py cfg1 cfg3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if d.get('mode') in ('inverted',): print(d['config'],d['B'],d['mode'],round(d['qps']),round(d['kernel_ms'],3))"
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --mode inverted --rows 2626916 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('2.6M-row shard inverted', round(d['value']), 'e2e', round(d['e2e']['value']))"
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --mode inverted 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('21M inverted', round(d['value']), 'e2e', round(d['e2e']['value']))"
