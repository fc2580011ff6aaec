#!/bin/bash
# scan kernel change check: full GPU tests, then the scan bench on an 8-GPU-sized shard and on the full index
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for R in 2626916 21015324; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-auto --rows $R 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('rows $R scan', round(d['value'],1), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1), d['clocks'])"
done
