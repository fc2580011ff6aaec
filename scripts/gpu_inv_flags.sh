#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for F in 0 1 2 3; do
  VSEARCH_B200_K3_FLAGS=$F timeout 600 python scripts/sweep_crossover.py cfg2_768 2>/dev/null | grep '"mode": "inverted"' | python -c "
import sys,json
print('flags $F', [(d['B'],d['qnnz'],d['qps'],d['kernel_ms']) for d in map(json.loads, sys.stdin)])"
done
