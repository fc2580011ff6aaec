#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-d3}
timeout 600 python -m pytest tests/test_gpu_dense.py -m gpu -x -q 2>&1 | tail -5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_topk -s 5 -c 1 -f -o gpurun_out/prof_dense_$TAG python scripts/prof_dense.py > gpurun_out/ncu_dense_$TAG.log 2>&1; echo "ncu dense rc=$?"; tail -2 gpurun_out/ncu_dense_$TAG.log
