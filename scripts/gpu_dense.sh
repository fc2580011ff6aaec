#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-d}
timeout 600 python -m pytest tests/test_gpu_dense.py -m gpu -x -q > gpurun_out/pytest_dense_$TAG.log 2>&1; echo "pytest dense rc=$?"; tail -30 gpurun_out/pytest_dense_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:inv_select -s 2 -c 1 -f -o gpurun_out/prof_select_$TAG \
    python bench.py --steps 1 --warmup 1 --mode inverted --no-cpu-baseline --batch 8 > gpurun_out/ncu_select_$TAG.log 2>&1; echo "ncu select rc=$?"
