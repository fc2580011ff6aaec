#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-d5}
timeout 300 python -m pytest tests/test_gpu_dense.py -m gpu -x -q 2>&1 | tail -8
echo "--- single-CTA path for comparison"
VSEARCH_B200_DENSE_PAIR=0 timeout 600 python scripts/bench_configs.py cfg4 2>gpurun_out/configs_${TAG}_cfg4_single.err | cut -c1-330
echo "--- pair path"
timeout 600 python scripts/bench_configs.py cfg4 > gpurun_out/configs_${TAG}_cfg4.jsonl 2> gpurun_out/configs_${TAG}_cfg4.err; echo "cfg4 rc=$?"; cut -c1-330 gpurun_out/configs_${TAG}_cfg4.jsonl; tail -3 gpurun_out/configs_${TAG}_cfg4.err
