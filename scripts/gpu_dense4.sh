#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-d4}
timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_retriever.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python scripts/prof_dense.py
timeout 1500 python scripts/bench_configs.py cfg4 > gpurun_out/configs_${TAG}_cfg4.jsonl 2> gpurun_out/configs_${TAG}_cfg4.err; echo "cfg4 rc=$?"; cat gpurun_out/configs_${TAG}_cfg4.jsonl | cut -c1-400; tail -3 gpurun_out/configs_${TAG}_cfg4.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_topk -s 5 -c 1 -f -o gpurun_out/prof_dense_$TAG python scripts/prof_dense.py > gpurun_out/ncu_dense_$TAG.log 2>&1; echo "ncu dense rc=$?"
