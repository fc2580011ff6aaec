#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dense.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/prof_dense_shard.py 2>&1 | tail -1
timeout 600 python scripts/bench_configs.py cfg4 2>/dev/null | cut -c1-330 | tee gpurun_out/configs_r1i_cfg4.jsonl
