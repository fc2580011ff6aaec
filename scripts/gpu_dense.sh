#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dense.py tests/test_gpu_retriever.py -m gpu -x -q 2>&1 | tail -8
for D in 0 1; do
  echo "pair dbg=$D: $(VSEARCH_B200_DENSE_DBG=$D timeout 200 python scripts/prof_dense.py 2>&1 | tail -1)"
done
echo "single: $(VSEARCH_B200_DENSE_PAIR=0 timeout 200 python scripts/prof_dense.py 2>&1 | tail -1)"
timeout 600 python scripts/bench_configs.py cfg4 > gpurun_out/configs_dense_cfg4.jsonl 2> gpurun_out/configs_dense_cfg4.err; echo "cfg4 rc=$?"; cut -c1-330 gpurun_out/configs_dense_cfg4.jsonl; tail -3 gpurun_out/configs_dense_cfg4.err
