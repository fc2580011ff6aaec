#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_topk_pair -s 5 -c 1 -f -o gpurun_out/prof_dense_pair python scripts/prof_dense.py > gpurun_out/ncu_dense_pair.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_dense_pair.log
