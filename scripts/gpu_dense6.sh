#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
VSEARCH_B200_DEBUG=1 timeout 300 python scripts/prof_dense.py 2>&1 | tail -12
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_dense6.csv python scripts/prof_dense.py > /dev/null 2>&1
grep -E "dense_topk" gpurun_out/launches_dense6.csv | awk -F'","' '{print $5, $NF}' | tail -8
