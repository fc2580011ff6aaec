#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-p}
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:inv_search -s 1 -c 1 -f -o gpurun_out/prof_invsearch_$TAG \
    python bench.py --steps 1 --warmup 1 --mode inverted --no-cpu-baseline --batch 16 > gpurun_out/ncu_invsearch_$TAG.log 2>&1; echo "ncu inv_search rc=$?"
tail -3 gpurun_out/ncu_invsearch_$TAG.log
