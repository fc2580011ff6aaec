#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-p}
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:inv_select -s 4 -c 1 -f -o gpurun_out/prof_select_$TAG \
    python bench.py --steps 1 --warmup 1 --mode inverted --no-cpu-baseline --batch 8 > gpurun_out/ncu_select_$TAG.log 2>&1; echo "ncu select rc=$?"
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:inv_accum -s 4 -c 1 -f -o gpurun_out/prof_accum_$TAG \
    python bench.py --steps 1 --warmup 1 --mode inverted --no-cpu-baseline --batch 8 > gpurun_out/ncu_accum_$TAG.log 2>&1; echo "ncu accum rc=$?"
