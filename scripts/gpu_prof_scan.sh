#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-ps}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 1 -c 1 -f -o gpurun_out/prof_scan_$TAG \
    python bench.py --steps 1 --warmup 1 --mode scan --no-auto --no-cpu-baseline --batch 16 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
