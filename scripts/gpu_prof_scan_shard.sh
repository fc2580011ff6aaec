#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 1 -c 1 -f -o gpurun_out/prof_scan_shard \
    python bench.py --steps 1 --warmup 1 --mode scan --no-auto --no-cpu-baseline --batch 64 --rows 2626916 > gpurun_out/ncu_scan_shard.log 2>&1; echo "ncu rc=$?"
