#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-s}
shift
for c in "$@"; do
  timeout 900 python scripts/sweep_crossover.py $c > gpurun_out/sweep_${TAG}_$c.jsonl 2> gpurun_out/sweep_${TAG}_$c.err; echo "$c rc=$?"
  tail -2 gpurun_out/sweep_${TAG}_$c.err
  head -40 gpurun_out/sweep_${TAG}_$c.jsonl | cut -c1-260
done
