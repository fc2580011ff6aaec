#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-e}
lst() { name=$1; shift
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}_$name.csv \
    python bench.py --steps 1 --warmup 1 --mode inverted --no-cpu-baseline --batch 16 "$@" > gpurun_out/ncu_launches_${TAG}_$name.log 2>&1; echo "ncu list $name rc=$?"
  grep -E "inv_accum|inv_select" gpurun_out/launches_${TAG}_$name.csv | awk -F'","' '{print $5, $NF}' | tail -4; }
lst q64 --qnnz 64
lst q0 --qnnz 0
VSEARCH_B200_DEBUG_NOZERO=1 lst q0_nozero --qnnz 0
VSEARCH_B200_DEBUG_NOZERO=1 lst q64_nozero --qnnz 64
timeout 1500 python scripts/bench_configs.py cfg4 > gpurun_out/configs_${TAG}_cfg4.jsonl 2> gpurun_out/configs_${TAG}_cfg4.err; echo "cfg4 rc=$?"; cat gpurun_out/configs_${TAG}_cfg4.jsonl | cut -c1-500; tail -5 gpurun_out/configs_${TAG}_cfg4.err
