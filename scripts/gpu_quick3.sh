#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-q}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
run() { name=$1; shift
  timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; echo "bench $name rc=$?"
  python -c "
import json,sys; d=json.load(open('gpurun_out/bench_${TAG}_$name.json')); print({k:d[k] for k in ('value','ms_per_step')}, 'GB/s', round(d['roofline']['achieved'],1), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value'],1), d['config'].get('mode_used'), d['clocks'])"; tail -3 gpurun_out/bench_${TAG}_$name.err; }
run scan --mode scan
run inv64 --mode inverted --qnnz 64
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}_inv.csv \
    python bench.py --steps 1 --warmup 1 --mode inverted --no-cpu-baseline --batch 32 > gpurun_out/ncu_launches_${TAG}_inv.log 2>&1; echo "ncu list inv rc=$?"
grep -E "inv_accum|inv_select|inv_extract|merge_topk|prep_query" gpurun_out/launches_${TAG}_inv.csv | awk -F'","' '{print $5, $NF}' | tail -7
timeout 900 python scripts/bench_configs.py cfg1 cfg3 > gpurun_out/configs_$TAG.jsonl 2> gpurun_out/configs_$TAG.err; echo "configs rc=$?"; cat gpurun_out/configs_$TAG.jsonl | cut -c1-420; tail -3 gpurun_out/configs_$TAG.err
