#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-k3c}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
for q in 64 768; do
timeout 900 python bench.py --steps 3 --warmup 2 --mode inverted --qnnz $q --no-cpu-baseline > gpurun_out/bench_${TAG}_inv$q.json 2> gpurun_out/bench_${TAG}_inv$q.err; echo "bench inverted qnnz=$q rc=$?"; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_${TAG}_inv$q.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['value'])"; tail -3 gpurun_out/bench_${TAG}_inv$q.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_${TAG}_inv.csv \
    python bench.py --steps 1 --warmup 1 --mode inverted --no-cpu-baseline --batch 32 > gpurun_out/ncu_launches_${TAG}_inv.log 2>&1; echo "ncu list inv rc=$?"
grep -E "inv_accum|inv_select|inv_extract|inv_build|merge_topk|prep_query" gpurun_out/launches_${TAG}_inv.csv | awk -F'","' '{print $5, $NF}' | tail -9
