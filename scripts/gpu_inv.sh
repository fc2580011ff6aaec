#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-i}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --mode inverted --qnnz 64 > gpurun_out/bench_${TAG}_inv64.json 2> gpurun_out/bench_${TAG}_inv64.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}_inv64.json')); print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value'],1), d['config'].get('mode_used'), d['config'].get('index_build_s'), d['clocks'])"; tail -3 gpurun_out/bench_${TAG}_inv64.err
timeout 900 python scripts/sweep_crossover.py cfg2_768 > gpurun_out/sweep_${TAG}_cfg2_768.jsonl 2> gpurun_out/sweep_${TAG}_cfg2_768.err; echo "sweep rc=$?"; tail -3 gpurun_out/sweep_${TAG}_cfg2_768.err; cut -c1-220 gpurun_out/sweep_${TAG}_cfg2_768.jsonl
