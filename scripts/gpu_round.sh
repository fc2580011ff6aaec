#!/bin/bash
# what the driver runs at round end, plus cfg4: gpu tests, smoke, reference arm, bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout 1500 python scripts/bench_configs.py cfg4 > gpurun_out/configs_${TAG}_cfg4.jsonl 2> gpurun_out/configs_${TAG}_cfg4.err; echo "cfg4 rc=$?"; cat gpurun_out/configs_${TAG}_cfg4.jsonl | cut -c1-400; tail -3 gpurun_out/configs_${TAG}_cfg4.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_reference.json 2> gpurun_out/bench_${TAG}_reference.err; echo "bench reference rc=$?"; cut -c1-300 gpurun_out/bench_${TAG}_reference.json
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; cat gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
