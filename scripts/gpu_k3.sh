#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-k3}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log
for mode in inverted auto; do
timeout 900 python bench.py --steps 3 --warmup 2 --mode $mode --no-cpu-baseline > gpurun_out/bench_${TAG}_$mode.json 2> gpurun_out/bench_${TAG}_$mode.err; echo "bench $mode rc=$?"; cat gpurun_out/bench_${TAG}_$mode.json; tail -3 gpurun_out/bench_${TAG}_$mode.err
done
timeout 900 python bench.py --steps 3 --warmup 2 --mode inverted --qnnz 768 --no-cpu-baseline > gpurun_out/bench_${TAG}_inv768.json 2> gpurun_out/bench_${TAG}_inv768.err; echo "bench inv768 rc=$?"; cat gpurun_out/bench_${TAG}_inv768.json; tail -3 gpurun_out/bench_${TAG}_inv768.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --mode inverted --no-cpu-baseline --batch 64 > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu list rc=$?"
grep -E "inv_|merge_topk|prep_query" gpurun_out/launches_$TAG.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -5
grep -E "inv_accum|inv_select|inv_extract|inv_build" gpurun_out/launches_$TAG.csv | awk -F'","' '{print $5, $NF}' | tail -12
