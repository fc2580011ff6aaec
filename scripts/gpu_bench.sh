#!/bin/bash
# bench + ncu launch list + one full ncu capture of the scan kernel (1 GPU)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu list rc=$?"
grep -E "scan_topk|merge_topk|prep_query" gpurun_out/launches_$TAG.csv | awk -F'","' '{print $5, $NF}' | head -20
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 1 -c 1 -f -o gpurun_out/prof_scan_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out/
