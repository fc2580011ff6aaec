#!/bin/bash
# round 2, final kernels: compute-sanitizer memcheck over the K3 / dense paths touched last, the ncu launch list of the
# bench command, one ncu --set full capture of the inverted-list kernel on the full config-2 index
set -x
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_search.py -x -q -m gpu \
  -k "inverted or auto_mode or sparse_queries_match or replay or fixed_point" 2>&1 | tail -8 | tee gpurun_out/r2n_sanitize_search.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_dense.py -x -q -m gpu \
  -k "stepwise or fallback or heavy_ties or fp32 or golden" 2>&1 | tail -8 | tee gpurun_out/r2n_sanitize_dense.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2n_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2n_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:inv_search -s 2 -c 1 -o gpurun_out/r2n_k3_full \
  python scripts/exp_scan.py --rows 21015324 --batch 256 --mode inverted --reps 1 > gpurun_out/r2n_k3_ncu.log 2>&1
ls -la gpurun_out/r2n*
