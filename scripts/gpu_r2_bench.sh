#!/bin/bash
# round 2: full GPU test-suite, bench smoke at reduced sizes, the real 1-GPU bench line, one ncu --set full capture of the scan
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 600 python bench.py --rows 400000 --batch 64 --extras-rows 300000 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -2 > gpurun_out/r2c_bench_smoke.json; tail -c 1500 gpurun_out/r2c_bench_smoke.json
( time timeout 900 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 300 gpurun_out/r2c_bench.err; tail -c 6000 gpurun_out/r2c_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_bin -s 2 -c 1 -o gpurun_out/r2c_scan_full python scripts/exp_scan.py --rows 21015324 --batch 16 --reps 1 > gpurun_out/r2c_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
