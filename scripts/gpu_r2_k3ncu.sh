#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:inv_search -s 2 -c 1 -o gpurun_out/r2l_k3_shard python scripts/exp_scan.py --rows 2626916 --batch 256 --mode inverted --reps 1 > gpurun_out/r2l_k3_ncu.log 2>&1
ls -la gpurun_out/r2l_k3_shard.ncu-rep
