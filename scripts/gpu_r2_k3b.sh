#!/bin/bash
# round 2: new K3 kernel (balanced accumulate, histogram thresholds): parity + breakdown
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 300 python scripts/exp_scan.py --rows 21015324 --batch 512 --mode inverted --prof 2>&1 | tail -1 | tee gpurun_out/r2f_k3_full.json
timeout 300 python scripts/exp_scan.py --rows 2626916 --batch 512 --mode inverted --prof 2>&1 | tail -1 | tee gpurun_out/r2f_k3_shard.json
timeout 300 python scripts/exp_scan.py --rows 21015324 --batch 256 --mode inverted --qnnz 768 --prof 2>&1 | tail -1 | tee gpurun_out/r2f_k3_full_768.json
