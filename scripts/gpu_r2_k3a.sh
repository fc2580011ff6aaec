#!/bin/bash
# round 2: K3 phase breakdown, shared-atomic ceiling, dense ncu with tensor-pipe counters, box facts for the loader
set -x
mkdir -p gpurun_out
./scripts/micro/smem_atomics | tee gpurun_out/r2e_smem_atomics.json
timeout 300 python scripts/exp_scan.py --rows 21015324 --batch 512 --mode inverted --prof 2>&1 | tail -1 | tee gpurun_out/r2e_k3_full.json
timeout 300 python scripts/exp_scan.py --rows 2626916 --batch 512 --mode inverted --prof 2>&1 | tail -1 | tee gpurun_out/r2e_k3_shard.json
timeout 300 python scripts/exp_scan.py --rows 21015324 --batch 256 --mode inverted --qnnz 768 --prof 2>&1 | tail -1 | tee gpurun_out/r2e_k3_full_768.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_topk_pair -s 2 -c 1 -o gpurun_out/r2e_dense_full python scripts/prof_dense.py > gpurun_out/r2e_dense_ncu.log 2>&1
free -g | head -2; nproc; df -h /tmp . | tail -2
