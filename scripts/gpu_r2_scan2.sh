#!/bin/bash
# round 2: histogram-threshold binary scan; warps-per-CTA variants
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_retriever.py -x -q -m gpu 2>&1 | tail -15
for lib in libvsearch_b200 libvs_w20 libvs_w16; do
  export VSEARCH_B200_LIB=$PWD/vsearch_b200/lib/$lib.so
  echo "== $lib shard"; timeout 300 python scripts/exp_scan.py --rows 2626916 --batch 512 --prof 2>&1 | tail -1 | tee gpurun_out/r2b_${lib}_shard.json
  echo "== $lib full"; timeout 400 python scripts/exp_scan.py --rows 21015324 --batch 1024 --reps 2 --prof 2>&1 | tail -1 | tee gpurun_out/r2b_${lib}_full.json
done
unset VSEARCH_B200_LIB
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "cfg2 or scan" 2>&1 | tail -5
