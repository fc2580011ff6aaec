#!/bin/bash
# round 2: native direct-to-device loader: parity tests, then the timing at 8 shards x 2,626,916 rows x 120 tokens
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 600 python scripts/bench_loader.py --shards 3 --rows 200000 --dir /tmp/vs_loader_small 2>&1 | grep '^{' | tee gpurun_out/r2h_loader_small.jsonl
timeout 1500 python scripts/bench_loader.py --shards 8 --rows 2626916 --dir /tmp/vs_loader 2>&1 | grep '^{' | tee gpurun_out/r2h_loader_full.jsonl
