#!/bin/bash
# round 2: compute-sanitizer memcheck over the small parity tests of every kernel family (new scan, K3, staged merge, loader, fp32 dense)
set -x
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_search.py -x -q -m gpu \
  -k "golden_csr or sparse_queries_match or raw_abi or host_pointers or limits or loader_direct or sharded or pathological or inverted_edge or replay or auto_mode or dense_to_csr or score_rows" 2>&1 | tail -15 | tee gpurun_out/r2i_sanitize_search.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_dense.py -x -q -m gpu -k "golden or heavy_ties or fp32" 2>&1 | tail -8 | tee gpurun_out/r2i_sanitize_dense.log
