#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_dense.py tests/test_gpu_search.py -x -q -m gpu 2>&1 | tail -12
