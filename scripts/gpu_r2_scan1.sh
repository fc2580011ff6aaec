#!/bin/bash
# round 2, first GPU job: parity of the transposed 8-chunk-per-lane layout, then timing + phase breakdown + conflict counters
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_retriever.py -x -q -m gpu 2>&1 | tail -15
timeout 300 python scripts/exp_scan.py --rows 2626916 --batch 256 --prof --check 2>&1 | tail -3 | tee gpurun_out/r2a_scan_shard.json
timeout 400 python scripts/exp_scan.py --rows 21015324 --batch 128 --prof 2>&1 | tail -3 | tee gpurun_out/r2a_scan_full.json
timeout 300 ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts.sum,smsp__inst_executed.sum,dram__bytes_read.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg \
  --clock-control none -k regex:scan_bin --csv --log-file gpurun_out/r2a_scan_counters.csv python scripts/exp_scan.py --rows 2626916 --batch 8 --reps 1 > gpurun_out/r2a_ncu.log 2>&1
tail -5 gpurun_out/r2a_scan_counters.csv
