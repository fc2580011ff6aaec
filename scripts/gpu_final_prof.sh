#!/bin/bash
# ncu evidence for profiles/: launch list of a bench step, full captures of the three scoring kernels; then the secondary configs
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TAG=${1:-r1g}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_topk -s 1 -c 1 -f -o gpurun_out/prof_scan_$TAG \
    python bench.py --steps 1 --warmup 1 --mode scan --no-auto --no-cpu-baseline --batch 16 > gpurun_out/ncu_scan_$TAG.log 2>&1; echo "ncu scan rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:inv_search -s 1 -c 1 -f -o gpurun_out/prof_invsearch_$TAG \
    python bench.py --steps 1 --warmup 1 --mode inverted --no-cpu-baseline --batch 16 > gpurun_out/ncu_inv_$TAG.log 2>&1; echo "ncu inv rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_topk_pair -s 5 -c 1 -f -o gpurun_out/prof_dense_$TAG \
    python scripts/prof_dense.py > gpurun_out/ncu_dense_$TAG.log 2>&1; echo "ncu dense rc=$?"
timeout 900 python scripts/bench_configs.py cfg1 cfg3 > gpurun_out/configs_${TAG}_cfg13.jsonl 2> gpurun_out/configs_${TAG}_cfg13.err; echo "cfg1/3 rc=$?"; cut -c1-260 gpurun_out/configs_${TAG}_cfg13.jsonl
for c in cfg2_768 cfg2_ragged cfg2_86 cfg2_zipf cfg5; do
  timeout 900 python scripts/sweep_crossover.py $c > gpurun_out/sweep_${TAG}_$c.jsonl 2> gpurun_out/sweep_${TAG}_$c.err; echo "$c rc=$?"; tail -2 gpurun_out/sweep_${TAG}_$c.err
done
