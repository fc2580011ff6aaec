#!/bin/bash
# multi-GPU: NCCL parity + strong-scaling bench at N ranks (N = number of visible GPUs)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); TAG=${1:-m}
echo "GPUs: $N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/nccl_parity.py > gpurun_out/nccl_parity_${TAG}_$N.log 2>&1; echo "nccl parity rc=$?"; grep -E "world=|Error|error" gpurun_out/nccl_parity_${TAG}_$N.log | head
for mode in scan auto; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 3 --warmup 3 --mode $mode > gpurun_out/bench_${TAG}_${N}gpu_$mode.json 2> gpurun_out/bench_${TAG}_${N}gpu_$mode.err; echo "bench $N gpu $mode rc=$?"
tail -1 gpurun_out/bench_${TAG}_${N}gpu_$mode.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'GB/s', d['roofline']['achieved'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['config'].get('mode_used'))"; tail -2 gpurun_out/bench_${TAG}_${N}gpu_$mode.err
done
