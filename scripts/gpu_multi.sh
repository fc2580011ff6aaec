#!/bin/bash
# multi-GPU: NCCL parity + strong-scaling bench at N ranks (N = number of visible GPUs), launched the way the driver does
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); TAG=${1:-m}
echo "GPUs: $N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/nccl_parity.py > gpurun_out/nccl_parity_${TAG}_$N.log 2>&1; echo "nccl parity rc=$?"; grep -E "world=|Error|error" gpurun_out/nccl_parity_${TAG}_$N.log | head
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_${N}gpu.json 2> gpurun_out/bench_${TAG}_${N}gpu.err; echo "bench $N gpu rc=$?"
tail -1 gpurun_out/bench_${TAG}_${N}gpu.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'GB/s', d['roofline']['achieved'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['config'].get('mode_used'), 'auto', d.get('auto_mode',{}).get('value'), d.get('auto_mode',{}).get('e2e',{}).get('value'), d['clocks'])"; tail -2 gpurun_out/bench_${TAG}_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29536 scripts/dense_multi.py 2>/dev/null | grep cfg4 | tee gpurun_out/dense_${TAG}_${N}gpu.json
