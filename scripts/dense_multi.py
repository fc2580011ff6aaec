#!/usr/bin/env python
"""Config 4 at N GPUs (run under torchrun): dense bf16 [21,015,324 x 768] row-sharded, 4096 queries, k=100.
Every rank scores its shard (tcgen05 GEMM + fused top-k), ONE all-gather of rank keys, merge.  Prints one JSON line
(time = max over ranks of CUDA-event time around ShardedIndex.search, queries resident on every GPU)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import vsearch_b200 as vs  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N, D, B, K = 21_015_324, 768, 4096, 100
lo, hi = vs.row_partition(N, world, rank)
g = torch.Generator(device=dev).manual_seed(7 + rank)
x = torch.empty((hi - lo, D), dtype=torch.bfloat16, device=dev)
for a in range(0, hi - lo, 1 << 20):
    b = min(hi - lo, a + (1 << 20))
    x[a:b] = torch.randn((b - a, D), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
q = torch.randn((B, D), generator=torch.Generator(device=dev).manual_seed(99), device=dev, dtype=torch.float32).to(torch.bfloat16)
idx = vs.Index(fp16=False)
idx.vector = x
idx.move_to_device(dev)
del x
sh = vs.ShardedIndex(idx, lo, N)
for _ in range(2):
    res = sh.search(q, K)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 3
e0.record()
for _ in range(steps):
    res = sh.search(q, K)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    t = float(ms)
    print(json.dumps({"config": "cfg4", "n_gpus": world, "rows_per_gpu": hi - lo, "B": B, "k": K, "ms_per_call": t,
                      "qps": B / t * 1e3, "TFLOPs_aggregate": 2.0 * B * N * D / (t * 1e-3) / 1e12}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
