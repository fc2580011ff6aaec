#!/usr/bin/env python
"""Multi-GPU parity of the row-sharded path over NCCL (run under torchrun, one rank per GPU):
every rank builds its contiguous shard of the same seeded index, ShardedIndex.search does local fused top-k ->
ONE all-gather of rank keys -> merge, and rank 0 checks ids/scores bit-exactly against the oracle on the full index.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/nccl_parity.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import vsearch_b200 as vs  # noqa: E402
from oracle import ref_search  # noqa: E402
from tests.util import V, sparse_queries, stratified_csr  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for binary, n, m, k, mode in [(True, 300_001, 60, 100, "scan"), (False, 120_000, 128, 1000, "scan"), (True, 300_001, 60, 100, "inverted")]:
    crow, col, val = stratified_csr(n, V, m, seed=5, grid=True, binary=binary, jitter=11)
    lo, hi = vs.row_partition(n, world, rank)
    c = crow[lo:hi + 1] - crow[lo]
    sl = slice(int(crow[lo]), int(crow[hi]))
    idx = (vs.BoTIndex if binary else vs.SparseIndex)()
    idx.vector = ref_search.torch_csr(c, col[sl], val[sl], (hi - lo, V))
    idx.move_to_device(dev)
    idx.search_mode = mode
    sh = vs.ShardedIndex(idx, lo, n)
    q = sparse_queries(9, V, 200, seed=3)
    if binary:
        q = (q != 0).float()  # heavy ties across shard boundaries
    res = sh.search(q, k)
    torch.cuda.synchronize()
    if rank == 0:
        X = ref_search.torch_csr(crow, col, val, (n, V))
        msg = ref_search.compare_results(res, ref_search.ref_scores(q, X), k, exact=True)
        print(f"world={world} binary={binary} n={n} k={k} mode={mode}:", "OK" if msg is None else msg, flush=True)
        ok = ok and msg is None
# dense index: thresholds pooled between the ranks after every sweep (vs_search_dense_step); grid values, massive ties
for n, d, B, k, dtype, steps in [(1_000_003, 64, 6, 100, torch.bfloat16, True), (1_000_003, 64, 6, 100, torch.bfloat16, False),
                                 (300_000, 128, 260, 10, torch.float16, True)]:
    g = torch.Generator().manual_seed(21)
    x = torch.randint(-8, 9, (n, d), generator=g).float() / 4.0
    q = torch.randint(-8, 9, (B, d), generator=g).float() / 4.0
    lo, hi = vs.row_partition(n, world, rank)
    idx = vs.Index()
    idx.vector = x[lo:hi].to(dtype)
    idx.move_to_device(dev)
    sh = vs.ShardedIndex(idx, lo, n)
    sh.dense_steps = steps
    assert sh._dense_steps_apply(idx._require_engine(), q, k) == (steps and world > 1)
    res = sh.search(q, k)
    torch.cuda.synchronize()
    if rank == 0:
        canon = ref_search.canonical_topk(ref_search.quantize_like(ref_search.ref_scores(q, x), dtype), k)
        good = torch.equal(res.ids.cpu(), canon.ids) and torch.equal(res.scores.float().cpu(), canon.scores)
        print(f"world={world} dense n={n} d={d} B={B} k={k} {dtype} stepwise={steps}:", "OK" if good else "MISMATCH", flush=True)
        ok = ok and good
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
