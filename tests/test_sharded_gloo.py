"""world_size-2 test of the row-sharded path's host logic on CPU (gloo): partition arithmetic, the one
all-gather of rank keys, padding of short shards, merge.  The CUDA kernels are replaced by the oracle
through the injection points ShardedIndex offers for exactly this purpose (the product default is the
CUDA merge; without CUDA it raises)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_search
from tests.util import sparse_queries, stratified_csr


class _OracleShard:
    """Stands in for a device-resident shard: same search_keys contract, computed by the oracle."""

    def __init__(self, crow, col, val, shape):
        self.X = ref_search.torch_csr(crow, col, val, shape)
        self.n_rows = shape[0]

    def _require_engine(self):
        return self

    def _value_dtype(self):
        return torch.float32

    def search_keys(self, q, k, id_offset=0):
        res = ref_search.canonical_topk(ref_search.ref_scores(q, self.X), k)
        return ref_search.pack_keys(res.ids + id_offset, res.scores)


def _worker(rank, world, port, n, v, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vsearch_b200 import ShardedIndex, row_partition

    crow, col, val = stratified_csr(n, v, 12, seed=7, grid=True, binary=True, jitter=6)
    lo, hi = row_partition(n, world, rank)
    c = crow[lo:hi + 1] - crow[lo]
    sl = slice(int(crow[lo]), int(crow[hi]))
    shard = _OracleShard(c, col[sl], val[sl], (hi - lo, v))
    sh = ShardedIndex(shard, lo, n, merge_fn=ref_search.merge_keys_oracle)
    q = (sparse_queries(3, v, 40, seed=3) != 0).float()
    res = sh.search(q, k)
    r1 = sh.search(q[0], k)
    assert tuple(r1.ids.shape) == (k,)
    if rank == 0:
        torch.save((res.ids, res.scores), out)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(n, k, tmp_path):
    v, world = 400, 2
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(world, _free_port(), n, v, k, out), nprocs=world, join=True)
    ids, scores = torch.load(out)
    crow, col, val = stratified_csr(n, v, 12, seed=7, grid=True, binary=True, jitter=6)
    X = ref_search.torch_csr(crow, col, val, (n, v))
    q = (sparse_queries(3, v, 40, seed=3) != 0).float()
    msg = ref_search.compare_results(ref_search.SearchResults(ids, scores), ref_search.ref_scores(q, X), k, exact=True)
    assert msg is None, msg


def test_two_ranks_equal_single_index(tmp_path):
    _run(1001, 25, tmp_path)


def test_k_larger_than_one_shard(tmp_path):
    _run(41, 30, tmp_path)  # shards of 21 and 20 rows, k = 30: short shards are padded with empty keys


class _OracleDenseShard:
    """Stands in for a dense shard with the stepwise contract of vs_search_dense_step: step s answers with the top-k of
    the first (s + 1) thirds of its rows, no weaker than what the pooled keys of the previous step allow.  `overflow`
    makes this rank report a survivor-list overflow at the end (every rank must then take the one-all-gather path)."""

    kind = 0
    device = torch.device("cpu")

    def __init__(self, x, overflow=False):
        self.x, self.n_rows, self.overflow, self.calls = x, x.shape[0], overflow, []

    def _require_engine(self):
        return self

    def _value_dtype(self):
        return torch.bfloat16

    def _score_round(self):
        return 0

    def _prep_q(self, q):
        return q.float()

    def search_dense_step(self, step, q, k, n_ranks, gathered, keys_out, status, score_round=0, id_offset=0):
        self.calls.append(step)
        assert (gathered is None) == (step == 0)
        if gathered is not None:
            assert tuple(gathered.shape) == (n_ranks, q.shape[0], k)
        seen = self.x[: max(k, (self.n_rows * (step + 1) + 2) // 3)]
        res = ref_search.canonical_topk(ref_search.ref_scores(q, seen), k)
        keys_out.copy_(ref_search.pack_keys(res.ids + id_offset, res.scores))
        if step == 2 and self.overflow:
            status.fill_(1)

    def search_keys(self, q, k, id_offset=0):
        self.calls.append("keys")
        res = ref_search.canonical_topk(ref_search.ref_scores(q.float(), self.x), k)
        return ref_search.pack_keys(res.ids + id_offset, res.scores)


def _dense_worker(rank, world, port, n, d, k, overflow_rank, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vsearch_b200 import ShardedIndex, row_partition

    g = torch.Generator().manual_seed(3)
    x = torch.randint(-4, 5, (n, d), generator=g).float()
    q = torch.randint(-4, 5, (5, d), generator=g).float()
    lo, hi = row_partition(n, world, rank)
    shard = _OracleDenseShard(x[lo:hi], overflow=(rank == overflow_rank))
    sh = ShardedIndex(shard, lo, n, merge_fn=ref_search.merge_keys_oracle)
    res = sh.search(q, k)
    want = [0, 1, 2] + (["keys"] if overflow_rank >= 0 else [])
    assert shard.calls == want, shard.calls
    if rank == 0:
        torch.save((res.ids, res.scores.float()), out)
    dist.barrier()
    dist.destroy_process_group()


def _run_dense(overflow_rank, tmp_path):
    n, d, k, world = 9000, 16, 20, 2
    out = str(tmp_path / "dense.pt")
    mp.spawn(_dense_worker, args=(world, _free_port(), n, d, k, overflow_rank, out), nprocs=world, join=True)
    ids, scores = torch.load(out)
    g = torch.Generator().manual_seed(3)
    x = torch.randint(-4, 5, (n, d), generator=g).float()
    q = torch.randint(-4, 5, (5, d), generator=g).float()
    msg = ref_search.compare_results(ref_search.SearchResults(ids, scores), ref_search.ref_scores(q, x), k, exact=True)
    assert msg is None, msg


def test_dense_stepwise_three_gathers(tmp_path):
    _run_dense(-1, tmp_path)


def test_dense_stepwise_overflow_on_one_rank_sends_every_rank_to_the_fallback(tmp_path):
    _run_dense(1, tmp_path)
