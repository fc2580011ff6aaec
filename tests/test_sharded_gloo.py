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
