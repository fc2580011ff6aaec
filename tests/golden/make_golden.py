"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

It imports /root/reference/src/ir/retriever/index.py by file path (the package
__init__ pulls in encoders whose dependencies are not installed -- SURVEY.md 8c)
and records, for small seeded inputs, the outputs of the reference's own
``SparseIndex.search`` / ``BoTIndex.search`` / ``Index.search`` (index.py:88-94),
``SparseIndex.init_index`` (index.py:163-179: sorted glob, ``[:, shift:]``,
vstack) and ``SparseIndex.save`` (index.py:181-202).

Each ``search_*.npz`` holds the inputs (CSR arrays or dense matrix, queries in
COO form), the reference's full score matrix (index.py:91), the reference's
``topk`` ids/values (index.py:92, arbitrary tie order) and k.
"""
import importlib.util
import os
import sys
import warnings

import numpy as np
import scipy.sparse as sp
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/ir/retriever/index.py"
V = 29523  # 30522 - 999 (src/ir/encoder/vdr.py:37,72)


def load_ref():
    spec = importlib.util.spec_from_file_location("ref_index", REF)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def rand_csr(rng, n, v, mean_nnz, grid, binary=False, neg=False, empty_every=0):
    lens = np.clip(rng.normal(mean_nnz, mean_nnz / 3, size=n).round().astype(np.int64), 0, v)
    if empty_every:
        lens[::empty_every] = 0
    lens[n // 2] = min(v, 4 * mean_nnz + 3)  # one long row
    crow = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=crow[1:])
    col = np.empty(crow[-1], dtype=np.int32)
    for r in range(n):
        c = np.sort(rng.choice(v, size=lens[r], replace=False))
        col[crow[r]:crow[r + 1]] = c
    if n > 2 and lens[1] > 1:  # columns 0 and V-1 present
        col[crow[1]] = 0
        col[crow[2] - 1] = v - 1
    if binary:
        val = np.ones(crow[-1], dtype=np.float32)
    elif grid:
        val = (rng.integers(1, 256, size=crow[-1]) / 64.0).astype(np.float32)
    else:
        val = rng.uniform(0.01, 2.0, size=crow[-1]).astype(np.float32)
    if neg:
        val *= rng.choice([-1.0, 1.0], size=val.shape).astype(np.float32)
    return crow, col, val


def rand_queries(rng, b, v, nnz, grid, neg=False):
    qi = np.stack([np.sort(rng.choice(v, size=nnz, replace=False)) for _ in range(b)]).astype(np.int32)
    if grid:
        qv = (rng.integers(1, 193, size=(b, nnz)) / 64.0).astype(np.float32)
    else:
        qv = rng.uniform(0.01, 3.0, size=(b, nnz)).astype(np.float32)
    if neg:
        qv *= rng.choice([-1.0, 1.0], size=qv.shape).astype(np.float32)
    return qi, qv


def dense_q(qi, qv, v):
    q = np.zeros((qi.shape[0], v), dtype=np.float32)
    np.put_along_axis(q, qi.astype(np.int64), qv, axis=1)
    return q


def run_sparse(m, cls, crow, col, val, shape, q, k):
    idx = cls()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        idx.vector = idx._scipy_csr_to_torch_csr(sp.csr_array((val, col, crow.astype(np.int32)), shape=shape))
        qt = torch.from_numpy(q)
        res = idx.search(qt, k)
        scores = torch.matmul(qt.to(idx.device).type(idx.vector.dtype), idx.vector.t())
    return scores.numpy(), res.ids.numpy(), res.scores.numpy()


def main():
    if not os.path.exists(REF):
        sys.exit("reference not mounted; fixtures can only be regenerated in the build container")
    m = load_ref()
    torch.manual_seed(0)

    cases = [
        # name, cls, n, v, mean_nnz, B, qnnz, k, grid, binary, neg, empty_every
        ("sparse_grid_k10", "SparseIndex", 3000, V, 48, 5, 64, 10, True, False, False, 97),
        ("sparse_grid_k100", "SparseIndex", 3000, V, 48, 5, 64, 100, True, False, False, 97),
        ("sparse_cont_k100", "SparseIndex", 3000, V, 48, 4, 64, 100, False, False, False, 0),
        ("sparse_neg_k25", "SparseIndex", 1200, V, 40, 4, 256, 25, True, False, True, 50),
        ("bot_binary_k100", "BoTIndex", 4000, V, 30, 6, 32, 100, True, True, False, 0),
        ("bot_binary_k1", "BoTIndex", 4000, V, 30, 3, 32, 1, True, True, False, 0),
        ("sparse_small_keqn", "SparseIndex", 40, 500, 12, 3, 20, 40, True, False, False, 7),
    ]
    for seed, (name, cls, n, v, mn, b, qn, k, grid, binary, neg, ee) in enumerate(cases):
        rng = np.random.default_rng(1000 + seed)
        crow, col, val = rand_csr(rng, n, v, mn, grid, binary=binary, neg=neg, empty_every=ee)
        qi, qv = rand_queries(rng, b, v, qn, grid, neg=neg)
        if name == "sparse_neg_k25":
            qv[1] = 0.0  # an all-zero query: every score 0 -> ids 0..k-1
        q = dense_q(qi, qv, v)
        scores, rids, rsc = run_sparse(m, getattr(m, cls), crow, col, val, (n, v), q, k)
        np.savez_compressed(os.path.join(HERE, f"search_{name}.npz"), kind="csr", crow=crow, col=col, val=val,
                            shape=np.array([n, v]), q_idx=qi, q_val=qv, k=k, ref_scores=scores,
                            ref_topk_ids=rids, ref_topk_scores=rsc, binary=binary)
        print(name, scores.shape, "nnz", len(col))

    # 1-D query -> [k] results
    rng = np.random.default_rng(77)
    crow, col, val = rand_csr(rng, 500, 2000, 20, True)
    qi, qv = rand_queries(rng, 1, 2000, 30, True)
    q = dense_q(qi, qv, 2000)[0]
    idx = m.SparseIndex()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        idx.vector = idx._scipy_csr_to_torch_csr(sp.csr_array((val, col, crow.astype(np.int32)), shape=(500, 2000)))
        res = idx.search(torch.from_numpy(q), 7)
        scores = torch.matmul(torch.from_numpy(q), idx.vector.t())
    np.savez_compressed(os.path.join(HERE, "search_sparse_1d_k7.npz"), kind="csr", crow=crow, col=col, val=val,
                        shape=np.array([500, 2000]), q_idx=qi, q_val=qv, k=7, ref_scores=scores.numpy(),
                        ref_topk_ids=res.ids.numpy(), ref_topk_scores=res.scores.numpy(), binary=False, one_d=True)
    print("1d", tuple(res.ids.shape))

    # dense index: idx = Index(); idx.vector = X (the upstream loader is broken)
    rng = np.random.default_rng(5)
    for name, n, d, b, k, grid in [("dense_grid_k20", 1500, 64, 4, 20, True), ("dense_cont_k50", 1500, 64, 4, 50, False)]:
        if grid:
            x = (rng.integers(-32, 33, size=(n, d)) / 16.0).astype(np.float32)
            q = (rng.integers(-32, 33, size=(b, d)) / 16.0).astype(np.float32)
        else:
            x = rng.standard_normal((n, d)).astype(np.float32)
            q = rng.standard_normal((b, d)).astype(np.float32)
        di = m.Index()
        di.vector = torch.from_numpy(x)
        res = di.search(torch.from_numpy(q), k)
        scores = torch.matmul(torch.from_numpy(q), di.vector.t())
        np.savez_compressed(os.path.join(HERE, f"search_{name}.npz"), kind="dense", x=x, q=q, k=k,
                            ref_scores=scores.numpy(), ref_topk_ids=res.ids.numpy(),
                            ref_topk_scores=res.scores.numpy())
        print(name, scores.shape)

    # loader: sorted-glob order (index10 < index2), shift, vstack  (index.py:172-175)
    rng = np.random.default_rng(9)
    shard_dir = os.path.join(HERE, "shards")
    os.makedirs(shard_dir, exist_ok=True)
    vfull = 700
    for i, n in zip([0, 1, 2, 10], [11, 7, 13, 5]):
        crow, col, val = rand_csr(rng, n, vfull, 9, True)
        sp.save_npz(os.path.join(shard_dir, f"index{i}.npz"),
                    sp.csr_array((val, col, crow.astype(np.int32)), shape=(n, vfull)))
    for shift in (0, 100):
        idx = m.SparseIndex(os.path.join(shard_dir, "index*.npz"), None, fp16=False, device="cpu", shift=shift)
        vv = idx.vector
        np.savez_compressed(os.path.join(HERE, f"load_shift{shift}.npz"),
                            crow=vv.crow_indices().numpy(), col=vv.col_indices().numpy(),
                            val=vv.values().numpy(), shape=np.array(vv.shape))
        print("load shift", shift, tuple(vv.shape))

    # saver: a file written by the reference's SparseIndex.save (index.py:181-202), int64 indices
    idx = m.SparseIndex(os.path.join(shard_dir, "index1.npz"), None, fp16=False, device="cpu")
    big = torch.sparse_csr_tensor(idx.vector.crow_indices().to(torch.int64), idx.vector.col_indices().to(torch.int64),
                                  idx.vector.values(), size=idx.vector.shape)
    idx.vector = big
    idx.save(os.path.join(HERE, "saved_by_reference.npz"))
    print("saved_by_reference.npz members:", list(np.load(os.path.join(HERE, "saved_by_reference.npz")).keys()))


if __name__ == "__main__":
    main()
