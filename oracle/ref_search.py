"""CPU oracle for the vsearch index-scoring hot path.  TEST INFRASTRUCTURE ONLY.

This file is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under ``vsearch_b200/`` imports it.

What it restates (all citations relative to the upstream repo jzhoubu/vsearch):

* ``Index.search``                       src/ir/retriever/index.py:88-94
      q = q.to(device).type(vector.dtype)              :89
      scores = torch.matmul(q, vector.t())             :91   (under no_grad)
      scores.topk(k) -> SearchResults(indices, values) :92-93
* ``SparseIndex._scipy_csr_to_torch_csr`` src/ir/retriever/index.py:144-161
* ``Retriever.process_query`` tensor/ndarray branches  src/ir/retriever/retriever.py:96-99
* ``Retriever.retrieve`` -> ``index.search``           src/ir/retriever/retriever.py:133-136

The arithmetic of that path lives in a third-party dependency that is NOT in
the reference tree: PyTorch (upstream pins torch 2.3.0, poetry.lock; this image
has torch 2.11.0).  ``torch.matmul(dense, sparse_csc)`` dispatches to ATen's
sparse-CSR addmm (MKL ``mkl_sparse_s_mm`` on CPU) and ``Tensor.topk`` to ATen's
partial sort.  The oracle therefore calls the *same torch entry points in the
same order* as the reference's own call sites, and pins itself two ways:

1. ``tests/golden/*.npz`` were produced by importing the unmodified reference
   file ``/root/reference/src/ir/retriever/index.py`` in the build container and
   calling ``SparseIndex.search`` / ``Index.search``
   (``tests/golden/make_golden.py``); ``tests/test_oracle.py`` checks this
   oracle -- and the independent plain-C restatement in ``oracle/oracle.c`` --
   against every one of them.
2. The reference has NO tests, golden vectors or fixtures of its own for this
   path, so beyond (1) parity is "pinned to outputs of
   the reference itself run here", not to upstream-published vectors.

Tie rule.  ``torch.topk`` returns an arbitrary order among equal scores
(measured: SURVEY.md 3.4b).  The contract (BASELINE.json north_star) is "ids
bit-exact, ties broken by lower id", so the canonical answer is defined on the
reference's *score matrix*: stable descending sort => (score desc, id asc).
``-0.0 == +0.0`` compare equal, as in torch.
"""
from __future__ import annotations

import warnings
from typing import NamedTuple, Optional

import numpy as np
import torch


class SearchResults(NamedTuple):  # index.py:16-18 (ids first, scores second)
    ids: torch.Tensor
    scores: torch.Tensor


def torch_csr(crow, col, val, shape) -> torch.Tensor:
    """index.py:154-159 -- build the torch sparse CSR tensor the reference searches."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return torch.sparse_csr_tensor(
            torch.as_tensor(crow), torch.as_tensor(col), torch.as_tensor(val), size=tuple(shape)
        )


def ref_scores(q_embs: torch.Tensor, vector: torch.Tensor) -> torch.Tensor:
    """index.py:89-91 -- the full [B, N] (or [N]) score matrix of the reference."""
    q = q_embs.to(vector.device).type(vector.dtype)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return torch.matmul(q, vector.t())


def ref_search(q_embs: torch.Tensor, vector: torch.Tensor, k: int) -> SearchResults:
    """index.py:88-94 verbatim semantics (tie order = whatever torch.topk gives)."""
    scores = ref_scores(q_embs, vector)
    top = scores.topk(k)  # raises RuntimeError when k > N, like the reference
    return SearchResults(top.indices, top.values)


def canonical_topk(scores: torch.Tensor, k: int) -> SearchResults:
    """Canonical top-k of a reference score matrix: score desc, then id asc.

    NaN is not supported by the contract (documented in DESIGN.md)."""
    if k > scores.shape[-1]:
        raise RuntimeError("selected index k out of range")
    s = scores + 0.0  # -0.0 -> +0.0 so both zeros tie
    order = torch.sort(s, dim=-1, descending=True, stable=True)
    ids = order.indices[..., :k].contiguous()
    vals = torch.gather(scores, -1, ids) + 0.0
    return SearchResults(ids, vals)


def oracle_search(q_embs: torch.Tensor, vector: torch.Tensor, k: int) -> SearchResults:
    """Reference scores + canonical tie order: the answer the CUDA path must give."""
    return canonical_topk(ref_scores(q_embs, vector), k)


def process_query(queries) -> torch.Tensor:
    """retriever.py:96-99 -- ndarray -> torch.Tensor (fp32 copy); Tensor passthrough."""
    if isinstance(queries, np.ndarray):
        return torch.Tensor(queries)
    if isinstance(queries, torch.Tensor):
        return queries
    raise NotImplementedError(f"Query type {type(queries)} not supported")


def quantize_like(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Parity rule 3: round to the storage dtype, come back to fp32."""
    return x.to(dtype).to(torch.float32)


# ---------------------------------------------------------------------------
# shard merge: the reference has no multi-GPU search; the oracle
# for it is "search the row-concatenated index", and this helper restates the
# merge the sharded path must be equivalent to.
# ---------------------------------------------------------------------------
def merge_shard_results(ids_per_shard, scores_per_shard, k: int) -> SearchResults:
    """Merge per-shard top-k lists (global ids) into the global canonical top-k."""
    ids = torch.cat(list(ids_per_shard), dim=-1)
    sc = torch.cat(list(scores_per_shard), dim=-1) + 0.0
    # sort by (score desc, id asc): first by id asc (stable), then by score desc (stable)
    o1 = torch.sort(ids, dim=-1, stable=True).indices
    ids1, sc1 = torch.gather(ids, -1, o1), torch.gather(sc, -1, o1)
    o2 = torch.sort(sc1, dim=-1, descending=True, stable=True).indices[..., :k]
    return SearchResults(torch.gather(ids1, -1, o2), torch.gather(sc1, -1, o2))


def compare_results(ours: SearchResults, ref_scores_mat: torch.Tensor, k: int,
                    rtol: float = 1e-5, exact: bool = False) -> Optional[str]:
    """Parity comparator (SURVEY.md 8c rules 1-2).  Returns None when OK, else a message.

    exact=True  : ids must equal the canonical ids, scores must be bit-equal.
    exact=False : scores within rtol of the reference's canonical scores; ids may
                  differ from the canonical ids only where the *reference* scores of
                  the two ids are within rtol of each other (near-tie), and every
                  returned id's reference score must match the returned score.
    """
    canon = canonical_topk(ref_scores_mat, k)
    ids = ours.ids.cpu().to(torch.int64)
    sc = ours.scores.cpu().to(torch.float32)
    if ids.shape != canon.ids.shape:
        return f"shape mismatch {tuple(ids.shape)} vs {tuple(canon.ids.shape)}"
    if exact:
        if not torch.equal(ids, canon.ids):
            bad = (ids != canon.ids).nonzero()[:5].tolist()
            return f"ids differ from canonical at {bad}"
        if not torch.equal(sc, canon.scores.to(torch.float32)):
            return "scores not bit-equal to the reference's"
        return None
    ref = canon.scores.to(torch.float32)
    tol = rtol * ref.abs().clamp_min(1e-30) + 1e-30
    if not bool(((sc - ref).abs() <= tol + rtol * sc.abs()).all()):
        return f"scores outside rtol={rtol}: max abs err {(sc - ref).abs().max().item()}"
    ref2d = ref_scores_mat.reshape(-1, ref_scores_mat.shape[-1]).to(torch.float32)
    ids2d = ids.reshape(-1, k)
    sc2d = sc.reshape(-1, k)
    got_ref = torch.gather(ref2d, 1, ids2d)
    if not bool(((got_ref - sc2d).abs() <= rtol * got_ref.abs() + rtol * sc2d.abs() + 1e-30).all()):
        return "a returned id's reference score does not match the returned score"
    for b in range(ids2d.shape[0]):
        if ids2d[b].unique().numel() != k:
            return f"duplicate ids in row {b}"
    mism = ids2d != canon.ids.reshape(-1, k)
    if mism.any():
        a = torch.gather(ref2d, 1, canon.ids.reshape(-1, k))[mism]
        g = got_ref[mism]
        if not bool(((a - g).abs() <= 4 * rtol * a.abs().clamp_min(1e-30)).all()):
            return "ids differ outside near-tie runs"
    return None


# ---------------------------------------------------------------------------
# rank-key wire format of the sharded path (include/vsearch_b200.h, vs_search_keys):
# key = ordered(score bits) << 32 | ~uint32(global id), carried as int64 bit patterns.
# ---------------------------------------------------------------------------
def pack_keys(ids: torch.Tensor, scores: torch.Tensor) -> torch.Tensor:
    s = (scores.to(torch.float32) + 0.0).contiguous().numpy().view(np.uint32).astype(np.uint64)
    neg = (s & np.uint64(0x80000000)) != 0
    ordered = np.where(neg, (~s) & np.uint64(0xFFFFFFFF), s | np.uint64(0x80000000))
    low = (~ids.numpy().astype(np.uint64)) & np.uint64(0xFFFFFFFF)
    return torch.from_numpy(((ordered << np.uint64(32)) | low).view(np.int64).copy())


def unpack_keys(keys: torch.Tensor):
    u = keys.contiguous().numpy().view(np.uint64)
    hi = (u >> np.uint64(32)).astype(np.uint32)
    bits = np.where(hi & np.uint32(0x80000000), hi & np.uint32(0x7FFFFFFF), ~hi)
    ids = ((~u) & np.uint64(0xFFFFFFFF)).astype(np.int64)
    return torch.from_numpy(ids), torch.from_numpy(bits.astype(np.uint32).view(np.float32).copy())


def merge_keys_oracle(gathered: torch.Tensor, k: int):
    """[P, B, k_in] keys -> (ids, scores) [B, k]: sort keys descending as unsigned, drop empty (0) keys."""
    P, B, kin = gathered.shape
    u = gathered.permute(1, 0, 2).reshape(B, P * kin).contiguous().numpy().view(np.uint64)
    order = np.argsort(u, axis=1, kind="stable")[:, ::-1][:, :k]
    top = np.take_along_axis(u, order, axis=1)
    ids, sc = unpack_keys(torch.from_numpy(top.view(np.int64).copy()))
    return ids, sc


def ref_bot_rows(token_lists, vocab_size=30522, num_shift=999, max_token=None):
    """Restates upstream Retriever._build_bot_vectors (src/ir/retriever/retriever.py:232-251) for ONE batch: dense
    zeros [n, vocab], ``emb[i, token_ids] = 1`` (first ``max_token`` distinct ids in order when given,
    index_utils.py:11-21), the ``[:, num_shift:]`` slice, ``to_sparse_coo`` -> ``to_sparse_csr``.
    Returns (crow int64, col int64, shape)."""
    n = len(token_lists)
    emb = torch.zeros([n, vocab_size], dtype=torch.float32)
    for i, ids in enumerate(token_lists):
        ids = list(ids)
        if max_token:
            seen, first = set(), []
            for e in ids:
                if e in seen:
                    continue
                seen.add(e)
                first.append(e)
                if len(seen) == max_token:
                    break
            ids = first
        if ids:
            emb[i, ids] = 1
    csr = emb[:, num_shift:].to_sparse_coo().to_sparse_csr()
    return csr.crow_indices().to(torch.int64), csr.col_indices().to(torch.int64), (n, vocab_size - num_shift)


def ref_topk_sparsify(emb: torch.Tensor, k: int, bow_ids=None, shift: int = 0) -> torch.Tensor:
    """Restates upstream utils/sparse.py:8-19 (topk -> scatter a bool mask -> multiply) with the logical_or of the
    row's own token columns (encoder/vdr.py:159-169); ties at the k-th value go to the lower column (upstream's
    torch.topk leaves them arbitrary)."""
    emb = emb.to(torch.float32)
    B, V = emb.shape
    mask = torch.zeros_like(emb, dtype=torch.bool)
    if k >= V:
        mask[:] = True
    elif k > 0:
        order = torch.sort(emb + 0.0, dim=-1, descending=True, stable=True).indices[:, :k]   # stable: lower column first
        mask.scatter_(-1, order, True)
    if bow_ids is not None:
        for b in range(B):
            for t in bow_ids[b].tolist():
                if 0 <= t - shift < V:
                    mask[b, t - shift] = True
    return emb * mask
