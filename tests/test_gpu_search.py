"""Parity of the CUDA path (through the C ABI, via the Python mirror of the reference classes) against
the oracle: golden vectors from the reference, seeded synthetic indices, and the reference's edge cases.
Bit-exact ids/scores on dyadic-grid data; 1e-5 relative (fp32) on continuous data."""
import numpy as np
import pytest
import torch

from oracle import ref_search
from tests.util import V, golden_search_cases, load_golden, sparse_queries, stratified_csr

pytestmark = pytest.mark.gpu


def _mk(cls_name, crow, col, val, shape, device="cuda:0", dtype=None):
    import vsearch_b200 as vs

    cls = getattr(vs, cls_name)
    idx = cls()
    v = torch.as_tensor(val)
    if dtype is not None:
        v = v.to(dtype)
    idx.vector = ref_search.torch_csr(torch.as_tensor(crow).to(torch.int64), torch.as_tensor(col).to(torch.int64), v, shape)
    idx.move_to_device(device)
    return idx


@pytest.mark.parametrize("name", golden_search_cases("csr"))
def test_golden_csr(name, cuda_device):
    z = load_golden(name)
    cls = "BoTIndex" if bool(z["binary"]) else "SparseIndex"
    idx = _mk(cls, z["crow"], z["col"], z["val"], z["shape"])
    q = torch.from_numpy(z["q"])
    ref = torch.from_numpy(z["ref_scores"])
    # scoring alone (index.py:91)
    sc = idx._require_engine().scores(q.reshape(-1, q.shape[-1])).cpu().reshape(ref.shape)
    exact = "cont" not in name
    if exact:
        assert torch.equal(sc, ref + 0.0)
    else:
        torch.testing.assert_close(sc, ref, rtol=1e-5, atol=1e-6)
    # scoring + selection (index.py:91-93)
    res = idx.search(q, z["k"])
    assert res.ids.dtype == torch.int64 and res.ids.is_cuda and tuple(res.ids.shape) == tuple(z["ref_topk_ids"].shape)
    msg = ref_search.compare_results(res, ref, z["k"], exact=exact)
    assert msg is None, msg
    if exact:  # the reference's own top-k VALUES (tie order only affects ids)
        assert torch.equal(res.scores.cpu(), torch.from_numpy(z["ref_topk_scores"]) + 0.0)


@pytest.mark.parametrize("n,m,B,k,binary", [
    (200_000, 120, 8, 100, True),     # config-2 shape, one shard slice
    (100_000, 256, 8, 100, False),    # config-1 shape
    (60_000, 256, 3, 1000, False),    # config-3 k
    (50_000, 120, 5, 1000, True),
])
def test_synthetic_grid_exact(n, m, B, k, binary, cuda_device):
    crow, col, val = stratified_csr(n, V, m, seed=11, grid=True, binary=binary, jitter=17)
    X = ref_search.torch_csr(crow, col, val, (n, V))
    idx = _mk("BoTIndex" if binary else "SparseIndex", crow, col, val, (n, V))
    q = sparse_queries(B, V, 64, seed=5)
    res = idx.search(q, k)
    msg = ref_search.compare_results(res, ref_search.ref_scores(q, X), k, exact=True)
    assert msg is None, msg


def test_synthetic_continuous_tolerance(cuda_device):
    n, m = 80_000, 256
    crow, col, val = stratified_csr(n, V, m, seed=3, grid=False)
    X = ref_search.torch_csr(crow, col, val, (n, V))
    idx = _mk("SparseIndex", crow, col, val, (n, V))
    q = sparse_queries(6, V, 128, seed=9, grid=False)
    res = idx.search(q, 100)
    msg = ref_search.compare_results(res, ref_search.ref_scores(q, X), 100, rtol=1e-5, exact=False)
    assert msg is None, msg


def test_dense_query_768_nnz(cuda_device):
    """the reference's default a=768 activations (retriever.py:134)"""
    n, m = 50_000, 120
    crow, col, val = stratified_csr(n, V, m, seed=21, grid=True, binary=True)
    X = ref_search.torch_csr(crow, col, val, (n, V))
    idx = _mk("BoTIndex", crow, col, val, (n, V))
    q = sparse_queries(4, V, 768, seed=2)
    res = idx.search(q, 100)
    assert ref_search.compare_results(res, ref_search.ref_scores(q, X), 100, exact=True) is None


def test_edge_cases(cuda_device):
    n, v = 300, 1000
    crow, col, val = stratified_csr(n, v, 20, seed=4, grid=True, jitter=20)  # includes empty rows
    X = ref_search.torch_csr(crow, col, val, (n, v))
    idx = _mk("SparseIndex", crow, col, val, (n, v))
    q = sparse_queries(3, v, 30, seed=6, neg=True)
    q[1] = 0  # all-zero query -> ids 0..k-1
    ref = ref_search.ref_scores(q, X)
    for k in (1, 5, 100, n):
        res = idx.search(q, k)
        msg = ref_search.compare_results(res, ref, k, exact=True)
        assert msg is None, f"k={k}: {msg}"
    assert idx.search(q, 7).ids[1].tolist() == list(range(7))
    with pytest.raises(RuntimeError):  # the reference raises RuntimeError for k > N (index.py:92)
        idx.search(q, n + 1)
    # 1-D query -> [k]
    r1 = idx.search(q[0], 9)
    assert tuple(r1.ids.shape) == (9,) and torch.equal(r1.ids, idx.search(q[:1], 9).ids[0])
    # numpy / half / bf16 / device queries
    rh = idx.search(q.to(torch.float16).cuda(), 5)
    assert ref_search.compare_results(rh, ref_search.ref_scores(q.half().float(), X), 5, exact=True) is None


def test_negative_values_untouched_rows_outrank(cuda_device):
    n, v = 5000, 2000
    crow, col, val = stratified_csr(n, v, 16, seed=8, grid=True)
    val = -val
    X = ref_search.torch_csr(crow, col, val, (n, v))
    idx = _mk("SparseIndex", crow, col, val, (n, v))
    q = sparse_queries(4, v, 200, seed=1)
    res = idx.search(q, 50)
    assert ref_search.compare_results(res, ref_search.ref_scores(q, X), 50, exact=True) is None
    assert float(res.scores.max()) <= 0.0


def test_heavy_ties_binary(cuda_device):
    """binary index x constant-valued queries: scores are small integers, thousands of exact ties that
    cross warp, CTA and staging-buffer boundaries; ids must still be (score desc, id asc)."""
    n = 120_000
    crow, col, val = stratified_csr(n, V, 60, seed=13, binary=True, jitter=30)
    X = ref_search.torch_csr(crow, col, val, (n, V))
    idx = _mk("BoTIndex", crow, col, val, (n, V))
    q = (sparse_queries(6, V, 512, seed=3) != 0).float()
    for k in (10, 100, 1000):
        res = idx.search(q, k)
        msg = ref_search.compare_results(res, ref_search.ref_scores(q, X), k, exact=True)
        assert msg is None, f"k={k}: {msg}"


def test_ascending_scores_worst_case_for_threshold(cuda_device):
    """scores increase with the row id: every row beats the running threshold -> exercises the in-kernel
    prune path continuously."""
    n, v = 40_000, 64
    crow = torch.arange(n + 1, dtype=torch.int64)
    col = torch.zeros(n, dtype=torch.int64)
    val = torch.arange(n, dtype=torch.float32) / 8.0
    X = ref_search.torch_csr(crow, col, val, (n, v))
    idx = _mk("SparseIndex", crow, col, val, (n, v))
    q = torch.zeros(2, v)
    q[0, 0], q[1, 0] = 1.0, -1.0
    for k in (100, 1500):
        res = idx.search(q, k)
        assert ref_search.compare_results(res, ref_search.ref_scores(q, X), k, exact=True) is None


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_half_precision_values(dtype, cuda_device):
    """Parity rule 3: quantise values and queries to the storage dtype, feed those numbers to the fp32
    reference; our scores are then rounded to the index dtype (index.py:89 dtype contract)."""
    n, m = 30_000, 64
    crow, col, val = stratified_csr(n, V, m, seed=17, grid=False)
    idx = _mk("SparseIndex", crow, col, val, (n, V), dtype=dtype)
    q = sparse_queries(4, V, 64, seed=2, grid=False)
    vq, qq = ref_search.quantize_like(val, dtype), ref_search.quantize_like(q, dtype)
    ref = ref_search.ref_scores(qq, ref_search.torch_csr(crow, col, vq, (n, V)))
    # (a) fp32 accumulate over the quantised numbers: within 1e-5 of the fp32 reference on the same numbers
    from vsearch_b200 import _native as nat
    ids, sc = idx._require_engine().search(qq, 20, mode="scan", score_round=nat.VS_F32)
    msg = ref_search.compare_results(ref_search.SearchResults(ids, sc), ref, 20, rtol=1e-5, exact=False)
    assert msg is None, msg
    # (b) the public call returns scores in the index dtype (index.py:89): equal to the reference's fp32
    # scores rounded to that dtype, up to one unit in the last place (2^-8 bf16 / 2^-11 fp16 relative)
    res = idx.search(q, 20)
    assert res.scores.dtype == dtype
    canon = ref_search.canonical_topk(ref, 20)
    ulp = 2.0 ** -7 if dtype == torch.bfloat16 else 2.0 ** -10
    torch.testing.assert_close(res.scores.float().cpu(), canon.scores.to(dtype).float(), rtol=ulp, atol=1e-6)


def test_export_roundtrip_and_save(tmp_path, cuda_device):
    import scipy.sparse as sp

    import vsearch_b200 as vs

    n, v = 2000, 3000
    crow, col, val = stratified_csr(n, v, 24, seed=5, grid=True, jitter=24)
    idx = _mk("SparseIndex", crow, col, val, (n, v))
    c2, j2, v2 = idx._require_engine().export_csr()
    assert torch.equal(c2.cpu(), crow) and torch.equal(j2.cpu(), col) and torch.equal(v2.cpu(), val)
    p = str(tmp_path / "idx.npz")
    idx.save(p)
    m = sp.load_npz(p)  # scipy must be able to read what we write
    assert m.shape == (n, v) and np.array_equal(m.indptr, crow.numpy()) and np.array_equal(m.data, val.numpy())
    r = vs.Retriever(device="cuda:0")
    r.load_index(p, index_type="sparse")
    q = sparse_queries(2, v, 40, seed=1)
    a, b = idx.search(q, 10), r.retrieve(q, k=10)
    assert torch.equal(a.ids, b.ids)


def test_virtual_shards_merge(cuda_device):
    """Row-split on one GPU + the same merge kernel = the multi-GPU path without a cluster."""
    import vsearch_b200 as vs

    n, W, k = 90_001, 4, 100
    crow, col, val = stratified_csr(n, V, 48, seed=19, binary=True, jitter=10)
    X = ref_search.torch_csr(crow, col, val, (n, V))
    q = (sparse_queries(5, V, 300, seed=4) != 0).float()  # heavy ties across shard boundaries
    keys = []
    for r in range(W):
        lo, hi = vs.row_partition(n, W, r)
        c = crow[lo:hi + 1] - crow[lo]
        sl = slice(int(crow[lo]), int(crow[hi]))
        shard = _mk("BoTIndex", c, col[sl], val[sl], (hi - lo, V))
        keys.append(shard.search_keys(q, k, id_offset=lo))
    ids, scores = vs.merge_keys(torch.stack(keys), k)
    msg = ref_search.compare_results(ref_search.SearchResults(ids, scores), ref_search.ref_scores(q, X), k, exact=True)
    assert msg is None, msg


def test_large_batch_chunks(cuda_device):
    """B > the internal 1024-query chunk."""
    n, v = 3000, 512
    crow, col, val = stratified_csr(n, v, 16, seed=2, grid=True)
    X = ref_search.torch_csr(crow, col, val, (n, v))
    idx = _mk("SparseIndex", crow, col, val, (n, v))
    q = sparse_queries(1100, v, 12, seed=3)
    res = idx.search(q, 10)
    assert ref_search.compare_results(res, ref_search.ref_scores(q, X), 10, exact=True) is None


# ---------------------------------------------------------------------------------------------------
# K3: token-major inverted lists (mode="inverted") and the auto crossover -- same oracle, same bars
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_search_cases("csr"))
def test_golden_csr_inverted(name, cuda_device):
    z = load_golden(name)
    idx = _mk("BoTIndex" if bool(z["binary"]) else "SparseIndex", z["crow"], z["col"], z["val"], z["shape"])
    idx.search_mode = "inverted"
    res = idx.search(torch.from_numpy(z["q"]), z["k"])
    msg = ref_search.compare_results(res, torch.from_numpy(z["ref_scores"]), z["k"], exact="cont" not in name)
    assert msg is None, msg


@pytest.mark.parametrize("n,m,B,k,binary,qnnz", [
    (200_000, 120, 19, 100, True, 64),
    (100_000, 256, 9, 100, False, 64),
    (60_000, 256, 3, 1000, False, 768),
    (50_000, 120, 5, 1000, True, 768),
])
def test_inverted_synthetic_grid_exact(n, m, B, k, binary, qnnz, cuda_device):
    crow, col, val = stratified_csr(n, V, m, seed=11, grid=True, binary=binary, jitter=17)
    X = ref_search.torch_csr(crow, col, val, (n, V))
    idx = _mk("BoTIndex" if binary else "SparseIndex", crow, col, val, (n, V))
    idx.search_mode = "inverted"
    q = sparse_queries(B, V, qnnz, seed=5, neg=not binary)
    res = idx.search(q, k)
    msg = ref_search.compare_results(res, ref_search.ref_scores(q, X), k, exact=True)
    assert msg is None, msg
    # and both kernel families agree bit for bit
    idx.search_mode = "scan"
    res2 = idx.search(q, k)
    assert torch.equal(res.ids, res2.ids) and torch.equal(res.scores, res2.scores)


def test_inverted_edge_cases(cuda_device):
    n, v = 300, 1000
    crow, col, val = stratified_csr(n, v, 20, seed=4, grid=True, jitter=20)
    X = ref_search.torch_csr(crow, col, val, (n, v))
    idx = _mk("SparseIndex", crow, col, val, (n, v))
    idx.search_mode = "inverted"
    q = sparse_queries(3, v, 30, seed=6, neg=True)
    q[1] = 0  # all-zero query: no postings at all, every score 0 -> ids 0..k-1
    ref = ref_search.ref_scores(q, X)
    for k in (1, 5, 100, n):
        msg = ref_search.compare_results(idx.search(q, k), ref, k, exact=True)
        assert msg is None, f"k={k}: {msg}"
    with pytest.raises(RuntimeError):
        idx.search(q, n + 1)
    r1 = idx.search(q[2], 9)
    assert tuple(r1.ids.shape) == (9,)
    # fewer than k matching rows: zeros fill the tail, lowest ids first (rows never touched still compete)
    q2 = torch.zeros(1, v)
    q2[0, int(col[0])] = 1.0
    res = idx.search(q2, 50)
    assert ref_search.compare_results(res, ref_search.ref_scores(q2, X), 50, exact=True) is None


def test_inverted_heavy_ties_and_continuous(cuda_device):
    n = 120_000
    crow, col, val = stratified_csr(n, V, 60, seed=13, binary=True, jitter=30)
    X = ref_search.torch_csr(crow, col, val, (n, V))
    idx = _mk("BoTIndex", crow, col, val, (n, V))
    idx.search_mode = "inverted"
    q = (sparse_queries(6, V, 512, seed=3) != 0).float()
    for k in (10, 1000):
        msg = ref_search.compare_results(idx.search(q, k), ref_search.ref_scores(q, X), k, exact=True)
        assert msg is None, f"k={k}: {msg}"
    crow, col, val = stratified_csr(50_000, V, 128, seed=3, grid=False)
    X = ref_search.torch_csr(crow, col, val, (50_000, V))
    idx = _mk("SparseIndex", crow, col, val, (50_000, V))
    idx.search_mode = "inverted"
    q = sparse_queries(5, V, 128, seed=9, grid=False)
    msg = ref_search.compare_results(idx.search(q, 100), ref_search.ref_scores(q, X), 100, rtol=1e-5, exact=False)
    assert msg is None, msg


def test_inverted_ascending_scores_replay_path(cuda_device):
    """K3 worst case: scores increase with the row id, so after the sampling phase EVERY row beats the threshold;
    the optimistic whole-block pass overflows the append region and each block is replayed stepwise with joins.
    One token owns the whole list (n postings per block > kLongList): the cooperative long-list path."""
    n, v = 300_000, 64
    crow = torch.arange(n + 1, dtype=torch.int64)
    col = torch.zeros(n, dtype=torch.int64)
    val = (torch.arange(n, dtype=torch.float32) % 65536) / 8.0 + (torch.arange(n) // 65536).float() * 8192.0
    X = ref_search.torch_csr(crow, col, val, (n, v))
    idx = _mk("SparseIndex", crow, col, val, (n, v))
    idx.search_mode = "inverted"
    q = torch.zeros(2, v)
    q[0, 0], q[1, 0] = 1.0, -1.0
    for k in (100, 1500):
        res = idx.search(q, k)
        assert idx.last_mode() == "inverted"
        assert ref_search.compare_results(res, ref_search.ref_scores(q, X), k, exact=True) is None


def test_inverted_multi_block_ctas_and_zipf_lists(cuda_device):
    """> 148 x 36,864 rows: several row blocks per CTA; heavy-tailed token popularity (one token in every row)."""
    n, m = 5_600_000, 6
    g = torch.Generator(device="cuda").manual_seed(5)
    col = torch.randint(1, 4000, (n, m), generator=g, device="cuda", dtype=torch.int64)
    col[:, 0] = 0                                   # token 0 is in every row
    col[:, 1] = torch.randint(1, 9, (n,), generator=g, device="cuda")   # 8 very popular tokens
    col[:, 2:] += 4000 * torch.arange(1, m - 1, device="cuda")[None, :]  # distinct ranges: no duplicates in a row
    col, _ = torch.sort(col, dim=1)
    val = torch.randint(1, 64, (n, m), generator=g, device="cuda").float() / 16.0
    crow = torch.arange(n + 1, device="cuda", dtype=torch.int64) * m
    idx = _mk("SparseIndex", crow, col.reshape(-1), val.reshape(-1), (n, V))
    q = torch.zeros(3, V)
    q[0, 0] = 1.0
    q[1, [0, 3, 5, 4100, 9000]] = torch.tensor([0.5, 2.0, 1.0, 3.0, 1.5])
    q[2, [2, 8, 5000, 13000, 17000, 20001]] = torch.tensor([1.0, -1.0, 2.0, 0.25, 4.0, 1.0])
    idx.search_mode = "inverted"
    a = idx.search(q, 100)
    assert idx.last_mode() == "inverted"
    idx.search_mode = "scan"
    b = idx.search(q, 100)
    assert torch.equal(a.ids, b.ids) and torch.equal(a.scores, b.scores)   # grid values: both paths are exact
    Xs = ref_search.torch_csr(crow[:200_001].cpu(), col[:200_000].reshape(-1).cpu(), val[:200_000].reshape(-1).cpu(), (200_000, V))
    sub = _mk("SparseIndex", crow[:200_001], col[:200_000].reshape(-1), val[:200_000].reshape(-1), (200_000, V))
    sub.search_mode = "inverted"
    assert ref_search.compare_results(sub.search(q, 100), ref_search.ref_scores(q, Xs), 100, exact=True) is None


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_inverted_half_precision_values(dtype, cuda_device):
    n, m = 30_000, 64
    crow, col, val = stratified_csr(n, V, m, seed=17, grid=True)
    idx = _mk("SparseIndex", crow, col, val, (n, V), dtype=dtype)  # grid values are exact in fp16 and bf16
    q = sparse_queries(4, V, 64, seed=2, grid=True)
    idx.search_mode = "inverted"
    a = idx.search(q, 20)
    idx.search_mode = "scan"
    b = idx.search(q, 20)
    assert torch.equal(a.ids, b.ids) and torch.equal(a.scores, b.scores)


def test_auto_mode_crossover(cuda_device):
    """auto picks inverted lists for sparse queries and the scan for dense ones; results do not depend on it."""
    n, m = 60_000, 120
    crow, col, val = stratified_csr(n, V, m, seed=23, grid=True, binary=True)
    X = ref_search.torch_csr(crow, col, val, (n, V))
    idx = _mk("BoTIndex", crow, col, val, (n, V))
    idx.search_mode = "auto"
    for qnnz in (8, 64, 768, 6000):
        q = sparse_queries(3, V, qnnz, seed=qnnz)
        msg = ref_search.compare_results(idx.search(q, 50), ref_search.ref_scores(q, X), 50, exact=True)
        assert msg is None, f"qnnz={qnnz}: {msg}"
    idx.search_mode = "inverted"
    with pytest.raises(NotImplementedError):  # > 4096 non-zeros per query: inverted lists refuse, auto falls back
        idx.search(sparse_queries(1, V, 6000, seed=1), 5)


@pytest.mark.gpu
@pytest.mark.parametrize("max_token", [None, 20])
def test_bot_index_from_token_ids(max_token, cuda_device):
    """GPU bag-of-token construction (vs_bot_from_tokens) against the restated reference builder, then a search on it."""
    import vsearch_b200 as vs

    g = torch.Generator().manual_seed(11)
    n, max_len, vocab, shift = 3000, 128, 30522, 999
    lens = torch.randint(0, max_len + 1, (n,), generator=g, dtype=torch.int32)
    lens[0], lens[1] = 0, max_len
    ids = torch.randint(0, vocab, (n, max_len), generator=g, dtype=torch.int64)
    ids[:, :60] = torch.randint(900, 1400, (n, 60), generator=g)      # many duplicates and ids around the shift
    ids[:, 0], ids[2, :] = 101, 102                                    # [CLS]; a row of nothing but [SEP]
    rows = [ids[i, :int(lens[i])].tolist() for i in range(n)]
    crow, col, shape = ref_search.ref_bot_rows(rows, vocab, shift, max_token)
    idx = vs.BoTIndex.from_token_ids(ids, lens, vocab_size=vocab, num_shift=shift, max_token=max_token, dtype=torch.float32)
    assert tuple(idx.vector.shape) == shape
    got = idx.vector.cpu()
    assert torch.equal(got.crow_indices().to(torch.int64), crow)
    assert torch.equal(got.col_indices().to(torch.int64), col)
    X = ref_search.torch_csr(crow, col, torch.ones(col.numel()), shape)
    q = sparse_queries(4, shape[1], 64, seed=5)
    assert ref_search.compare_results(idx.search(q, 10), ref_search.ref_scores(q, X), 10, exact=True) is None
    ids32 = vs.BoTIndex.from_token_ids(ids.to(torch.int32), None, vocab_size=vocab, num_shift=shift, max_token=max_token)
    full, _, _ = ref_search.ref_bot_rows(ids.tolist(), vocab, shift, max_token)
    assert torch.equal(ids32.vector.cpu().crow_indices().to(torch.int64), full)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16])
def test_pathological_rows_for_the_bank_placement(dtype, cuda_device):
    """Rows the build-time bank placement cannot balance, mixed with ordinary ones: a 20,000-entry row (spans ~40
    steps), rows whose columns all fall in ONE shared-memory bank (multiples of 32), duplicate columns inside a row,
    empty rows, a single-entry row -- through the scan and the inverted lists, fp32 (one chunk per lane) and fp16
    (two chunks per lane) value layouts, against the reference."""
    g = torch.Generator().manual_seed(21)
    rows = []
    for i in range(400):
        kind = i % 8
        if i == 7:
            c = torch.randperm(V, generator=g)[:20_000].sort().values          # very long row
        elif kind == 1:
            c = (torch.randperm(V // 32, generator=g)[:150] * 32).sort().values  # one bank
        elif kind == 2:
            c = torch.randint(0, 40, (60,), generator=g).sort().values          # duplicates (CSR sums them)
        elif kind == 3:
            c = torch.zeros(0, dtype=torch.int64)                               # empty
        elif kind == 4:
            c = torch.randint(0, V, (1,), generator=g)
        else:
            c = torch.randperm(V, generator=g)[:int(torch.randint(1, 300, (1,), generator=g))].sort().values
        rows.append(c)
    lens = torch.tensor([r.numel() for r in rows])
    crow = torch.zeros(len(rows) + 1, dtype=torch.int64)
    crow[1:] = torch.cumsum(lens, 0)
    col = torch.cat(rows)
    val = torch.randint(1, 64, (col.numel(),), generator=g).float() / 16.0     # exact in fp16
    X = ref_search.torch_csr(crow, col, val, (len(rows), V))
    idx = _mk("SparseIndex", crow, col, val, (len(rows), V), dtype=dtype)
    q = sparse_queries(5, V, 400, seed=12)
    q[1, ::32] = 0.5                                                            # hits the one-bank rows hard
    q[2, :40] = 1.0
    ref = ref_search.ref_scores(q, X)
    for mode in ("scan", "inverted"):
        idx.search_mode = mode
        res = idx.search(q, 50)
        assert idx.last_mode() == mode
        msg = ref_search.compare_results(ref_search.SearchResults(res.ids, res.scores.float()),
                                         ref_search.quantize_like(ref, dtype), 50, exact=True)
        assert msg is None, f"{mode}: {msg}"


@pytest.mark.gpu
def test_topk_sparsify_matches_reference(cuda_device):
    """vs_sparsify_topk against the restated upstream sparsifier (utils/sparse.py:8-19, vdr.py:159-169): continuous
    activations, heavy ties at the k-th value, k = 0 / k >= V, the lexical OR, then a search with the sparse queries."""
    import vsearch_b200 as vs

    g = torch.Generator().manual_seed(3)
    emb = torch.rand(6, V, generator=g) * 3
    emb[1] = torch.randint(0, 4, (V,), generator=g).float()          # thousands of ties at the threshold
    emb[2, 100:] = 0                                                  # fewer non-zeros than k
    emb[3] = -emb[3]                                                  # negative activations
    bow = torch.randint(0, V + 999, (6, 40), generator=g, dtype=torch.int32)
    for k in (768, 1, 0, V, V + 5):
        got = vs.topk_sparsify(emb.cuda(), k).cpu()
        assert torch.equal(got, ref_search.ref_topk_sparsify(emb, k)), k
        assert int((got[0] != 0).sum()) == min(k, V)
        got = vs.topk_sparsify(emb.cuda(), k, bow_ids=bow, shift=999).cpu()
        assert torch.equal(got, ref_search.ref_topk_sparsify(emb, k, bow, 999)), k
    assert vs.topk_sparsify(emb[0].cuda(), 10).shape == (V,)
    crow, col, val = stratified_csr(20_000, V, 40, seed=6, grid=True, binary=True)
    idx = _mk("BoTIndex", crow, col, val, (20_000, V))
    q = vs.topk_sparsify(emb.cuda(), 768)
    idx.search_mode = "inverted"                                     # 768 non-zeros per query: inverted lists apply
    res = idx.search(q, 20)
    assert idx.last_mode() == "inverted"
    X = ref_search.torch_csr(crow, col, val, (20_000, V))
    assert ref_search.compare_results(res, ref_search.ref_scores(q.cpu(), X), 20, rtol=1e-5, exact=False) is None


@pytest.mark.gpu
def test_malformed_csr_is_rejected_not_read_out_of_bounds(cuda_device):
    """vs_index_create_csr validates what torch.sparse_csr_tensor validates: row pointers start at 0, never decrease and
    stay within nnz; columns lie in [0, n_cols).  Errors come back as ValueError, nothing is read or written out of range."""
    from vsearch_b200.index import _Engine

    dev = torch.device("cuda:0")
    col = torch.tensor([1, 5, 2, 7], dtype=torch.int64, device=dev)
    val = torch.ones(4, device=dev)
    good = torch.tensor([0, 2, 2, 4], dtype=torch.int64, device=dev)
    _Engine.from_csr(good, col, val, (3, 10), dev)
    for bad_crow in ([0, 3, 2, 4], [1, 2, 2, 4], [0, 2, 2, 4000], [0, -2, 2, 4]):
        with pytest.raises(ValueError):
            _Engine.from_csr(torch.tensor(bad_crow, dtype=torch.int64, device=dev), col, val, (3, 10), dev)
    for bad_col in ([1, 5, 2, 10], [1, -1, 2, 7]):
        with pytest.raises(ValueError):
            _Engine.from_csr(good, torch.tensor(bad_col, dtype=torch.int64, device=dev), val, (3, 10), dev)
