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


def test_inverted_binary_index_fixed_point_and_fp32_queries(cuda_device):
    """K3 on a binary index accumulates in 32-bit fixed point when every query weight is positive and keeps >= 17 bits
    under the scale of the weights' total (native integer shared-memory adds), in fp32 otherwise: grid weights must come
    out bit-exact, continuous ones within the 1e-5 contract, and queries that cannot go to fixed point (a negative
    weight, a 2^20 dynamic range) must take the fp32 path with the same guarantees -- all in one batch."""
    n = 150_000
    crow, col, val = stratified_csr(n, V, 60, seed=23, binary=True, jitter=30)
    X = ref_search.torch_csr(crow, col, val, (n, V))
    idx = _mk("BoTIndex", crow, col, val, (n, V))
    idx.search_mode = "inverted"
    q = sparse_queries(4, V, 64, seed=5)                         # multiples of 1/64: exact in either arithmetic
    for k in (10, 300):
        msg = ref_search.compare_results(idx.search(q, k), ref_search.ref_scores(q, X), k, exact=True)
        assert msg is None, f"grid weights, k={k}: {msg}"
    q = sparse_queries(6, V, 64, seed=6, grid=False)             # U(0.01, 3): fixed point
    q[1] = sparse_queries(1, V, 64, seed=7, grid=False, neg=True)[0]      # mixed signs: fp32
    nz = q[2].nonzero().flatten()
    q[2, nz[0]] = 3.0e-6                                         # 2^20 below the largest weight: fp32
    q[3] = sparse_queries(1, V, 500, seed=8, grid=False)[0]      # a long query, still one token tile
    assert idx.last_mode() == "inverted"
    for k in (10, 300):
        res = idx.search(q, k)
        msg = ref_search.compare_results(res, ref_search.ref_scores(q, X), k, rtol=1e-5)
        assert msg is None, f"continuous weights, k={k}: {msg}"
    # the same batch through the scan: same ids wherever the reference scores are not near-ties
    idx.search_mode = "scan"
    msg = ref_search.compare_results(idx.search(q, 300), ref_search.ref_scores(q, X), 300, rtol=1e-5)
    assert msg is None, msg


def test_inverted_binary_index_many_blocks_per_cta_and_popular_tokens(cuda_device, monkeypatch):
    """K3 with several row blocks per CTA on a binary index (block size forced down to 2,048 rows: 196 blocks on 148
    SMs): blocks after a CTA's first one catch the rows that cross the pre-filter while the postings are added instead
    of scanning the sums, and three tokens that sit in (almost) every row have lists longer than 1,024 postings per
    block -- the vectorised all-threads path.  Grid weights: bit-exact against the oracle, for fixed-point queries and
    for one with a negative weight (fp32 path, scanned blocks)."""
    monkeypatch.setenv("VSEARCH_B200_K3_BLOCK_ROWS", "2048")
    n, m = 400_000, 24
    crow, col, val = stratified_csr(n, V, m, seed=51, binary=True, jitter=5)
    # popular tokens: column 7 in every row, 11 in two rows of three, 13 in every row but the first 1,000 (odd list starts)
    rows = torch.arange(n)
    extra_r = torch.cat([rows, rows[rows % 3 != 0], rows[1000:]])
    extra_c = torch.cat([torch.full((n,), 7), torch.full((int((rows % 3 != 0).sum()),), 11), torch.full((n - 1000,), 13)])
    crow_t = torch.as_tensor(crow).to(torch.int64)
    r_all = torch.cat([torch.repeat_interleave(rows, crow_t[1:] - crow_t[:-1]), extra_r])
    c_all = torch.cat([torch.as_tensor(col).to(torch.int64), extra_c])
    X = torch.sparse_coo_tensor(torch.stack([r_all, c_all]), torch.ones(r_all.numel()), (n, V)).coalesce()
    X = (X.to_sparse_csr())
    crow2, col2 = X.crow_indices(), X.col_indices()
    val2 = torch.ones(col2.numel())
    Xb = ref_search.torch_csr(crow2, col2, val2, (n, V))
    idx = _mk("BoTIndex", crow2, col2, val2, (n, V))
    idx.search_mode = "inverted"
    q = sparse_queries(5, V, 48, seed=17)
    q[:, 7] = torch.tensor([1.0, 0.5, 2.0, 0.25, 1.5])
    q[:, 11] = torch.tensor([0.75, 1.0, 0.0, 3.0, 0.5])
    q[:, 13] = torch.tensor([2.0, 0.0, 1.0, 1.0, 0.25])
    q[4, 11] = -0.5                                   # one query off the fixed-point path
    ref = ref_search.ref_scores(q, Xb)
    for k in (10, 200):
        msg = ref_search.compare_results(idx.search(q, k), ref, k, exact=True)
        assert msg is None, f"k={k}: {msg}"
    assert idx.last_mode() == "inverted"
    idx.search_mode = "scan"
    assert ref_search.compare_results(idx.search(q, 200), ref, 200, exact=True) is None


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
    idx.search_mode = "inverted"   # > 4096 non-zeros per query: the lists cannot serve it, the scan answers (and says so)
    q = sparse_queries(1, V, 6000, seed=1)
    assert ref_search.compare_results(idx.search(q, 5), ref_search.ref_scores(q, X), 5, exact=True) is None
    assert idx.last_mode() == "scan"


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


# ---------------------------------------------------------------------------------------------------------------
# round 2: sparse-query entry point (vs_search_sparse), no host wait, limits, larger oracle-backed cases
def _sparse_forms(q):
    """the same queries as a torch CSR tensor, a COO tensor, and host-resident CSR parts"""
    csr = q.to_sparse_csr()
    return {"csr": csr, "coo": q.to_sparse_coo(), "csr_cuda": csr.cuda()}


@pytest.mark.gpu
@pytest.mark.parametrize("binary", [True, False])
def test_sparse_queries_match_dense_queries_and_oracle(binary, cuda_device):
    """Index.search with (token, weight) lists (torch sparse tensors -> vs_search_sparse) through the scan, the inverted
    lists and auto: bit-exact against the oracle on grid data, and equal to the dense-query path."""
    n, m = 90_000, 96
    crow, col, val = stratified_csr(n, V, m, seed=31, grid=True, binary=binary, jitter=9)
    X = ref_search.torch_csr(crow, col, val, (n, V))
    idx = _mk("BoTIndex" if binary else "SparseIndex", crow, col, val, (n, V))
    q = sparse_queries(7, V, 64, seed=8, neg=not binary)
    q[3] = 0                                  # an empty query: all scores 0, ids 0..k-1
    ref = ref_search.ref_scores(q, X)
    for mode in ("scan", "inverted", "auto"):
        idx.search_mode = mode
        dense = idx.search(q, 50)
        for name, qs in _sparse_forms(q).items():
            res = idx.search(qs, 50)
            msg = ref_search.compare_results(res, ref, 50, exact=True)
            assert msg is None, f"{mode}/{name}: {msg}"
            assert torch.equal(res.ids, dense.ids) and torch.equal(res.scores, dense.scores), f"{mode}/{name}"
    assert idx.search(q[3].to_sparse(), 4).ids.tolist() == [0, 1, 2, 3]   # 1-D sparse query


@pytest.mark.gpu
def test_sparse_query_lists_raw_abi_host_and_device(cuda_device):
    """vs_search_sparse with raw lists: host pointers (staged through the workspace), int32 / int64 offsets, tokens
    outside [0, V) ignored, zero weights ignored, duplicate tokens of a query added up."""
    n = 40_000
    crow, col, val = stratified_csr(n, V, 64, seed=33, grid=True, binary=False)
    X = ref_search.torch_csr(crow, col, val, (n, V))
    idx = _mk("SparseIndex", crow, col, val, (n, V))
    eng = idx._require_engine()
    g = torch.Generator().manual_seed(4)
    toks, ws, ptr = [], [], [0]
    qd = torch.zeros(5, V)
    for b in range(5):
        t = torch.randint(0, V, (40,), generator=g).to(torch.int32)
        w = torch.randint(1, 65, (40,), generator=g).float() / 32.0
        t[3] = t[5]                       # a duplicate token: weights add
        w[7] = 0.0                        # ignored
        t = torch.cat([t, torch.tensor([-5, V, V + 100], dtype=torch.int32)])   # ignored
        w = torch.cat([w, torch.tensor([1.0, 2.0, 3.0])])
        for tt, ww in zip(t.tolist(), w.tolist()):
            if 0 <= tt < V:
                qd[b, tt] += ww
        toks.append(t); ws.append(w); ptr.append(ptr[-1] + t.numel())
    tok, w = torch.cat(toks), torch.cat(ws)
    ref = ref_search.ref_scores(qd, X)
    for mode in ("scan", "inverted"):
        for ptr_dtype in (torch.int64, torch.int32):
            for dev in ("cpu", "cuda:0"):
                p_ = torch.tensor(ptr, dtype=ptr_dtype, device=dev)
                ids, sc = eng.search_sparse(p_, tok.to(dev), w.to(dev), 30, mode=mode)
                torch.cuda.synchronize()
                msg = ref_search.compare_results(ref_search.SearchResults(ids, sc), ref, 30, exact=True)
                assert msg is None, f"{mode}/{ptr_dtype}/{dev}: {msg}"


@pytest.mark.gpu
def test_binary_scan_sparse_query_lists_of_many_lengths(cuda_device):
    """The binary scan on an index large enough that every warp's part has several 256-chunk steps, fed (token, weight)
    lists of 0, 1, 64, 300, 640 and 641 entries (the CTA has 640 threads) with duplicates and out-of-range tokens,
    int32 and int64 offsets, host and device lists; bit-exact against the oracle (grid weights) and equal to the
    dense-row path.  (A variant of the kernel that copied the next query's list to shared memory with cp.async during
    the pass was measured 0.9 % slower on a 1/8 shard and dropped: staging is bound by the first step's loads.)"""
    n, m = 400_000, 64
    crow, col, val = stratified_csr(n, V, m, seed=41, grid=True, binary=True, jitter=9)
    X = ref_search.torch_csr(crow, col, val, (n, V))
    idx = _mk("BoTIndex", crow, col, val, (n, V))
    eng = idx._require_engine()
    g = torch.Generator().manual_seed(14)
    lens = [64, 0, 640, 1, 641, 64, 300, 64]
    toks, ws, ptr = [], [], [0]
    qd = torch.zeros(len(lens), V)
    for b, ln in enumerate(lens):
        t = torch.randint(0, V, (ln,), generator=g).to(torch.int32)
        w = torch.randint(1, 65, (ln,), generator=g).float() / 32.0
        if ln >= 8:
            t[3] = t[5]                   # a duplicate token: weights add
            t[6] = V + 7                  # ignored
        for tt, ww in zip(t.tolist(), w.tolist()):
            if 0 <= tt < V:
                qd[b, tt] += ww
        toks.append(t); ws.append(w); ptr.append(ptr[-1] + ln)
    tok, w = torch.cat(toks), torch.cat(ws)
    ref = ref_search.ref_scores(qd, X)
    idx.search_mode = "scan"
    dense = idx.search(qd, 20)
    assert ref_search.compare_results(dense, ref, 20, exact=True) is None
    for ptr_dtype in (torch.int64, torch.int32):
        for dev in ("cuda:0", "cpu"):
            p_ = torch.tensor(ptr, dtype=ptr_dtype, device=dev)
            ids, sc = eng.search_sparse(p_, tok.to(dev), w.to(dev), 20, mode="scan")
            torch.cuda.synchronize()
            msg = ref_search.compare_results(ref_search.SearchResults(ids, sc), ref, 20, exact=True)
            assert msg is None, f"{ptr_dtype}/{dev}: {msg}"
            assert torch.equal(ids, dense.ids)


@pytest.mark.gpu
def test_dense_queries_from_host_pointers_with_a_leading_dimension(cuda_device):
    """vs_search straight from host memory (ctypes, no torch copy): staged through the workspace, ldq > n_cols."""
    import ctypes

    from vsearch_b200 import _native as nat

    n = 30_000
    crow, col, val = stratified_csr(n, V, 48, seed=35, grid=True, binary=True)
    X = ref_search.torch_csr(crow, col, val, (n, V))
    idx = _mk("BoTIndex", crow, col, val, (n, V))
    eng = idx._require_engine()
    q = sparse_queries(6, V, 64, seed=2)
    ld = V + 13
    qh = torch.zeros(6, ld)
    qh[:, :V] = q
    ids = torch.empty((6, 20), dtype=torch.int64, device="cuda:0")
    sc = torch.empty((6, 20), dtype=torch.float32, device="cuda:0")
    for mode in ("scan", "auto"):
        ws = eng.workspace(6, 20)
        rc = nat.LIB.vs_search(eng.handle, ctypes.c_void_p(qh.data_ptr()), nat.VS_F32, 6, ld, 20, nat.MODES[mode], nat.VS_F32, 0,
                               ids.data_ptr(), sc.data_ptr(), ws.data_ptr(), ws.numel(), None)
        nat.check(rc)
        torch.cuda.synchronize()
        msg = ref_search.compare_results(ref_search.SearchResults(ids, sc), ref_search.ref_scores(q, X), 20, exact=True)
        assert msg is None, f"{mode}: {msg}"


@pytest.mark.gpu
def test_search_does_not_wait_for_the_stream(cuda_device):
    """The no-synchronisation contract of the header: with device queries, search (scan, inverted, auto) returns while
    the stream is still busy with earlier work -- enqueue ~0.5 s of spinning first and time the calls."""
    import time

    n = 50_000
    crow, col, val = stratified_csr(n, V, 64, seed=37, grid=True, binary=True)
    idx = _mk("BoTIndex", crow, col, val, (n, V))
    q = sparse_queries(4, V, 64, seed=2).cuda()
    qs = q.to_sparse_csr()
    for mode in ("scan", "inverted", "auto"):          # warm: inverted lists built, workspace allocated
        idx.search_mode = mode
        idx.search(q, 10); idx.search(qs, 10)
    torch.cuda.synchronize()
    spin = int(0.5 * 1.9e9)
    for mode in ("scan", "inverted", "auto"):
        idx.search_mode = mode
        torch.cuda._sleep(spin)
        t0 = time.perf_counter()
        r1 = idx.search(q, 10)
        r2 = idx.search(qs, 10)
        dt = time.perf_counter() - t0
        busy = not torch.cuda.current_stream().query()
        torch.cuda.synchronize()
        assert busy and dt < 0.1, f"{mode}: search blocked the host for {dt * 1e3:.1f} ms (stream busy: {busy})"
        assert torch.equal(r1.ids, r2.ids)


@pytest.mark.gpu
def test_limits_k_2048_and_32767_columns(cuda_device):
    """The two documented hard limits at their edge: k = VS_MAX_K = 2048 and n_cols = 32,767 (15 column bits + the
    row-end flag), scan and inverted lists, against the oracle."""
    g = torch.Generator().manual_seed(5)
    n, vmax = 30_000, 32767
    crow, col, val = stratified_csr(n, vmax, 40, seed=41, grid=True, binary=False, jitter=5)
    col[-1] = vmax - 1                       # the last column is in use
    X = ref_search.torch_csr(crow, col, val, (n, vmax))
    idx = _mk("SparseIndex", crow, col, val, (n, vmax))
    q = torch.zeros(3, vmax)
    for b in range(3):
        q[b, torch.randperm(vmax, generator=g)[:300]] = torch.randint(1, 64, (300,), generator=g).float() / 16.0
    q[0, vmax - 1] = 2.0
    ref = ref_search.ref_scores(q, X)
    for mode in ("scan", "inverted"):
        idx.search_mode = mode
        for k in (2048, 7):
            msg = ref_search.compare_results(idx.search(q, k), ref, k, exact=True)
            assert msg is None, f"{mode} k={k}: {msg}"
    with pytest.raises(NotImplementedError):
        idx.search(q, 2049)
    crow, col, val = stratified_csr(9000, V, 120, seed=42, grid=True, binary=True)   # binary scan at k = 2048
    idx = _mk("BoTIndex", crow, col, val, (9000, V))
    idx.search_mode = "scan"
    qb = sparse_queries(2, V, 200, seed=6)
    msg = ref_search.compare_results(idx.search(qb, 2048), ref_search.ref_scores(qb, ref_search.torch_csr(crow, col, val, (9000, V))),
                                     2048, exact=True)
    assert msg is None, msg


@pytest.mark.gpu
def test_inverted_and_scan_continuous_2m_rows(cuda_device):
    """Oracle-backed continuous-data case at 2M rows x 256 nnz/row (512M postings): shared-memory float atomics make the
    inverted lists' summation order vary from run to run, the scan's is fixed; both within 1e-5 of the reference with
    the near-tie-aware id comparison."""
    n, m = 2_000_000, 256
    g = torch.Generator().manual_seed(77)
    w = V // m
    base = (torch.arange(m, dtype=torch.int32) * V) // m
    col = (torch.randint(0, w, (n, m), generator=g, dtype=torch.int32) + base[None, :]).reshape(-1)
    val = torch.rand(n * m, generator=g) * 1.99 + 0.01
    crow = torch.arange(n + 1, dtype=torch.int64) * m
    q = sparse_queries(4, V, 256, seed=19, grid=False)
    X = ref_search.torch_csr(crow, col.to(torch.int64), val, (n, V))
    ref = ref_search.ref_scores(q, X)
    del X
    import vsearch_b200 as vs
    from vsearch_b200.index import _Engine

    idx = vs.SparseIndex()
    idx._engine = _Engine.from_csr(crow.cuda(), col.cuda(), val.cuda(), (n, V), torch.device("cuda:0"))
    idx.device = "cuda:0"
    for mode in ("inverted", "scan"):
        idx.search_mode = mode
        for k in (100, 1000):
            res = idx.search(q, k)
            assert idx.last_mode() == mode
            msg = ref_search.compare_results(res, ref, k, rtol=1e-5, exact=False)
            assert msg is None, f"{mode} k={k}: {msg}"


@pytest.mark.gpu
def test_score_rows_large_batch(cuda_device):
    """vs_score_rows prepares the whole batch at once: B = 2,500 needs more than one search chunk's workspace."""
    n = 20_000
    crow, col, val = stratified_csr(n, V, 32, seed=43, grid=True, binary=False)
    X = ref_search.torch_csr(crow, col, val, (n, V))
    idx = _mk("SparseIndex", crow, col, val, (n, V))
    B = 2500
    q = sparse_queries(B, V, 16, seed=3)
    g = torch.Generator().manual_seed(9)
    ids = torch.randint(0, n, (B, 3), generator=g)
    got = idx.score_rows(q, ids).cpu()
    ref = torch.gather(ref_search.ref_scores(q, X), 1, ids)
    assert torch.equal(got, ref + 0.0)


@pytest.mark.gpu
def test_dense_to_csr_and_sparse_sparsifier_output(cuda_device):
    """vs_dense_to_csr against torch's to_sparse_csr (fp32 / fp16 / bf16, empty rows, a strided view), the sparsifier's
    as_sparse output, and build_index from a dense [N, V] matrix (sparsified on the GPU, .vector exports CSR)."""
    import vsearch_b200 as vs

    g = torch.Generator().manual_seed(11)
    x = torch.rand(37, 1000, generator=g)
    x[x < 0.9] = 0
    x[5] = 0
    for dt in (torch.float32, torch.float16, torch.bfloat16):
        xd = x.to(dt).cuda()
        crow, col, val = vs.dense_to_csr(xd)
        ref = x.to(dt).float().to_sparse_csr()
        assert torch.equal(crow.cpu(), ref.crow_indices()) and torch.equal(col.cpu().long(), ref.col_indices())
        assert torch.equal(val.cpu(), ref.values())
    wide = torch.zeros(37, 1200)
    wide[:, :1000] = x
    crow, col, val = vs.dense_to_csr(wide.cuda()[:, :1000])                 # row stride 1200
    assert torch.equal(col.cpu().long(), x.to_sparse_csr().col_indices())
    emb = torch.rand(5, V, generator=g)
    sp = vs.topk_sparsify(emb.cuda(), 100, as_sparse=True)
    de = vs.topk_sparsify(emb.cuda(), 100)
    assert sp.layout == torch.sparse_csr and torch.equal(sp.to_dense(), de)
    # build_index(vectors=dense [N, V]) -> GPU sparsification, same results as the CSR route
    dense = torch.zeros(300, V)
    crow, col, val = stratified_csr(300, V, 30, seed=44, grid=True, binary=False)
    Xc = ref_search.torch_csr(crow, col, val, (300, V))
    dense = Xc.to_dense()
    r = vs.Retriever(device="cuda:0")
    r.build_index(vectors=dense, index_type="sparse")
    assert r.index.vector.layout == torch.sparse_csr and tuple(r.index.vector.shape) == (300, V)
    q = sparse_queries(3, V, 64, seed=5)
    assert ref_search.compare_results(r.retrieve(q, k=10), ref_search.ref_scores(q, Xc), 10, exact=True) is None
    r.build_index(vectors=dense, index_type="bag_of_token")
    Xb = ref_search.torch_csr(crow, col, torch.ones(col.numel()), (300, V))
    res = r.retrieve(q, k=10)
    assert res.scores.dtype == torch.float16
    msg = ref_search.compare_results(ref_search.SearchResults(res.ids, res.scores.float()),
                                     ref_search.quantize_like(ref_search.ref_scores(q, Xb), torch.float16), 10, exact=True)
    assert msg is None, msg


@pytest.mark.gpu
def test_sharded_rank_without_rows(cuda_device):
    """row_partition(10, 8): ranks 5..7 own no rows.  Such a rank contributes empty keys instead of failing on k = 0."""
    import vsearch_b200 as vs
    from vsearch_b200.index import _Engine

    assert vs.row_partition(10, 8, 6) == (10, 10)
    dev = torch.device("cuda:0")
    idx = vs.SparseIndex()
    idx._engine = _Engine.from_csr(torch.zeros(1, dtype=torch.int64, device=dev), torch.zeros(0, dtype=torch.int64, device=dev),
                                   torch.zeros(0, device=dev), (0, V), dev)
    idx.device = "cuda:0"
    res = vs.ShardedIndex(idx, 10, 10).search(sparse_queries(2, V, 8, seed=1), 3)
    assert tuple(res.ids.shape) == (2, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("shift", [0, 100])
def test_loader_direct_to_device_matches_reference_loader(shift, tmp_path, cuda_device):
    """SparseIndex(index_file, device='cuda'): the native loader (vs_index_load_npz: parallel inflate -> pinned staging ->
    device CSR -> index build with the column shift on the GPU) against what the unmodified reference loader produced
    from the same shard files (tests/golden/load_shift*.npz: sorted-glob order index10 < index2, `[:, shift:]`, row
    concatenation; index.py:172-176), then float64 / int64 shard files and the fp16 default."""
    import os

    import scipy.sparse as sp

    import vsearch_b200 as vs
    from tests.util import GOLDEN

    z = np.load(os.path.join(GOLDEN, f"load_shift{shift}.npz"))
    idx = vs.SparseIndex(os.path.join(GOLDEN, "shards", "index*.npz"), None, fp16=False, device="cuda:0", shift=shift)
    assert idx._engine is not None and idx._vector is None          # no host copy was made
    assert sum(idx.shard_rows) == int(z["shape"][0])
    v = idx.vector.cpu()                                             # exported from the engine on demand
    assert tuple(v.shape) == tuple(z["shape"]) and v.values().dtype == torch.float32
    assert np.array_equal(v.crow_indices().numpy(), z["crow"])
    assert np.array_equal(v.col_indices().numpy(), z["col"])
    assert np.array_equal(v.values().numpy(), z["val"])
    X = ref_search.torch_csr(z["crow"], z["col"], z["val"], tuple(z["shape"]))
    q = sparse_queries(3, int(z["shape"][1]), 40, seed=4)
    assert ref_search.compare_results(idx.search(q, 5), ref_search.ref_scores(q, X), 5, rtol=1e-5, exact=False) is None
    half = vs.SparseIndex(os.path.join(GOLDEN, "shards", "index1.npz"), device="cuda:0")    # fp16=True is upstream's default
    assert half.vector.values().dtype == torch.float16 and half._engine.store_dtype == 1
    # a shard as scipy writes it by default: float64 values, int64 indices after a vstack of large parts
    m = sp.load_npz(os.path.join(GOLDEN, "shards", "index1.npz")).astype(np.float64)
    m.indices, m.indptr = m.indices.astype(np.int64), m.indptr.astype(np.int64)
    sp.save_npz(str(tmp_path / "f64.npz"), m)
    wide = vs.SparseIndex(str(tmp_path / "f64.npz"), fp16=False, device="cuda:0")
    assert np.array_equal(wide.vector.cpu().values().numpy(), m.data.astype(np.float32))
    bot = vs.BoTIndex(str(tmp_path / "f64.npz"), fp16=False, device="cuda:0")                # values are not all 1: stays valued
    assert bot._engine.kind == 1
    m.data[:] = 1.0
    sp.save_npz(str(tmp_path / "ones.npz"), m)
    bot = vs.BoTIndex(str(tmp_path / "ones.npz"), device="cuda:0")
    assert bot._engine.kind == 2                                                             # all ones: column ids only


@pytest.mark.gpu
def test_sharded_index_from_shard_files(cuda_device):
    """ShardedIndex.from_shard_files: the shard files are dealt to the ranks in contiguous groups of upstream's sorted-glob
    order, each rank loads its own files onto its GPU and gets its global id offset from the headers of the files before
    it.  Two virtual ranks on one GPU + the merge = the search of the whole index."""
    import os

    import vsearch_b200 as vs
    from tests.util import GOLDEN
    from vsearch_b200.index import merge_keys

    pattern = os.path.join(GOLDEN, "shards", "index*.npz")
    z = np.load(os.path.join(GOLDEN, "load_shift0.npz"))
    X = ref_search.torch_csr(z["crow"], z["col"], z["val"], tuple(z["shape"]))
    q = sparse_queries(3, int(z["shape"][1]), 40, seed=4)
    whole = vs.ShardedIndex.from_shard_files(pattern, "cuda:0", fp16=False)
    assert whole.row_offset == 0 and whole.n_rows_total == int(z["shape"][0])
    assert ref_search.compare_results(whole.search(q, 6), ref_search.ref_scores(q, X), 6, rtol=1e-5, exact=False) is None
    parts = [vs.ShardedIndex.from_shard_files(pattern, "cuda:0", world=2, rank=r, fp16=False) for r in range(2)]
    assert parts[0].row_offset == 0 and parts[1].row_offset == parts[0].local._require_engine().n_rows
    assert parts[1].row_offset + parts[1].local._require_engine().n_rows == int(z["shape"][0])
    keys = torch.stack([p.local.search_keys(q, 6, id_offset=p.row_offset) for p in parts])
    ids, sc = merge_keys(keys, 6)
    msg = ref_search.compare_results(ref_search.SearchResults(ids, sc), ref_search.ref_scores(q, X), 6, rtol=1e-5, exact=False)
    assert msg is None, msg
