"""K4: dense index (tcgen05 GEMM + fused top-k) against the oracle (torch fp32 matmul + canonical top-k)."""
import numpy as np
import pytest
import torch

from oracle import ref_search
from tests.util import golden_search_cases, load_golden

pytestmark = pytest.mark.gpu


def _grid(shape, seed, scale=16.0, lim=32):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(-lim, lim + 1, shape, generator=g).float() / scale


def _dense_index(x, dtype=torch.bfloat16):
    import vsearch_b200 as vs

    idx = vs.Index()
    idx.vector = x.to(dtype)
    idx.move_to_device("cuda:0")
    return idx


@pytest.mark.parametrize("name", golden_search_cases("dense"))
def test_golden_dense(name, cuda_device):
    z = load_golden(name)
    x, q, k = torch.from_numpy(z["x"]), torch.from_numpy(z["q"]), z["k"]
    exact = "grid" in name  # grid values are exact in bf16: every accumulation order gives the same fp32 bits
    idx = _dense_index(x)
    res = idx.search(q, k)
    assert res.ids.dtype == torch.int64 and tuple(res.ids.shape) == (q.shape[0], k)
    if exact:  # scores come back in the index dtype (index.py:89): round the reference the same way, then rank
        canon = ref_search.canonical_topk(ref_search.quantize_like(torch.from_numpy(z["ref_scores"]), torch.bfloat16), k)
        assert torch.equal(res.ids.cpu(), canon.ids)
        assert torch.equal(res.scores.float().cpu(), canon.scores)
    else:
        # bf16 storage, continuous data (SURVEY.md 8c rule 3): quantise values and queries like the index does, feed those
        # numbers to the fp32 reference; returned scores are the fp32 accumulations rounded to bf16 once (half an ulp =
        # 2^-9 ~ 2e-3 relative), ids must be the canonical ones except inside runs of reference scores closer than that
        ref = ref_search.ref_scores(ref_search.quantize_like(q, torch.bfloat16), ref_search.quantize_like(x, torch.bfloat16))
        msg = ref_search.compare_results(ref_search.SearchResults(res.ids, res.scores.float()), ref, k, rtol=2e-3, exact=False)
        assert msg is None, msg


@pytest.mark.parametrize("n,d,B,k,dtype", [
    (5000, 768, 7, 10, torch.bfloat16),       # one sample sweep only (N < sample)
    (70_000, 768, 130, 100, torch.bfloat16),  # sample + filtered sweep, 2 query tiles, ragged N
    (40_000, 128, 33, 1000, torch.float16),   # k = 1000, fp16 storage, D = 128
    (300, 96, 3, 300, torch.bfloat16),        # k == N, D not a multiple of 64
    (1_200_000, 64, 5, 100, torch.bfloat16),  # all three sweeps: sample, 64x sample with tau1, the rest with tau2
    (300_000, 64, 700, 100, torch.bfloat16),  # CTA-pair kernel (cta_group::2): 3 pair tiles of 256 queries, last one ragged
    (70_000, 192, 300, 10, torch.float16),    # CTA-pair kernel, odd number of 128-query tiles, 3 k-blocks
])
def test_dense_grid_exact(n, d, B, k, dtype, cuda_device):
    x, q = _grid((n, d), 1), _grid((B, d), 2)
    idx = _dense_index(x, dtype)
    res = idx.search(q, k)
    assert res.scores.dtype == dtype
    ref = ref_search.ref_scores(q, x)
    # scores come back in the index dtype (index.py:89): compare ids exactly and scores after the same rounding
    canon = ref_search.canonical_topk(ref_search.quantize_like(ref, dtype), k)
    assert torch.equal(res.ids.cpu(), canon.ids), (res.ids.cpu() != canon.ids).nonzero()[:5]
    assert torch.equal(res.scores.float().cpu(), canon.scores)


def test_dense_heavy_ties_and_edges(cuda_device):
    n, d = 30_000, 64
    x = (_grid((n, d), 3, scale=1.0, lim=1)).float()   # entries in {-1, 0, 1}: integer scores, massive ties
    q = (_grid((5, d), 4, scale=1.0, lim=1)).float()
    q[1] = 0
    idx = _dense_index(x)
    ref = ref_search.ref_scores(q, x)
    for k in (1, 100):
        res = idx.search(q, k)
        msg = ref_search.compare_results(ref_search.SearchResults(res.ids, res.scores.float()), ref, k, exact=True)
        assert msg is None, f"k={k}: {msg}"
    assert idx.search(q, 5).ids[1].tolist() == [0, 1, 2, 3, 4]
    with pytest.raises(RuntimeError):
        idx.search(q, n + 1)
    assert tuple(idx.search(q[0], 4).ids.shape) == (4,)


def test_dense_sorted_index_overflow_retry(cuda_device):
    """Passages sorted by increasing score: the sample prefix gives a useless threshold, every row of the sweep
    survives it, the candidate lists overflow and the kernel must tighten and repeat."""
    n, d = 1_000_000, 64
    x = torch.zeros(n, d)
    x[:, 0] = torch.arange(n).float() / 4.0
    q = torch.zeros(2, d)
    q[0, 0], q[1, 0] = 1.0, -1.0
    idx = _dense_index(x, torch.float16 if False else torch.bfloat16)
    xq = ref_search.quantize_like(x, torch.bfloat16)
    res = idx.search(q, 10)
    canon = ref_search.canonical_topk(ref_search.quantize_like(ref_search.ref_scores(q, xq), torch.bfloat16), 10)
    assert torch.equal(res.ids.cpu(), canon.ids)


def test_dense_continuous_ids_and_scores_at_2e3(cuda_device):
    """Continuous N(0,1) data at a size where all three sweeps run (sample, tau1 sweep, tau2 sweep): ids through the
    near-tie-aware comparator, scores within 2e-3 of the fp32 reference on the bf16-quantised inputs."""
    g = torch.Generator().manual_seed(12)
    n, d, B, k = 1_300_000, 128, 48, 100
    x = torch.randn(n, d, generator=g)
    q = torch.randn(B, d, generator=g)
    idx = _dense_index(x)
    res = idx.search(q, k)
    ref = ref_search.ref_scores(ref_search.quantize_like(q, torch.bfloat16), ref_search.quantize_like(x, torch.bfloat16))
    msg = ref_search.compare_results(ref_search.SearchResults(res.ids, res.scores.float()), ref, k, rtol=2e-3, exact=False)
    assert msg is None, msg


def test_dense_fp32_vector_keeps_fp32_semantics(cuda_device):
    """upstream's Index(fp16=False) keeps an fp32 matrix and does an fp32 GEMM (index.py:36-44, 88-94).  Here the tensor
    cores sweep a bf16 copy and every passage within the bf16 error bound of the k-th score is re-scored exactly from
    its fp32 row: ids and scores must match the fp32 reference at 1e-5 (near-tie-aware), on data where a plain bf16
    index ranks differently."""
    import vsearch_b200 as vs

    g = torch.Generator().manual_seed(21)
    for n, d, B, k in ((150_000, 96, 37, 50), (1_100_000, 64, 9, 100), (3000, 200, 5, 100)):   # the last one: the sample sweep covers the whole index
        x = torch.randn(n, d, generator=g)
        q = torch.randn(B, d, generator=g)
        idx = vs.Index(fp16=False)
        idx.vector = x
        idx.move_to_device("cuda:0")
        assert idx._engine.store_dtype == 0
        res = idx.search(q, k)
        assert res.scores.dtype == torch.float32
        ref = ref_search.ref_scores(q, x)
        msg = ref_search.compare_results(res, ref, k, rtol=1e-5, exact=False)
        assert msg is None, f"n={n}: {msg}"
        if n == 150_000:   # the same vectors narrowed to bf16 do not reproduce the fp32 ranking: the re-score matters
            narrow = vs.Index()
            narrow.vector = x.to(torch.bfloat16)
            narrow.move_to_device("cuda:0")
            got = narrow.search(q, k)
            assert ref_search.compare_results(ref_search.SearchResults(got.ids, got.scores.float()), ref, k, rtol=1e-5, exact=False) is not None
    # grid data: exact in every precision -> bit-equal ids and scores, fp16 queries against an fp32 index
    xg, qg = _grid((40_000, 128), 5), _grid((6, 128), 6)
    idx = vs.Index(fp16=False)
    idx.vector = xg
    idx.move_to_device("cuda:0")
    res = idx.search(qg.to(torch.float16), 20)
    assert ref_search.compare_results(res, ref_search.ref_scores(qg, xg), 20, exact=True) is None


@pytest.mark.parametrize("n,d,B,k,world,dtype", [
    (1_300_000, 64, 5, 100, 2, torch.bfloat16),   # all three steps sweep rows on both ranks
    (400_000, 64, 300, 10, 4, torch.float16),     # CTA-pair kernel; last rank shorter
    (40_000, 96, 9, 100, 3, torch.bfloat16),      # step 1 ends the rows of a rank: step 2 sweeps nothing
    (260_000, 64, 4, 1000, 1, torch.bfloat16),    # one rank: the steps alone
])
def test_dense_stepwise_simulated_ranks(n, d, B, k, world, dtype, cuda_device):
    """vs_search_dense_step (the row-sharded dense path: thresholds pooled between the ranks after every sweep), with the
    ranks played one after the other on one GPU and torch.stack standing in for the all-gather; grid values, so ids and
    scores must equal the reference's bit for bit.  Massive ties at the k-th score included (integer grid)."""
    import vsearch_b200 as vs

    x, q = _grid((n, d), 11, scale=4.0, lim=8), _grid((B, d), 12, scale=4.0, lim=8)
    parts = [vs.row_partition(n, world, r) for r in range(world)]
    shards = [_dense_index(x[lo:hi], dtype) for lo, hi in parts]
    engines = [s._require_engine() for s in shards]
    rnd = shards[0]._score_round()   # scores are ranked after rounding to the index dtype (index.py:89)
    qd = engines[0]._prep_q(q)
    keys = [torch.empty((B, k), dtype=torch.int64, device="cuda:0") for _ in range(world)]
    status = [torch.zeros(1, dtype=torch.int32, device="cuda:0") for _ in range(world)]
    gathered = None
    for step in range(3):
        for r, eng in enumerate(engines):
            eng.search_dense_step(step, qd, k, world, gathered, keys[r], status[r], score_round=rnd, id_offset=parts[r][0])
        gathered = torch.stack(keys).contiguous()
    assert all(int(s.item()) == 0 for s in status)
    ids, scores = vs.merge_keys(gathered, k)
    canon = ref_search.canonical_topk(ref_search.quantize_like(ref_search.ref_scores(q, x), dtype), k)
    assert torch.equal(ids.cpu(), canon.ids), (ids.cpu() != canon.ids).nonzero()[:5]
    assert torch.equal(ref_search.quantize_like(scores.cpu(), dtype), canon.scores)


def test_dense_overflowing_lists_take_the_checked_fallback(cuda_device):
    """The dense call enqueues all its sweeps without reading anything back and polls one status word at the end.  Rows
    that beat everything the sample saw, 100,000 of them in the second sweep's range (the survivor lists hold 65,536),
    raise it, and the call must still come out exact through the checked, retrying path."""
    n, d, B, k = 1_200_000, 64, 4, 10
    g = torch.Generator().manual_seed(5)
    x = torch.zeros(n, d)
    x[:16384, 0] = 1.0
    hot = 20_000 + torch.randperm(900_000, generator=g)[:100_000]
    x[hot, 0] = 2.0 + (torch.arange(100_000) % 200).float()          # integers <= 201: exact in bf16
    q = torch.zeros(B, d)
    q[:, 0] = torch.tensor([1.0, 2.0, 0.5, 4.0])
    idx = _dense_index(x)
    res = idx.search(q, k)
    canon = ref_search.canonical_topk(ref_search.quantize_like(ref_search.ref_scores(q, x), torch.bfloat16), k)
    assert torch.equal(res.ids.cpu(), canon.ids)
    assert torch.equal(res.scores.float().cpu(), canon.scores)
