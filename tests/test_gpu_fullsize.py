"""BASELINE.json's full sizes (the oracle cannot run there): size-independent properties.

* config 2 (21,015,324 x 29,523 binary, 120 tokens/row): the passage-major scan and the token-major inverted lists are
  two independent kernel families over two different index layouts; on dyadic-grid queries every summation order
  gives the same bits, so their ids and scores must be IDENTICAL, sorted, and reproduced by re-scoring the returned
  rows (vs_score_rows); two virtual shards merged = the whole index.
* config 4 (21,015,324 x 768 bf16 dense): returned scores = <q, X[id]> recomputed by torch on the gathered rows,
  sorted with ties -> lower id, and no row of a random 1M-row slice beats the k-th result."""
import pytest
import torch

pytestmark = pytest.mark.gpu
V, N = 29523, 21_015_324


def _strat_cols(n, m, seed, dev):
    g = torch.Generator(device=dev).manual_seed(seed)
    w = V // m
    base = ((torch.arange(m, device=dev, dtype=torch.int64) * V) // m).to(torch.int32)
    out = torch.empty((n, m), dtype=torch.int32, device=dev)
    for lo in range(0, n, 1 << 20):
        hi = min(n, lo + (1 << 20))
        out[lo:hi] = torch.randint(0, w, (hi - lo, m), generator=g, device=dev, dtype=torch.int32) + base[None, :]
    return out


def test_cfg2_full_size_scan_and_inverted_agree(cuda_device):
    import vsearch_b200 as vs

    dev = torch.device("cuda:0")
    m, B, k = 120, 12, 100
    cols = _strat_cols(N, m, 1234, dev).reshape(-1)
    crow = torch.arange(N + 1, device=dev, dtype=torch.int64) * m
    idx = vs.BoTIndex.from_token_csr(crow, cols, (N, V), device=dev, dtype=torch.float32)
    g = torch.Generator().manual_seed(5)
    q = torch.zeros(B, V)
    for b in range(B):
        nnz = [8, 64, 300, 768][b % 4]
        c = torch.randperm(V, generator=g)[:nnz]
        q[b, c] = torch.randint(1, 256, (nnz,), generator=g).float() / 64.0      # dyadic grid: exact sums
    q[3, torch.randperm(V, generator=g)[:5]] = -2.0                              # some negative weights
    out = {}
    for mode in ("scan", "inverted"):
        idx.search_mode = mode
        out[mode] = idx.search(q, k)
        assert idx.last_mode() == mode
    a, b_ = out["scan"], out["inverted"]
    assert torch.equal(a.ids, b_.ids) and torch.equal(a.scores, b_.scores)
    s = a.scores.float()
    assert bool((s[:, :-1] >= s[:, 1:]).all())
    tie = s[:, :-1] == s[:, 1:]
    assert bool((a.ids[:, :-1][tie] < a.ids[:, 1:][tie]).all())                  # ties -> lower id first
    assert int(a.ids.min()) >= 0 and int(a.ids.max()) < N
    assert torch.equal(idx.score_rows(q, a.ids).float(), s)                      # scores belong to the returned rows
    # two virtual shards (rows [0, h) and [h, N)) merged = the whole index
    h = 10_000_000
    lo = vs.BoTIndex.from_token_csr(crow[:h + 1], cols[:h * m], (h, V), device=dev, dtype=torch.float32)
    hi = vs.BoTIndex.from_token_csr(crow[h:] - h * m, cols[h * m:], (N - h, V), device=dev, dtype=torch.float32)
    keys = torch.stack([lo.search_keys(q, k, id_offset=0), hi.search_keys(q, k, id_offset=h)])
    ids, sc = vs.merge_keys(keys, k)
    assert torch.equal(ids, a.ids) and torch.equal(sc.float(), s)


def test_cfg4_full_size_dense_properties(cuda_device):
    import vsearch_b200 as vs

    dev = torch.device("cuda:0")
    D, B, k = 768, 300, 100
    g = torch.Generator(device=dev).manual_seed(7)
    x = torch.empty((N, D), dtype=torch.bfloat16, device=dev)
    for lo in range(0, N, 1 << 20):
        hi = min(N, lo + (1 << 20))
        x[lo:hi] = (torch.randint(-8, 9, (hi - lo, D), generator=g, device=dev, dtype=torch.int32).float() / 8.0).to(torch.bfloat16)
    q = (torch.randint(-8, 9, (B, D), generator=g, device=dev, dtype=torch.int32).float() / 8.0).to(torch.bfloat16)
    idx = vs.Index(fp16=False)
    idx.vector = x
    idx.move_to_device(dev)
    res = idx.search(q, k)
    assert res.scores.dtype == torch.bfloat16 and int(res.ids.min()) >= 0 and int(res.ids.max()) < N
    # scores = <q, X[id]> (multiples of 1/64, |s| <= 768: exact in fp32), rounded to bf16 like the index dtype
    re = torch.einsum("bkd,bd->bk", x[res.ids].float(), q.float()).to(torch.bfloat16)
    assert torch.equal(re, res.scores)
    s = res.scores.float()
    assert bool((s[:, :-1] >= s[:, 1:]).all())
    tie = s[:, :-1] == s[:, 1:]
    assert bool((res.ids[:, :-1][tie] < res.ids[:, 1:][tie]).all())
    assert all(len(set(r)) == k for r in res.ids[:8].tolist())                   # no duplicate ids
    # nothing in a random slice of the index beats the k-th result (ties at the k-th score may only have larger ids)
    start = 7_654_321
    sl = (q.float() @ x[start:start + 1_000_000].float().t()).to(torch.bfloat16).float()
    assert _no_missed(sl, s, res.ids, start)


def _no_missed(sl, s, ids, start):
    """exact check: every slice row scoring above the k-th result must be among the returned ids"""
    B = sl.shape[0]
    for b in range(B):
        better = (sl[b] > s[b, -1]).nonzero().flatten() + start
        if better.numel() and not bool(torch.isin(better, ids[b]).all()):
            return False
    return True
