"""The Retriever facade end to end on the GPU: build_index / retrieve / save_index / load_index for the three
index types (upstream retriever.py:107-148, 284-348), numpy and tensor queries, rerank with a stand-in encoder."""
import json

import numpy as np
import pytest
import torch

from oracle import ref_search
from tests.util import sparse_queries, stratified_csr

pytestmark = pytest.mark.gpu
V = 4000


def _corpus(n=6000, m=24, seed=3, binary=False):
    crow, col, val = stratified_csr(n, V, m, seed=seed, grid=True, binary=binary, jitter=8)
    return ref_search.torch_csr(crow, col, val, (n, V))


def test_sparse_build_retrieve_save_load(tmp_path, cuda_device):
    import vsearch_b200 as vs

    X = _corpus()
    texts = [f"passage {i}" for i in range(X.shape[0])]
    r = vs.Retriever(device="cuda:0")
    r.build_index(texts, index_type="sparse", vectors=X.to_dense())  # dense -> to_sparse_csr(), like upstream :304
    assert r.index_type == vs.IndexType.SPARSE and len(r.index) == len(texts)
    q = sparse_queries(5, V, 40, seed=1)
    ref = ref_search.ref_scores(q, X)
    res = r.retrieve(q.numpy(), k=7)  # ndarray queries (upstream process_query :96-97)
    assert ref_search.compare_results(res, ref, 7, exact=True) is None
    assert r.retrieve(q, k=5).ids.shape == (5, 5)  # default-k style call with a tensor
    assert r.index.get_sample(int(res.ids[0, 0])) == texts[int(res.ids[0, 0])]
    # explicit index= argument is honoured (upstream resolves it and then ignores it)
    other = vs.SparseIndex()
    other.vector = _corpus(seed=9)
    other.move_to_device("cuda:0")
    res_o = r.retrieve(q, k=7, index=other)
    assert ref_search.compare_results(res_o, ref_search.ref_scores(q, other.vector), 7, exact=True) is None
    # save -> load (fp16=True is upstream's default on load: grid values survive the fp16 round trip)
    p = str(tmp_path / "index0.npz")
    r.save_index(p)
    data = tmp_path / "texts.jsonl"
    data.write_text("\n".join(json.dumps(t) for t in texts) + "\n")
    r2 = vs.Retriever(device="cuda:0")
    r2.load_index(p, str(data))
    assert isinstance(r2.index, vs.SparseIndex) and r2.index.vector.values().dtype == torch.float16
    res2 = r2.retrieve(q, k=7)
    # an fp16 index returns (and ranks by) fp16 scores: index.py:89 dtype contract
    canon16 = ref_search.canonical_topk(ref_search.quantize_like(ref, torch.float16), 7)
    assert torch.equal(res2.ids.cpu(), canon16.ids) and res2.scores.dtype == torch.float16
    assert torch.equal(res2.scores.float().cpu(), canon16.scores)
    assert "SparseIndex" in str(r2.index) and "cuda" in str(r2.index)


def test_bag_of_token_build_and_rerank(cuda_device):
    import vsearch_b200 as vs

    X = _corpus(binary=True)
    dense_params = torch.rand(X.shape[0], V, generator=torch.Generator().manual_seed(5))  # "parametric" passage vectors
    texts = list(range(X.shape[0]))

    class FakeEncoderP:  # stands in for upstream's encoder_p.embed(texts) (retriever.py:139)
        def embed(self, t, batch_size=32, require_grad=False):
            return dense_params[torch.tensor(t)]

    r = vs.Retriever(device="cuda:0", encoder_p=FakeEncoderP())
    r.build_index(texts, index_type=vs.IndexType.BAG_OF_TOKEN, vectors=X)
    assert isinstance(r.index, vs.BoTIndex) and r.index._require_engine().kind == 2  # binary: no values on device
    q = (sparse_queries(4, V, 60, seed=2) != 0).float()
    k = 20
    res = r.retrieve(q, k=k)
    assert ref_search.compare_results(res, ref_search.ref_scores(q, X), k, exact=True) is None
    rr = r.retrieve(q, k=k, rerank=True)
    # reranked ids are a permutation of the retrieved ids, ordered by the parametric score q . p
    assert torch.equal(rr.ids.sort(dim=1).values.cpu(), res.ids.sort(dim=1).values.cpu())
    want = torch.gather(torch.matmul(q, dense_params.t()), 1, rr.ids.cpu())
    torch.testing.assert_close(rr.scores.cpu(), want, rtol=1e-5, atol=1e-5)
    assert bool((rr.scores[:, 1:] <= rr.scores[:, :-1]).all())


def test_dense_build_save_load(tmp_path, cuda_device):
    import vsearch_b200 as vs

    g = torch.Generator().manual_seed(11)
    x = torch.randint(-16, 17, (5000, 128), generator=g).float() / 8.0
    q = torch.randint(-16, 17, (6, 128), generator=g).float() / 8.0
    r = vs.Retriever(device="cuda:0")
    r.build_index(None, index_type="dense", vectors=x.to(torch.bfloat16))
    res = r.retrieve(q, k=9)
    canon = ref_search.canonical_topk(ref_search.quantize_like(ref_search.ref_scores(q, x), torch.bfloat16), 9)
    assert torch.equal(res.ids.cpu(), canon.ids) and res.scores.dtype == torch.bfloat16
    # two .pt shards, concatenated in sorted-glob order (intent of upstream index.py:36-44)
    torch.save(x[:3000].to(torch.float16), str(tmp_path / "emb0.pt"))
    torch.save(x[3000:].to(torch.float16), str(tmp_path / "emb1.pt"))
    r2 = vs.Retriever(device="cuda:0")
    r2.load_index(str(tmp_path / "emb*.pt"))
    assert r2.index_type == vs.IndexType.DENSE and tuple(r2.index.vector.shape) == (5000, 128)
    res2 = r2.retrieve(q, k=9)
    canon16 = ref_search.canonical_topk(ref_search.quantize_like(ref_search.ref_scores(q, x), torch.float16), 9)
    assert torch.equal(res2.ids.cpu(), canon16.ids)
    r2.save_index(str(tmp_path / "all.pt"))
    assert torch.equal(torch.load(str(tmp_path / "all.pt")), x.to(torch.float16))


class _WordTokenizer:
    """Stand-in for the BERT tokenizer of upstream's encoder_p: [CLS] words... [SEP], ids from a fixed vocabulary."""

    def __init__(self, vocab_size=3000):
        self.vocab = {f"w{i}": i for i in range(vocab_size)}

    def __call__(self, texts, max_length=128, truncation=True):
        out = []
        for t in texts:
            ids = [101] + [self.vocab[w] for w in t.split() if w in self.vocab]
            out.append((ids[:max_length - 1] if truncation else ids) + [102])
        return {"input_ids": out}


class _TokEncoder:
    def __init__(self):
        self.tokenizer = _WordTokenizer()


def test_bag_of_token_build_from_texts(cuda_device):
    """build_index(texts, bag_of_token) with a tokenizer-bearing encoder: upstream retriever.py:208-253, 307-312."""
    import vsearch_b200 as vs

    g = torch.Generator().manual_seed(2)
    texts = [" ".join(f"w{int(j)}" for j in torch.randint(900, 3000, (int(torch.randint(1, 200, (1,), generator=g)),),
                                                          generator=g)) for _ in range(500)]
    r = vs.Retriever(encoder_p=_TokEncoder(), device="cuda:0")
    r.build_index(texts, index_type="bag_of_token")
    assert isinstance(r.index, vs.BoTIndex) and r.index.data is texts
    rows = _WordTokenizer()(texts)["input_ids"]
    crow, col, shape = ref_search.ref_bot_rows(rows, vocab_size=3000, num_shift=999)
    got = r.index.vector.cpu()
    assert tuple(got.shape) == shape and got.values().dtype == torch.float16
    assert torch.equal(got.crow_indices().to(torch.int64), crow) and torch.equal(got.col_indices().to(torch.int64), col)
    q = sparse_queries(3, shape[1], 30, seed=4)
    X = ref_search.torch_csr(crow, col, torch.ones(col.numel()), shape)
    res = r.retrieve(q, k=5)
    assert ref_search.compare_results(ref_search.SearchResults(res.ids, res.scores.float()), ref_search.ref_scores(q, X), 5,
                                      exact=True) is None


def test_rerank_with_precomputed_vectors(cuda_device):
    """First stage on the bag-of-token index, second stage = gather the candidates' parametric rows (vs_score_rows):
    the same ids/scores as upstream's rerank (retriever.py:137-147) with the re-embedding replaced by row lookups."""
    import vsearch_b200 as vs

    n, k = 20_000, 50
    crow, col, val = stratified_csr(n, V, 30, seed=4, grid=True, jitter=10)
    Xp = ref_search.torch_csr(crow, col, val, (n, V))                      # parametric vectors (valued)
    Xb = ref_search.torch_csr(crow, col, torch.ones_like(val), (n, V))     # their bag-of-token support
    r = vs.Retriever(device="cuda:0")
    r.build_index([f"p{i}" for i in range(n)], index_type="bag_of_token", vectors=Xb)
    par = vs.SparseIndex()
    par.vector = Xp
    par.move_to_device("cuda:0")
    q = sparse_queries(6, V, 40, seed=8)
    first = r.retrieve(q, k=k)
    res = r.retrieve(q, k=k, rerank=True, rerank_index=par)
    # restated upstream rerank on the same candidates: p_emb = Xp[ids]; bmm with q; sort (ties keep first-stage order)
    dense_p = Xp.to_dense()
    cand = first.ids.cpu()
    sc = torch.einsum("bkv,bv->bk", dense_p[cand], q)
    order = torch.sort(sc, dim=-1, descending=True, stable=True)
    assert torch.equal(res.ids.cpu(), torch.gather(cand, 1, order.indices))
    assert torch.equal(res.scores.float().cpu(), order.values)
    # ids outside the index score -inf; 1-D query / 1-D ids round-trip
    bad = par.score_rows(q[:1], torch.tensor([[0, -1, n, 5]]))
    assert torch.isinf(bad[0, 1]) and torch.isinf(bad[0, 2]) and bad[0, 0] == (dense_p[0] * q[0]).sum()
    assert par.score_rows(q[0], torch.tensor([3, 4])).shape == (2,)
    bot_scores = r.index.score_rows(q, first.ids)                          # binary index: recovers the first-stage scores
    assert torch.equal(bot_scores.float(), first.scores.float())
