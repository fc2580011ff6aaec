"""The oracle (torch restatement + plain-C restatement) against the golden vectors that the
unmodified reference produced (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import c_oracle, ref_search
import os

from tests.util import GOLDEN, golden_search_cases, load_golden


@pytest.mark.parametrize("name", golden_search_cases("csr"))
def test_torch_oracle_matches_reference_csr(name):
    z = load_golden(name)
    X = ref_search.torch_csr(z["crow"], z["col"].astype(np.int64), z["val"], z["shape"])
    q = torch.from_numpy(z["q"])
    scores = ref_search.ref_scores(q, X)
    assert torch.equal(scores, torch.from_numpy(z["ref_scores"]))  # same torch entry points -> same bits
    res = ref_search.ref_search(q, X, z["k"])
    assert torch.equal(res.scores, torch.from_numpy(z["ref_topk_scores"]))
    canon = ref_search.canonical_topk(scores, z["k"])
    # reference topk values == canonical values (tie order only affects ids)
    assert torch.equal(canon.scores, torch.from_numpy(z["ref_topk_scores"]) + 0.0)
    # canonical ids are a valid answer for the reference: gathering reproduces its values
    assert torch.equal(torch.gather(scores, -1, canon.ids) + 0.0, canon.scores)
    # ids strictly ascending inside ties
    s, i = canon.scores.reshape(-1, z["k"]), canon.ids.reshape(-1, z["k"])
    tie = s[:, 1:] == s[:, :-1]
    assert bool((i[:, 1:][tie] > i[:, :-1][tie]).all())


@pytest.mark.parametrize("name", golden_search_cases("dense"))
def test_torch_oracle_matches_reference_dense(name):
    z = load_golden(name)
    scores = ref_search.ref_scores(torch.from_numpy(z["q"]), torch.from_numpy(z["x"]))
    assert torch.equal(scores, torch.from_numpy(z["ref_scores"]))
    res = ref_search.ref_search(torch.from_numpy(z["q"]), torch.from_numpy(z["x"]), z["k"])
    assert torch.equal(res.scores, torch.from_numpy(z["ref_topk_scores"]))


@pytest.mark.parametrize("name", golden_search_cases("csr"))
def test_c_oracle_matches_reference_csr(name):
    z = load_golden(name)
    q = z["q"].reshape(-1, z["q"].shape[-1])
    val = None if bool(z["binary"]) else z["val"]
    sc = c_oracle.csr_scores(z["crow"], z["col"], val, z["shape"], q)
    ref = z["ref_scores"].reshape(q.shape[0], -1)
    grid = "cont" not in name
    if grid:
        assert np.array_equal(sc, ref)
    else:
        np.testing.assert_allclose(sc, ref, rtol=1e-5, atol=1e-6)
    k = z["k"]
    ids, tsc = c_oracle.topk(ref, k)
    canon = ref_search.canonical_topk(torch.from_numpy(ref), k)
    assert np.array_equal(ids, canon.ids.numpy())
    assert np.array_equal(tsc, canon.scores.numpy())
    ids2, tsc2 = c_oracle.csr_search(z["crow"], z["col"], val, z["shape"], q, k)
    if grid:
        assert np.array_equal(ids2, canon.ids.numpy())
        assert np.array_equal(tsc2, canon.scores.numpy())
    else:
        msg = ref_search.compare_results(
            ref_search.SearchResults(torch.from_numpy(ids2), torch.from_numpy(tsc2)), torch.from_numpy(ref), k)
        assert msg is None, msg


@pytest.mark.parametrize("name", golden_search_cases("dense"))
def test_c_oracle_matches_reference_dense(name):
    z = load_golden(name)
    sc = c_oracle.dense_scores(z["x"], z["q"])
    if "grid" in name:
        assert np.array_equal(sc, z["ref_scores"])
    else:
        np.testing.assert_allclose(sc, z["ref_scores"], rtol=1e-5, atol=1e-5)


def test_k_out_of_range_raises_like_reference():
    z = load_golden("sparse_small_keqn")
    X = ref_search.torch_csr(z["crow"], z["col"].astype(np.int64), z["val"], z["shape"])
    with pytest.raises(RuntimeError):
        ref_search.ref_search(torch.from_numpy(z["q"]), X, int(z["shape"][0]) + 1)
    with pytest.raises(RuntimeError):
        ref_search.oracle_search(torch.from_numpy(z["q"]), X, int(z["shape"][0]) + 1)
    with pytest.raises(RuntimeError):
        c_oracle.topk(z["ref_scores"], int(z["shape"][0]) + 1)


def test_all_zero_query_gives_lowest_ids():
    z = load_golden("sparse_neg_k25")
    canon = ref_search.canonical_topk(torch.from_numpy(z["ref_scores"]), z["k"])
    assert canon.ids[1].tolist() == list(range(z["k"]))
    assert bool((canon.scores[1] == 0).all())


def test_merge_shard_results_equals_global():
    g = torch.Generator().manual_seed(3)
    scores = torch.randint(0, 6, (4, 300), generator=g).float()  # heavy ties
    k = 17
    full = ref_search.canonical_topk(scores, k)
    bounds = [0, 90, 91, 200, 300]
    ids, sc = [], []
    for a, b in zip(bounds[:-1], bounds[1:]):
        kk = min(k, b - a)
        part = ref_search.canonical_topk(scores[:, a:b], kk)
        ids.append(part.ids + a)
        sc.append(part.scores)
    merged = ref_search.merge_shard_results(ids, sc, k)
    assert torch.equal(merged.ids, full.ids) and torch.equal(merged.scores, full.scores)


def test_one_d_query_shape():
    z = load_golden("sparse_1d_k7")
    X = ref_search.torch_csr(z["crow"], z["col"].astype(np.int64), z["val"], z["shape"])
    res = ref_search.oracle_search(torch.from_numpy(z["q"]), X, z["k"])
    assert tuple(res.ids.shape) == (7,) and tuple(z["ref_topk_ids"].shape) == (7,)
    assert torch.equal(res.scores, torch.from_numpy(z["ref_topk_scores"]))


def test_ref_bot_rows_matches_reference_semantics():
    """hand-checked case of the bag-of-token construction (retriever.py:232-251): duplicates collapse, ids below the
    shift vanish, columns are renumbered and ascending, max_token keeps the first distinct ids in order."""
    rows = [[101, 2054, 2003, 2054, 1996, 102], [101, 102], [5000, 999, 998, 5000]]
    crow, col, shape = ref_search.ref_bot_rows(rows, vocab_size=30522, num_shift=999)
    assert shape == (3, 29523)
    assert crow.tolist() == [0, 3, 3, 5]
    assert col.tolist() == [1996 - 999, 2003 - 999, 2054 - 999, 0, 5000 - 999]
    crow, col, _ = ref_search.ref_bot_rows(rows, vocab_size=30522, num_shift=999, max_token=3)
    assert crow.tolist() == [0, 2, 2, 4]          # row 0 keeps {101, 2054, 2003}; row 2 keeps {5000, 999, 998}
    assert col.tolist() == [2003 - 999, 2054 - 999, 0, 5000 - 999]


@pytest.mark.parametrize("tag", ["all", "first20"])
def test_ref_bot_rows_matches_reference_output(tag):
    """oracle restatement of the bag-of-token builder vs the output of the reference's own _build_bot_vectors
    (tests/golden/make_golden_neighbours.py executes the unmodified function body)."""
    z = np.load(os.path.join(GOLDEN, f"bot_rows_{tag}.npz"))
    # the tokenizer truncates to max_len (retriever.py:238) before the builder sees the ids
    rows = [z["token_ids"][i, :min(int(z["lengths"][i]), int(z["max_len"]))].tolist() for i in range(z["token_ids"].shape[0])]
    crow, col, shape = ref_search.ref_bot_rows(rows, int(z["vocab"]), int(z["shift"]), int(z["max_token"]) or None)
    assert shape == tuple(z["shape"])
    assert np.array_equal(crow.numpy(), z["crow"]) and np.array_equal(col.numpy(), z["col"])
    assert str(z["val_dtype"]) == "torch.float16" and bool((z["val"] == 1).all())


def test_ref_topk_sparsify_matches_reference_output():
    """oracle restatement of the sparsifier vs the reference's topk_sparsify / build_topk_mask / build_bow_mask
    (utils/sparse.py) combined as in encoder/vdr.py:159-169."""
    z = np.load(os.path.join(GOLDEN, "sparsify_k768.npz"))
    emb, ids = torch.from_numpy(z["emb"]), torch.from_numpy(z["token_ids"])
    for name, out in (("plain", ref_search.ref_topk_sparsify(emb, int(z["k"]))),
                      ("lexical", ref_search.ref_topk_sparsify(emb, int(z["k"]), ids, int(z["shift"])))):
        assert np.array_equal(out.nonzero().numpy(), z[f"{name}_idx"]), name
        assert np.array_equal(out[out != 0].numpy(), z[f"{name}_val"]), name
