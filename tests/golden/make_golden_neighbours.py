"""Golden fixtures for the neighbours of the hot path, produced by the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_neighbours.py

* ``sparsify_k768.npz``   -- ``topk_sparsify`` / ``build_topk_mask`` / ``build_bow_mask`` imported from
  /root/reference/src/ir/utils/sparse.py, combined the way encoder/vdr.py:159-169 combines them
  (``mask = logical_or(bow_mask, topk_mask); emb *= mask``), on a seeded continuous batch (no ties at the k-th value).
* ``bot_rows_*.npz``      -- ``Retriever._build_bot_vectors`` (retriever.py:208-253).  The class cannot be imported
  (its module pulls in the encoders), so the function's own source is taken from the reference file with
  ``ast`` at generation time and executed unmodified against a stand-in tokenizer; ``get_first_unique_n`` is imported
  from /root/reference/src/ir/retriever/index_utils.py.  Nothing of the reference is copied into the repository.
"""
import ast
import importlib.util
import os
from typing import List  # noqa: F401  (name used by the extracted function's annotations)

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/ir"


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


class FakeTokenizer:
    def __init__(self, rows, vocab_size):
        self.rows, self.vocab = rows, {i: i for i in range(vocab_size)}

    def __call__(self, texts, max_length=128, truncation=True):
        return {"input_ids": [self.rows[int(t)][:max_length] for t in texts]}


def main():
    sparse = load(os.path.join(REF, "utils", "sparse.py"), "ref_sparse")
    iu = load(os.path.join(REF, "retriever", "index_utils.py"), "ref_index_utils")

    # ---- sparsifier
    g = torch.Generator().manual_seed(17)
    V, shift, k = 29523, 999, 768
    emb = torch.rand(5, V, generator=g) * 3 - 0.5
    ids = torch.randint(0, V + shift, (5, 24), generator=g)
    plain = sparse.topk_sparsify(emb.clone(), k)
    topk_mask = sparse.build_topk_mask(emb, k)
    bow_mask = sparse.build_bow_mask(ids, vocab_size=V + shift, shift_num=shift).bool()
    lexical = emb * torch.logical_or(bow_mask, topk_mask)
    np.savez_compressed(os.path.join(HERE, "sparsify_k768.npz"), emb=emb.numpy(), token_ids=ids.numpy(), k=k, shift=shift,
                        plain_idx=plain.nonzero().numpy(), plain_val=plain[plain != 0].numpy(),
                        lexical_idx=lexical.nonzero().numpy(), lexical_val=lexical[lexical != 0].numpy())
    print("sparsify", int((plain != 0).sum()), int((lexical != 0).sum()))

    # ---- bag-of-token rows: the reference's own function body, executed as is
    src = open(os.path.join(REF, "retriever", "retriever.py")).read()
    fn = next(n for n in ast.walk(ast.parse(src)) if isinstance(n, ast.FunctionDef) and n.name == "_build_bot_vectors")
    ns = {"torch": torch, "List": List, "tqdm": (lambda x, **kw: x), "get_first_unique_n": iu.get_first_unique_n,
          "csr_array": None}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "reference:_build_bot_vectors", "exec"), ns)
    rng = np.random.default_rng(5)
    vocab = 30522
    rows = []
    for i in range(70):
        n = int(rng.integers(0, 140))
        r = np.concatenate(([101], rng.integers(900, 1500, n // 2), rng.integers(0, vocab, n - n // 2), [102])).tolist()
        rows.append(r)
    rows[3] = [101, 102]
    maxlen = max(len(r) for r in rows)
    padded = np.zeros((len(rows), maxlen), dtype=np.int32)
    for i, r in enumerate(rows):
        padded[i, :len(r)] = r

    class Self:
        pass

    me = Self()
    me.encoder_p = Self()
    me.encoder_p.tokenizer = FakeTokenizer(rows, vocab)
    for tag, max_token in (("all", None), ("first20", 20)):
        # ONE batch: across batches the reference appends views of one reused buffer (retriever.py:234-247) and converts
        # them after the loop, so earlier batches come out as copies of the last one -- an upstream bug that is not part of
        # the contract and is not reproduced
        csr = ns["_build_bot_vectors"](me, [str(i) for i in range(len(rows))], batch_size=128, max_len=128, max_token=max_token)
        np.savez_compressed(os.path.join(HERE, f"bot_rows_{tag}.npz"), token_ids=padded,
                            lengths=np.array([len(r) for r in rows], dtype=np.int32), max_len=128, vocab=vocab, shift=999,
                            max_token=0 if max_token is None else max_token, crow=csr.crow_indices().numpy(),
                            col=csr.col_indices().numpy(), val=csr.values().float().numpy(), val_dtype=str(csr.values().dtype),
                            shape=np.array(csr.shape))
        print("bot rows", tag, tuple(csr.shape), csr.values().dtype, int(csr.crow_indices()[-1]))


if __name__ == "__main__":
    main()
