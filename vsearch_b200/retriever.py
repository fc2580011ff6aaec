"""Encoder-less ``Retriever`` facade: the part of upstream ``src/ir/retriever/retriever.py`` that sits on
the index-scoring hot path -- ``process_query`` (tensor / ndarray branches, :96-103), ``retrieve``
(:107-148), ``build_index`` (:284-317), ``save_index`` (:319-320), ``load_index`` (:322-348).

The encoders (BERT / CLIP towers) are out of scope: string queries and text corpora
need user-supplied ``encoder_q`` / ``encoder_p`` objects exposing upstream's ``embed(texts, ...)``; tensor
queries and pre-computed passage vectors need none.
"""
from __future__ import annotations

from typing import List, Optional, Union

import numpy as np
import torch
import torch.nn.functional as F

from .index import BoTIndex, Index, IndexType, SearchResults, SparseIndex


class Retriever:
    def __init__(self, device: str = "cuda", encoder_q=None, encoder_p=None, topk: int = 768):
        self.device = device
        self.encoder_q = encoder_q
        self.encoder_p = encoder_p
        self.topk = topk  # upstream: encoder_q.config.topk (conf/biencoder/vdr.yaml:11)
        self.index: Optional[Index] = None
        self.index_type: Optional[IndexType] = None

    # ---- queries ----------------------------------------------------------------------------------
    def process_query(self, queries: Union[str, List[str], np.ndarray, torch.Tensor], dropout: float = 0,
                      a: Optional[int] = None, batch_size: int = 32) -> torch.Tensor:
        num_activation = a or self.topk
        if isinstance(queries, str):
            queries = [queries]
        if isinstance(queries, list) and queries and isinstance(queries[0], str):
            if self.encoder_q is None:
                raise NotImplementedError("string queries need an encoder_q (encoders are out of scope here)")
            q_emb = self.encoder_q.embed(queries, batch_size=batch_size, topk=num_activation)
        elif isinstance(queries, np.ndarray):
            q_emb = torch.Tensor(queries)
        elif isinstance(queries, torch.Tensor):
            q_emb = queries
        else:
            raise NotImplementedError(f"Query type {type(queries)} not supported")
        if dropout:
            q_emb = F.dropout(q_emb, p=dropout)
        return q_emb

    def retrieve(self, queries, k: int = 5, dropout: float = 0, a: Optional[int] = None,
                 index: Optional[Index] = None, rerank: bool = False, batch_size: int = 32,
                 rerank_index: Optional[Index] = None) -> SearchResults:
        """``rerank_index`` (extension): a device-resident sparse index holding the passages' parametric vectors, row
        for row with the bag-of-token index.  With it the rerank stage re-scores the k candidates by gathering their
        rows on the GPU instead of re-embedding their texts (upstream retriever.py:137-147)."""
        index = index if index is not None else self.index  # upstream resolves this, then ignores it (:133,136)
        if index is None:
            raise RuntimeError("no index: call build_index() or load_index() first")
        q_emb = self.process_query(queries, dropout, a, batch_size=batch_size)
        results = index.search(q_emb, k=k)
        if rerank and index.index_type == IndexType.BAG_OF_TOKEN:
            if rerank_index is not None:
                results = self._rerank_rows(rerank_index, q_emb, results)
            elif self.encoder_p is None:
                raise NotImplementedError("rerank needs an encoder_p to re-embed the retrieved passages, or a rerank_index")
            else:
                results = self._rerank(index, q_emb, results, k, batch_size)
        return results

    @staticmethod
    def _rerank_rows(rerank_index, q_emb, results):
        """Re-score the candidates against their precomputed vectors (``vs_score_rows``), re-sort (stable: equal scores
        keep the first-stage order)."""
        ids = results.ids.reshape(-1, results.ids.shape[-1])
        sc = rerank_index.score_rows(q_emb.reshape(ids.shape[0], -1), ids)
        order = torch.sort(sc.float(), dim=-1, descending=True, stable=True)
        new_ids = torch.gather(ids, 1, order.indices).reshape(results.ids.shape)
        return SearchResults(new_ids, torch.gather(sc, 1, order.indices).reshape(results.ids.shape))

    def _rerank(self, index, q_emb, results, k, batch_size):
        """upstream retriever.py:137-147: re-embed the k texts, dot with the query, re-sort."""
        ret_indices = results.ids
        texts = [index.get_sample(i) for i in ret_indices.flatten().tolist()]
        p_emb = self.encoder_p.embed(texts, batch_size=batch_size, require_grad=False)
        q2 = q_emb.reshape(-1, q_emb.shape[-1])
        p_emb = p_emb.view(-1, k, q2.shape[-1]).to(ret_indices.device)
        sc = torch.bmm(p_emb.float(), q2.unsqueeze(-1).to(ret_indices.device).float()).squeeze(-1)
        order = torch.sort(sc, dim=-1, descending=True, stable=True)
        ids = torch.gather(ret_indices.reshape(-1, k), 1, order.indices)
        return SearchResults(ids, order.values)

    # ---- index construction -------------------------------------------------------------------------
    def build_index(self, texts=None, batch_size: int = 32, index_type=IndexType.DENSE, bag_of_token: bool = False,
                    vectors: Optional[torch.Tensor] = None):
        """Assemble an index from pre-computed passage vectors (``vectors``: dense ``[N, V]`` / ``[N, D]`` tensor or
        a torch CSR tensor) or, with an ``encoder_p``, from texts (upstream retriever.py:284-317)."""
        if isinstance(index_type, str):
            index_type = IndexType(str(index_type).lower())
        elif not isinstance(index_type, IndexType):
            raise TypeError("index_type must be an instance of IndexType, int, or str.")
        self.index_type = index_type
        if vectors is None and index_type == IndexType.BAG_OF_TOKEN and getattr(self.encoder_p, "tokenizer", None) is not None:
            # upstream _build_bot_vectors (retriever.py:208-253): tokenize (max_length 128, truncation), set of token
            # ids per passage, drop ids < 999.  The tokenizer runs on the host; the rows are built on the GPU.
            self.index = self._build_bot_index(list(texts), batch_size=batch_size)
            self.index.data = texts
            return
        if vectors is None:
            if self.encoder_p is None:
                raise NotImplementedError("build_index from texts needs an encoder_p; pass vectors= instead")
            vectors = self.encoder_p.embed(list(texts), batch_size=batch_size)
        if index_type == IndexType.DENSE:
            self.index = Index()
            self.index.vector = vectors
        elif index_type == IndexType.SPARSE:
            self.index = SparseIndex()
            # a dense [N, V] matrix is sparsified on the GPU when the index moves there (vs_dense_to_csr) instead of
            # upstream's vectors.to_sparse_csr() (retriever.py:304)
            self.index.vector = vectors if vectors.layout in (torch.sparse_csr, torch.strided) else vectors.to_sparse_csr()
        elif index_type == IndexType.BAG_OF_TOKEN:
            self.index = BoTIndex()
            if vectors.layout == torch.strided:
                vectors = (vectors != 0).to(torch.float16)  # upstream: fp16 ones (:232-251); sparsified on the GPU
            elif vectors.layout != torch.sparse_csr:
                vectors = vectors.to_sparse_csr()
            self.index.vector = vectors
        else:
            raise NotImplementedError
        self.index.data = texts
        self.index.move_to_device(self.device)

    def _build_bot_index(self, texts, batch_size: int = 32, max_len: int = 128, max_token=None, num_shift: int = 999):
        tok = self.encoder_p.tokenizer
        vocab_size = len(tok.vocab)
        ids = torch.zeros((len(texts), max_len), dtype=torch.int32)
        lens = torch.zeros(len(texts), dtype=torch.int32)
        for b0 in range(0, len(texts), batch_size):
            for i, row in enumerate(tok(texts[b0:b0 + batch_size], max_length=max_len, truncation=True)["input_ids"]):
                row = row[:max_len]
                ids[b0 + i, :len(row)] = torch.as_tensor(row, dtype=torch.int32)
                lens[b0 + i] = len(row)
        return BoTIndex.from_token_ids(ids, lens, vocab_size=vocab_size, num_shift=num_shift, max_token=max_token,
                                       device=self.device, dtype=torch.float16)

    def save_index(self, path):
        self.index.save(path)

    def load_index(self, index_file=None, data_file=None, index_type=None):
        if index_type is None:
            if index_file.endswith(".pt"):
                index_type = IndexType.DENSE
            elif index_file.endswith(".npz"):
                index_type = IndexType.SPARSE
            else:
                raise ValueError("Cannot infer index type from file extension. Please provide 'index_type' explicitly.")
        elif isinstance(index_type, str):
            index_type = IndexType(index_type.lower())
        elif not isinstance(index_type, IndexType):  # upstream rejects the enum too; accepting it is harmless
            raise TypeError("index_type must be an instance of IndexType, int, or str.")
        self.index_type = index_type
        cls = {IndexType.DENSE: Index, IndexType.SPARSE: SparseIndex, IndexType.BAG_OF_TOKEN: BoTIndex}[index_type]
        self.index = cls(index_file, data_file, device=self.device)
