"""Index containers of the B200 engine: same public surface as upstream
``src/ir/retriever/index.py`` (``SearchResults``, ``IndexType``, ``Index``, ``SparseIndex``,
``BoTIndex``), with ``search`` served by hand-written sm_100a kernels behind the C ABI in
``include/vsearch_b200.h`` instead of ``torch.matmul`` + ``topk`` (upstream index.py:88-94).

Differences that are deliberate:
  * ties are ranked (score desc, id asc) -- upstream's order inside ties is arbitrary;
  * the dense ``.pt`` loader and the ``low_memory`` text store work (both are broken upstream);
  * an index searches only on a CUDA device: there is no CPU fallback, ``search`` on a CPU-resident
    index raises.
``.vector`` stays a torch tensor (CSR or strided) for drop-in use (``save``, ``str``); the engine
keeps its own compact device copy and never reads ``.vector`` during a search.
"""
from __future__ import annotations

import ctypes
import glob
import json
import logging
from enum import Enum
from typing import NamedTuple, Optional

import numpy as np
import torch

from . import _native as nat

logger = logging.getLogger(__name__)


class SearchResults(NamedTuple):
    """(ids, scores): int64 ``[B, k]`` ids and ``[B, k]`` scores in the index dtype, on the index
    device, ranked score-descending (upstream index.py:16-18, README.md:160-164)."""
    ids: torch.Tensor
    scores: torch.Tensor


class IndexType(Enum):
    DENSE = "dense"
    SPARSE = "sparse"
    BAG_OF_TOKEN = "bag_of_token"


_TORCH2VS = {torch.float32: nat.VS_F32, torch.float16: nat.VS_F16, torch.bfloat16: nat.VS_BF16,
             torch.int32: nat.VS_I32, torch.int64: nat.VS_I64}


def _is_cuda(device) -> bool:
    return torch.device(device).type == "cuda"


def _stream_ptr(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(torch.device(device)).cuda_stream)


class _Engine:
    """Owns one ``vs_index`` handle plus the reusable search workspace."""

    def __init__(self, handle: ctypes.c_void_p, device: torch.device):
        self.handle = handle
        self.device = device
        self._ws = None
        info = [ctypes.c_int64() for _ in range(3)]
        kind, sdt = ctypes.c_int(), ctypes.c_int()
        dbytes, sbytes = ctypes.c_int64(), ctypes.c_int64()
        nat.check(nat.LIB.vs_index_info(handle, info[0], info[1], info[2], kind, sdt, dbytes, sbytes))
        self.n_rows, self.n_cols, self.nnz = (int(x.value) for x in info)
        self.kind, self.store_dtype = kind.value, sdt.value
        self.device_bytes, self.stream_bytes = int(dbytes.value), int(sbytes.value)

    @classmethod
    def from_csr(cls, crow, col, val, shape, device, store_dtype=None) -> "_Engine":
        """crow/col(/val) torch tensors on any device; val None => binary bag-of-token."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("vsearch_b200 searches on CUDA devices only (no CPU fallback)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        crow, col = crow.contiguous(), col.contiguous()
        if crow.dtype not in (torch.int32, torch.int64):
            crow = crow.to(torch.int64)
        if col.dtype not in (torch.int32, torch.int64):
            col = col.to(torch.int64)
        if val is not None:
            val = val.contiguous()
            if val.dtype not in (torch.float32, torch.float16, torch.bfloat16):
                val = val.to(torch.float32)
            if store_dtype is None:
                store_dtype = val.dtype
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            rc = nat.LIB.vs_index_create_csr(
                device.index, int(shape[0]), int(shape[1]), int(col.numel()),
                crow.data_ptr(), _TORCH2VS[crow.dtype], col.data_ptr(), _TORCH2VS[col.dtype],
                None if val is None else val.data_ptr(), nat.VS_NONE if val is None else _TORCH2VS[val.dtype],
                nat.VS_NONE if val is None else _TORCH2VS[store_dtype], _stream_ptr(device), ctypes.byref(handle))
        nat.check(rc)
        return cls(handle, device)

    @classmethod
    def from_npz(cls, files, device, shift: int = 0, value_dtype=torch.float32, binary_if_ones: bool = False,
                 threads: int = 0):
        """Shard files -> device index without a host copy of the index (``vs_index_load_npz``).  Returns
        ``(engine, rows_per_file)``."""
        import os

        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("vsearch_b200 searches on CUDA devices only (no CPU fallback)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        paths = (ctypes.c_char_p * len(files))(*[os.fsencode(f) for f in files])
        rows = (ctypes.c_int64 * len(files))()
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            rc = nat.LIB.vs_index_load_npz(device.index, paths, len(files), int(shift), _TORCH2VS[value_dtype], int(binary_if_ones),
                                           int(threads), _stream_ptr(device), ctypes.byref(handle), rows)
        nat.check(rc)
        return cls(handle, device), [int(r) for r in rows]

    @classmethod
    def from_dense(cls, x: torch.Tensor, device, store_dtype) -> "_Engine":
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("vsearch_b200 searches on CUDA devices only (no CPU fallback)")
        if x.dim() != 2:
            raise ValueError("dense index must be [N, D]")
        if x.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            x = x.to(torch.float32)
        if x.stride(1) != 1:
            x = x.contiguous()
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            rc = nat.LIB.vs_index_create_dense(device.index, x.shape[0], x.shape[1], x.data_ptr(), _TORCH2VS[x.dtype],
                                               x.stride(0), _TORCH2VS[store_dtype], _stream_ptr(device),
                                               ctypes.byref(handle))
        nat.check(rc)
        return cls(handle, device)

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h and nat is not None and getattr(nat, "LIB", None) is not None:  # module may be gone at interpreter exit
            nat.LIB.vs_index_destroy(h)   # restores the caller's current device itself

    def workspace(self, B: int, k: int) -> torch.Tensor:
        need = int(nat.LIB.vs_search_workspace_bytes(self.handle, B, k))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def _sparse_q(self, q: torch.Tensor):
        """A sparse ``[B, V]`` query tensor (CSR or COO layout) -> (crow, col int32, val fp32) on the index device."""
        if q.dim() != 2 or q.shape[1] != self.n_cols:
            raise RuntimeError(f"query shape {tuple(q.shape)} does not match index width {self.n_cols}")
        q = q.to(self.device)
        if q.layout != torch.sparse_csr:
            q = q.coalesce().to_sparse_csr() if q.layout == torch.sparse_coo else q.to_sparse_csr()
        crow = q.crow_indices()
        if crow.dtype not in (torch.int32, torch.int64):
            crow = crow.to(torch.int64)
        return crow.contiguous(), q.col_indices().to(torch.int32).contiguous(), q.values().to(torch.float32).contiguous()

    def search_sparse(self, crow: torch.Tensor, col: torch.Tensor, val: torch.Tensor, k: int, mode: str = "auto",
                      score_round: int = nat.VS_F32, id_offset: int = 0, keys_only: bool = False):
        """Search with CSR-style (token, weight) query lists (``vs_search_sparse``): ``crow`` ``[B+1]`` int32/int64 offsets,
        ``col`` int32 tokens, ``val`` fp32 weights, on this device (or all on the host)."""
        B = crow.numel() - 1
        k = int(k)
        with torch.cuda.device(self.device):
            ws = self.workspace(B, max(k, 1))
            st = _stream_ptr(self.device)
            keys = ids = scores = None
            if keys_only:
                keys = torch.empty((B, max(k, 0)), dtype=torch.int64, device=self.device)
            else:
                ids = torch.empty((B, max(k, 0)), dtype=torch.int64, device=self.device)
                scores = torch.empty((B, max(k, 0)), dtype=torch.float32, device=self.device)
            rc = nat.LIB.vs_search_sparse(self.handle, crow.data_ptr(), _TORCH2VS[crow.dtype], col.data_ptr(), val.data_ptr(), B, k,
                                          nat.MODES[mode], score_round, id_offset,
                                          None if ids is None else ids.data_ptr(), None if scores is None else scores.data_ptr(),
                                          None if keys is None else keys.data_ptr(), ws.data_ptr(), ws.numel(), st)
        nat.check(rc)
        return keys if keys_only else (ids, scores)

    def _prep_q(self, q: torch.Tensor):
        q = q.to(self.device, non_blocking=True)
        if q.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            q = q.to(torch.float32)
        if q.dim() != 2 or q.shape[1] != self.n_cols:
            raise RuntimeError(f"query shape {tuple(q.shape)} does not match index width {self.n_cols}")
        return q.contiguous()

    def search(self, q: torch.Tensor, k: int, mode: str = "auto", score_round: int = nat.VS_F32,
               id_offset: int = 0, keys_only: bool = False):
        if q.layout != torch.strided:   # sparse queries: (token, weight) lists, no dense [B, V] rows
            if self.kind == 0:
                raise TypeError("sparse queries need a SparseIndex / BoTIndex")
            return self.search_sparse(*self._sparse_q(q), k, mode=mode, score_round=score_round, id_offset=id_offset,
                                      keys_only=keys_only)
        q = self._prep_q(q)
        B = q.shape[0]
        k = int(k)
        with torch.cuda.device(self.device):
            ws = self.workspace(B, max(k, 1))
            st = _stream_ptr(self.device)
            if keys_only:
                keys = torch.empty((B, max(k, 0)), dtype=torch.int64, device=self.device)
                rc = nat.LIB.vs_search_keys(self.handle, q.data_ptr(), _TORCH2VS[q.dtype], B, q.stride(0), k,
                                            nat.MODES[mode], score_round, id_offset, keys.data_ptr(),
                                            ws.data_ptr(), ws.numel(), st)
                nat.check(rc)
                return keys
            ids = torch.empty((B, max(k, 0)), dtype=torch.int64, device=self.device)
            scores = torch.empty((B, max(k, 0)), dtype=torch.float32, device=self.device)
            rc = nat.LIB.vs_search(self.handle, q.data_ptr(), _TORCH2VS[q.dtype], B, q.stride(0), k,
                                   nat.MODES[mode], score_round, id_offset, ids.data_ptr(), scores.data_ptr(),
                                   ws.data_ptr(), ws.numel(), st)
        nat.check(rc)
        return ids, scores

    def search_dense_step(self, step: int, q: torch.Tensor, k: int, n_ranks: int, gathered: Optional[torch.Tensor],
                          keys_out: torch.Tensor, status: Optional[torch.Tensor], score_round: int = nat.VS_F32,
                          id_offset: int = 0) -> None:
        """One step of the row-sharded dense search (``vs_search_dense_step``); ``q`` prepared ``[B <= 4096, dim]`` on
        this device, ``gathered`` ``[n_ranks, B, k]`` int64 keys of all ranks (steps 1, 2), ``keys_out`` ``[B, k]``."""
        B = q.shape[0]
        with torch.cuda.device(self.device):
            ws = self.workspace(B, k)
            rc = nat.LIB.vs_search_dense_step(self.handle, step, q.data_ptr(), _TORCH2VS[q.dtype], B, q.stride(0), k, score_round,
                                              id_offset, n_ranks, None if gathered is None else gathered.data_ptr(),
                                              keys_out.data_ptr(), None if status is None else status.data_ptr(),
                                              ws.data_ptr(), ws.numel(), _stream_ptr(self.device))
        nat.check(rc)

    def scores(self, q: torch.Tensor, score_round: int = nat.VS_F32) -> torch.Tensor:
        """Diagnostic: the full [B, N] score matrix (upstream index.py:91)."""
        q = self._prep_q(q)
        B = q.shape[0]
        out = torch.empty((B, self.n_rows), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            ws = self.workspace(B, 1)
            rc = nat.LIB.vs_scores(self.handle, q.data_ptr(), _TORCH2VS[q.dtype], B, q.stride(0), score_round,
                                   out.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(self.device))
        nat.check(rc)
        return out

    def score_rows(self, q: torch.Tensor, ids: torch.Tensor, score_round: int = nat.VS_F32) -> torch.Tensor:
        """scores[b, j] = <q_b, row ids[b, j]> (rerank stage, upstream retriever.py:137-141)."""
        q = self._prep_q(q)
        ids = ids.to(self.device, torch.int64).contiguous()
        if ids.dim() != 2 or ids.shape[0] != q.shape[0]:
            raise ValueError("ids must be [B, k] with one row per query")
        B, k = ids.shape
        out = torch.empty((B, k), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            need = int(nat.LIB.vs_score_rows_workspace_bytes(self.handle, B))   # the whole batch is prepared at once
            if self._ws is None or self._ws.numel() < need:
                self._ws = None
                self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            ws = self._ws
            rc = nat.LIB.vs_score_rows(self.handle, q.data_ptr(), _TORCH2VS[q.dtype], B, q.stride(0), ids.data_ptr(), k,
                                       score_round, out.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(self.device))
        nat.check(rc)
        return out

    def export_csr(self):
        crow = torch.empty(self.n_rows + 1, dtype=torch.int64, device=self.device)
        col = torch.empty(self.nnz, dtype=torch.int64, device=self.device)
        val = torch.empty(self.nnz, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            nat.check(nat.LIB.vs_index_export_csr(self.handle, crow.data_ptr(), col.data_ptr(), val.data_ptr(),
                                                  _stream_ptr(self.device)))
        return crow, col, val

    def kernel_timer(self, reset: bool = True):
        """(summed device ms, launches) of the scan kernels launched since the last reset."""
        ms, n = ctypes.c_float(), ctypes.c_int()
        nat.check(nat.LIB.vs_kernel_timer(self.handle, int(reset), ms, n))
        return float(ms.value), int(n.value)


def dense_to_csr(x: torch.Tensor):
    """Non-zeros of a dense ``[n, V]`` CUDA tensor as ``(crow int64 [n+1], col int32, val fp32)`` on the same device
    (``vs_dense_to_csr``: count -> scan -> fill on the GPU; upstream uses ``Tensor.to_sparse_csr``, retriever.py:299-305)."""
    if x.device.type != "cuda":
        raise RuntimeError("vsearch_b200.dense_to_csr runs on CUDA tensors only (no CPU fallback)")
    if x.dim() != 2:
        raise ValueError("dense_to_csr needs a [n, V] matrix")
    if x.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        x = x.to(torch.float32)
    if x.stride(1) != 1:
        x = x.contiguous()
    n, v = x.shape
    dev = x.device
    crow = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    args = (dev.index, x.data_ptr(), _TORCH2VS[x.dtype], n, x.stride(0) if n else v, v)
    with torch.cuda.device(dev):
        if n:
            nat.check(nat.LIB.vs_dense_to_csr(*args, crow[1:].data_ptr(), None, None, _stream_ptr(dev)))
            torch.cumsum(crow[1:], 0, out=crow[1:])
        nnz = int(crow[-1]) if n else 0
        col = torch.empty(nnz, dtype=torch.int32, device=dev)
        val = torch.empty(nnz, dtype=torch.float32, device=dev)
        if nnz:
            nat.check(nat.LIB.vs_dense_to_csr(*args, crow.data_ptr(), col.data_ptr(), val.data_ptr(), _stream_ptr(dev)))
    return crow, col, val


def topk_sparsify(emb_dense: torch.Tensor, k: int, bow_ids: Optional[torch.Tensor] = None, shift: int = 0,
                  as_sparse: bool = False) -> torch.Tensor:
    """Keep the ``k`` largest activations of every row of ``emb_dense`` ``[B, V]`` (ties -> lower column), zero the
    rest; with ``bow_ids`` ``[B, L]`` the columns ``id - shift`` of the row's own tokens survive too.  Upstream
    ``utils/sparse.py:8-19`` (``topk_sparsify`` / ``build_topk_mask``) and the ``logical_or(bow_mask, topk_mask)`` of
    ``encoder/vdr.py:159-169``; returns a new fp32 tensor on the embedding's CUDA device -- dense like upstream's, or
    with ``as_sparse=True`` a ``torch.sparse_csr`` tensor of the survivors (the (token, weight) lists ``Index.search``
    hands to ``vs_search_sparse`` without a dense ``[B, V]`` detour)."""
    if emb_dense.device.type != "cuda":
        raise RuntimeError("vsearch_b200.topk_sparsify runs on CUDA tensors only (no CPU fallback)")
    one_d = emb_dense.dim() == 1
    out = (emb_dense.unsqueeze(0) if one_d else emb_dense).to(torch.float32).clone().contiguous()
    B, V = out.shape
    ids = None
    if bow_ids is not None:
        ids = bow_ids.to(out.device, torch.int32).reshape(B, -1).contiguous()
    with torch.cuda.device(out.device):
        nat.check(nat.LIB.vs_sparsify_topk(out.device.index, out.data_ptr(), B, out.stride(0), V, int(k),
                                           None if ids is None else ids.data_ptr(), 0 if ids is None else ids.shape[1],
                                           int(shift), _stream_ptr(out.device)))
    if as_sparse:
        crow, col, val = dense_to_csr(out)
        return torch.sparse_csr_tensor(crow, col.to(torch.int64), val, size=(B, V))
    return out[0] if one_d else out


def merge_keys(keys: torch.Tensor, k_out: int):
    """Merge gathered rank keys ``[P, B, k_in]`` (int64 bit patterns) into ``(ids, scores)`` ``[B, k_out]``."""
    if not keys.is_cuda:
        raise RuntimeError("vsearch_b200 merges on CUDA devices only (no CPU fallback)")
    keys = keys.contiguous()
    P, B, k_in = keys.shape
    ids = torch.empty((B, k_out), dtype=torch.int64, device=keys.device)
    scores = torch.empty((B, k_out), dtype=torch.float32, device=keys.device)
    with torch.cuda.device(keys.device):
        nat.check(nat.LIB.vs_merge_keys(keys.device.index, keys.data_ptr(), P, B * k_in, k_in, B, k_in, k_out,
                                        ids.data_ptr(), scores.data_ptr(), _stream_ptr(keys.device)))
    return ids, scores


class Index:
    """Dense index (``[N, D]`` strided ``.vector``) and base class (upstream index.py:25-126)."""

    index_type = IndexType.DENSE

    def __init__(self, index_file: Optional[str] = None, data_file: Optional[str] = None, fp16: bool = True,
                 device: str = "cpu", low_memory: bool = False):
        self.data = None
        self._vector = None
        self._engine = None
        self.low_memory = low_memory
        self.device = device
        self.search_mode = "auto"  # "auto" | "scan" | "inverted"
        self.init_index(index_file, fp16)
        self.load_data(data_file)

    # ---- the logical vector ---------------------------------------------------------------
    @property
    def vector(self):
        if self._vector is None and self._engine is not None and self._engine.kind != 0:
            crow, col, val = self._engine.export_csr()
            self._vector = torch.sparse_csr_tensor(crow, col, val.to(self._value_dtype()),
                                                   size=(self._engine.n_rows, self._engine.n_cols))
        return self._vector

    @vector.setter
    def vector(self, value):
        self._vector = value
        self._engine = None
        self._logical_dtype = None

    def _value_dtype(self):
        if self._vector is not None:
            return self._vector.dtype
        if getattr(self, "_logical_dtype", None) is not None:   # engine-only index: dtype the logical vector would have
            return self._logical_dtype
        sd = self._engine.store_dtype if self._engine is not None else nat.VS_F32
        return {nat.VS_F16: torch.float16, nat.VS_BF16: torch.bfloat16}.get(sd, torch.float32)

    # ---- loading ----------------------------------------------------------------------------
    def init_index(self, index_path: Optional[str], fp16: bool = True):
        """Dense shards: glob ``*.pt`` in sorted order, concatenate rows (intent of upstream index.py:36-44)."""
        if not index_path:
            return
        files = sorted(glob.glob(index_path))
        if not files:
            raise FileNotFoundError(f"no index files match {index_path!r}")
        logger.info("***** Loading %s Index from %d files *****", self.index_type.value, len(files))
        shards = [torch.load(f, map_location="cpu") for f in files]
        vec = torch.cat(shards, dim=0) if len(shards) > 1 else shards[0]
        self.vector = vec.to(torch.float16) if fp16 else vec
        self.move_to_device(self.device)

    def load_data(self, data_file: Optional[str]):
        if not data_file:
            return
        if not self.low_memory:
            with open(data_file, "r") as f:
                self.data = [json.loads(line) for line in f]
        else:
            self.offsets = self._calculate_offsets(data_file)
            self.data_file = data_file

    @staticmethod
    def _calculate_offsets(data_file: str):
        offsets, pos = [], 0
        with open(data_file, "rb") as f:
            for line in f:
                offsets.append(pos)
                pos += len(line)
        return offsets

    def get_sample(self, index: int):
        if not self.low_memory:
            return self.data[index]
        with open(self.data_file, "rb") as f:
            f.seek(self.offsets[index])
            return json.loads(f.readline().decode("utf-8"))

    # ---- device placement ----------------------------------------------------------------------
    def move_to_device(self, device: str):
        logger.info("Moving index to %s.", device)
        self.device = device
        if self._vector is None and self._engine is None:
            return
        if _is_cuda(device):
            if self._engine is None or self._engine.device != self._resolve(device):
                self._build_engine(device)
        else:
            if self._vector is None:
                self._vector = self.vector  # export before dropping the engine
            self._vector = self._vector.to(device)
            self._engine = None

    @staticmethod
    def _resolve(device) -> torch.device:
        d = torch.device(device)
        if d.type == "cuda" and d.index is None:
            d = torch.device("cuda", torch.cuda.current_device())
        return d

    def _build_engine(self, device):
        """Dense ``[N, D]`` vector -> bf16 / fp16 K-major device copy for the tcgen05 kernel (K4)."""
        dev = self._resolve(device)
        v = self._vector
        if v.layout != torch.strided:
            raise TypeError("the dense Index needs a strided [N, D] vector; use SparseIndex / BoTIndex for CSR")
        if v.dtype not in (torch.float16, torch.bfloat16, torch.float32):
            v = v.to(torch.float32)
        # bf16 / fp16 vectors are scored as they are on the tensor cores.  An fp32 vector (upstream's Index(fp16=False):
        # an fp32 GEMM, index.py:36-44, 88-94) keeps fp32 semantics: the tensor cores sweep a bf16 copy, every passage
        # within the bf16 error bound of the k-th score is re-scored exactly from its fp32 row (csrc/dense.cu).
        self._engine = _Engine.from_dense(v.to(dev), dev, v.dtype)

    def _require_engine(self) -> _Engine:
        if self._engine is None:
            if self._vector is None:
                raise RuntimeError("index is empty: nothing to search")
            if not _is_cuda(self.device):
                raise RuntimeError(
                    "vsearch_b200 has no CPU search path: call index.move_to_device('cuda') first")
            self._build_engine(self.device)
        return self._engine

    # ---- search -------------------------------------------------------------------------------------
    def _score_round(self) -> int:
        return {torch.float16: nat.VS_F16, torch.bfloat16: nat.VS_BF16}.get(self._value_dtype(), nat.VS_F32)

    def search(self, q_embs: torch.Tensor, k: int) -> SearchResults:
        """``scores = q @ vector.t(); scores.topk(k)`` (upstream index.py:88-94) without the [B, N] matrix."""
        eng = self._require_engine()
        one_d = q_embs.dim() == 1
        q = q_embs.unsqueeze(0) if one_d else q_embs   # strided, or a sparse [B, V] tensor of (token, weight) lists
        ids, scores = eng.search(q, k, mode=self.search_mode, score_round=self._score_round())
        scores = scores.to(self._value_dtype())
        if one_d:
            ids, scores = ids[0], scores[0]
        return SearchResults(ids, scores)

    def last_mode(self) -> str:
        """Kernel family the last search on this index used ("scan" | "inverted")."""
        m = ctypes.c_int()
        nat.check(nat.LIB.vs_index_last_mode(self._require_engine().handle, m))
        return "inverted" if m.value == nat.VS_MODE_INVERTED else "scan"

    def score_rows(self, q_embs: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
        """``scores[b, j] = <q_b, vector[ids[b, j]]>`` in the index dtype: the re-scoring half of upstream's
        ``retrieve(rerank=True)`` (retriever.py:137-141) for candidates whose vectors live in this index."""
        eng = self._require_engine()
        q = q_embs.unsqueeze(0) if q_embs.dim() == 1 else q_embs
        ids2 = ids.reshape(q.shape[0], -1)
        return eng.score_rows(q, ids2, score_round=self._score_round()).to(self._value_dtype()).reshape(ids.shape)

    def search_keys(self, q_embs: torch.Tensor, k: int, id_offset: int = 0) -> torch.Tensor:
        """Packed rank keys ``[B, k]`` of this shard with global ids (row-sharded path, sharded.py)."""
        eng = self._require_engine()
        return eng.search(q_embs, k, mode=self.search_mode, score_round=self._score_round(),
                          id_offset=id_offset, keys_only=True)

    # ---- persistence / introspection ------------------------------------------------------------------
    def save(self, path):
        """Dense: one ``.pt`` tensor (upstream index.py:96-109)."""
        try:
            torch.save(self.vector.cpu(), path)
            logger.info("Index successfully saved to %s", path)
        except Exception as e:  # noqa: BLE001 - upstream logs then re-raises
            logger.error("Failed to save index to %s: %s", path, e)
            raise

    def __len__(self):
        return len(self.data) if self.data else 0

    def __repr__(self):
        return repr(self.vector)

    def _shape(self):
        if self._vector is not None:
            return self._vector.shape
        return torch.Size((self._engine.n_rows, self._engine.n_cols))

    def __str__(self):
        layout = self._vector.layout if self._vector is not None else torch.sparse_csr
        return (
            f"Index Type        : {type(self).__name__}\n"
            f"Vector Shape      : {self._shape()}\n"
            f"Vector Dtype      : {self._value_dtype()}\n"
            f"Vector Layout     : {layout}\n"
            f"Number of Texts   : {len(self.data) if self.data else 0}\n"
            f"Device            : {self.device}\n"
        )


class SparseIndex(Index):
    """``[N, V]`` sparse CSR index (upstream index.py:128-202)."""

    index_type = IndexType.SPARSE

    def __init__(self, index_file: Optional[str] = None, data_file: Optional[str] = None, fp16: bool = True,
                 device: str = "cpu", low_memory: bool = False, shift: int = 0):
        self.shift = shift
        super().__init__(index_file, data_file, fp16, device, low_memory)

    def _scipy_csr_to_torch_csr(self, mat) -> torch.Tensor:
        """scipy CSR -> torch CSR on ``self.device`` (upstream index.py:144-161).  With a CUDA device the
        torch tensor stays on the host and the engine gets the compact device copy."""
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t = torch.sparse_csr_tensor(torch.from_numpy(np.ascontiguousarray(mat.indptr)),
                                        torch.from_numpy(np.ascontiguousarray(mat.indices)),
                                        torch.from_numpy(np.ascontiguousarray(mat.data)), size=mat.shape)
        return t

    def init_index(self, index_file: Optional[str], fp16: bool = True):
        """glob -> sorted -> load_npz -> ``[:, shift:]`` -> vstack -> optional fp16 (upstream index.py:163-179)."""
        if not index_file:
            return
        from .npz_io import load_csr_shards, load_csr_shards_native

        # a glob pattern like upstream's, or an explicit list of shard files in row order (a rank of the row-sharded layout)
        files = sorted(glob.glob(index_file)) if isinstance(index_file, str) else [str(f) for f in index_file]
        if not files:
            raise FileNotFoundError(f"no index files match {index_file!r}")
        logger.info("***** Loading %s Index from %d files *****", self.index_type.value, len(files))
        if _is_cuda(self.device):
            # straight to the device (csrc/npz.cu): parallel inflate -> pinned staging -> the device CSR arrays -> index
            # build, with the column shift applied on the GPU; no scipy matrices, no host copy of the index.  `.vector`
            # is exported from the engine on demand (save, repr).
            self._vector = None
            self._engine, self.shard_rows = _Engine.from_npz(files, self._resolve(self.device), shift=max(self.shift, 0),
                                                             value_dtype=torch.float16 if fp16 else torch.float32,
                                                             binary_if_ones=self.index_type == IndexType.BAG_OF_TOKEN)
            self._logical_dtype = torch.float16 if fp16 else torch.float32
            return
        if self.shift <= 0:
            # native reader (csrc/npz.cu): all members of all shards inflated in parallel into the final arrays,
            # int64 -> int32 narrowing and astype(float16) on the fly
            indptr, indices, data, shape = load_csr_shards_native(files, fp16=fp16)
        else:  # the column slice changes the row pointers: numpy path
            indptr, indices, data, shape = load_csr_shards(files, self.shift)
            if fp16:
                data = data.astype(np.float16)
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            self.vector = torch.sparse_csr_tensor(torch.from_numpy(indptr), torch.from_numpy(indices),
                                                  torch.from_numpy(data), size=shape)
        self.move_to_device(self.device)

    def _csr_parts(self):
        v = self._vector
        if v.layout != torch.sparse_csr:
            v = v.to_sparse_csr()
        return v.crow_indices(), v.col_indices(), v.values(), v.shape

    def _device_csr(self, dev):
        """(crow, col, val, shape) on ``dev``; a strided ``.vector`` is sparsified there (``vs_dense_to_csr``)."""
        v = self._vector
        if v.layout == torch.strided:
            crow, col, val = dense_to_csr(v.to(dev))
            return crow, col, val.to(v.dtype if v.dtype in (torch.float16, torch.bfloat16) else torch.float32), v.shape
        crow, col, val, shape = self._csr_parts()
        return crow.to(dev), col.to(dev), val.to(dev), shape

    def _drop_dense_vector(self):
        """A strided ``.vector`` was sparsified on the GPU: keep only the engine; ``.vector`` exports CSR on demand
        (upstream's ``.vector`` of a sparse index is a CSR tensor, retriever.py:304)."""
        if self._vector is not None and self._vector.layout == torch.strided:
            dt = self._vector.dtype
            self._vector = None
            self._logical_dtype = dt

    def _build_engine(self, device):
        dev = self._resolve(device)
        crow, col, val, shape = self._device_csr(dev)
        self._engine = _Engine.from_csr(crow, col, val, shape, dev)
        self._drop_dense_vector()

    def save(self, path):
        """scipy-loadable ``.npz`` (upstream index.py:181-202)."""
        try:
            from .npz_io import save_csr_npz_native

            crow, col, val, shape = self._csr_parts() if self._vector is not None else (*self._engine.export_csr(), self._shape())
            if val.dtype == torch.bfloat16:
                val = val.to(torch.float32)  # numpy has no bfloat16
            # same members / order / dtypes as scipy.sparse.save_npz, deflated by a thread pool (csrc/npz.cu)
            save_csr_npz_native(path, crow.cpu().numpy(), col.cpu().numpy(), val.cpu().numpy(), tuple(shape))
            logger.info("Index successfully saved to %s", path)
        except Exception as e:  # noqa: BLE001
            logger.error("Failed to save index to %s: %s", path, e)
            raise


class BoTIndex(SparseIndex):
    """Binary bag-of-token index (upstream index.py:205-218): all stored values are 1, so the device
    copy keeps column ids only (2 bytes per entry)."""

    index_type = IndexType.BAG_OF_TOKEN

    def _build_engine(self, device):
        dev = self._resolve(device)
        crow, col, val, shape = self._device_csr(dev)
        binary = bool((val == 1).all().item()) if val.numel() else True
        self._engine = _Engine.from_csr(crow, col, None if binary else val, shape, dev)
        self._drop_dense_vector()

    @classmethod
    def from_token_csr(cls, crow: torch.Tensor, col: torch.Tensor, shape, device="cuda", dtype=torch.float32):
        """Build straight from (crow, col) device arrays -- no values, no torch CSR copy (large synthetic /
        pre-tokenised corpora).  ``dtype`` is the logical value dtype reported by ``.vector`` / scores."""
        self = cls(device=device)
        self._engine = _Engine.from_csr(crow, col, None, shape, self._resolve(device))
        self._logical_dtype = dtype
        return self

    @classmethod
    def from_token_ids(cls, token_ids: torch.Tensor, lengths: Optional[torch.Tensor] = None, vocab_size: int = 30522,
                       num_shift: int = 999, max_token: Optional[int] = None, device="cuda", dtype=torch.float16):
        """Bag-of-token index straight from tokenizer output (upstream ``Retriever._build_bot_vectors``,
        retriever.py:208-253): ``token_ids`` ``[N, max_len]`` int32/int64 (padded; ``lengths`` = valid ids per row),
        row = the distinct ids of the passage (the first ``max_token`` distinct ones when given), ids below
        ``num_shift`` dropped, the rest renumbered ``id - num_shift``.  Built on the GPU (``vs_bot_from_tokens``):
        no dense ``[batch, vocab]`` matrix, no COO concatenation."""
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("vsearch_b200 builds and searches on CUDA devices only (no CPU fallback)")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        ids = token_ids.to(dev)
        if ids.dtype not in (torch.int32, torch.int64):
            ids = ids.to(torch.int64)
        ids = ids.contiguous()
        if ids.dim() != 2:
            raise ValueError("token_ids must be [N, max_len]")
        n, ld = ids.shape
        lens = None if lengths is None else lengths.to(dev, torch.int32).contiguous()
        crow = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        _sp = _stream_ptr
        args = (dev.index, ids.data_ptr(), _TORCH2VS[ids.dtype], n, ld, None if lens is None else lens.data_ptr(),
                int(vocab_size), int(num_shift), int(max_token or 0))
        with torch.cuda.device(dev):
            nat.check(nat.LIB.vs_bot_from_tokens(*args, crow[1:].data_ptr() if n else None, None, _sp(dev)))
            torch.cumsum(crow[1:], 0, out=crow[1:])
            col = torch.empty(int(crow[-1]) if n else 0, dtype=torch.int32, device=dev)
            if n:
                nat.check(nat.LIB.vs_bot_from_tokens(*args, crow.data_ptr(), col.data_ptr(), _sp(dev)))
        return cls.from_token_csr(crow, col, (n, int(vocab_size) - int(num_shift)), device=dev, dtype=dtype)

