"""``.npz`` index files, byte-compatible with ``scipy.sparse.save_npz`` / ``load_npz``.

Layout (measured on files written by upstream ``SparseIndex.save``, index.py:181-202): a deflated zip
with members ``indices.npy``, ``indptr.npy``, ``format.npy`` (``|S3`` ``b'csr'``), ``shape.npy``
(int64[2]), ``data.npy``, ``_is_array.npy`` (bool).  ``indices``/``indptr`` are int32 or int64, ``data``
float32 or float16.

The members are read with ``numpy.load`` rather than ``scipy.sparse.load_npz`` so that float16 ``data``
works (current scipy refuses float16 in the ``[:, shift:]`` slice the upstream loader performs,
index.py:174) and so that no scipy matrix copy of a 21M-row index is ever made.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def read_csr_npz(path: str):
    with np.load(path, allow_pickle=False) as z:
        fmt = z["format"].item()
        fmt = fmt.decode("ascii") if isinstance(fmt, bytes) else str(fmt)
        if fmt != "csr":
            raise ValueError(f"{path}: expected a CSR .npz, found format {fmt!r}")
        shape = tuple(int(x) for x in z["shape"])
        return z["indptr"], z["indices"], z["data"], shape


def _shift_columns(indptr, indices, data, shape, shift: int):
    """Equivalent of ``mat[:, shift:]`` on a CSR matrix: drop columns < shift, renumber the rest."""
    if shift <= 0:
        return indptr, indices, data, shape
    keep = indices >= shift
    kept_before = np.concatenate(([0], np.cumsum(keep, dtype=np.int64)))
    new_indptr = kept_before[indptr.astype(np.int64)]
    return (new_indptr.astype(indptr.dtype), (indices[keep] - shift).astype(indices.dtype), data[keep],
            (shape[0], max(shape[1] - shift, 0)))


def load_csr_shards(files: Sequence[str], shift: int = 0) -> Tuple[np.ndarray, np.ndarray, np.ndarray, Tuple[int, int]]:
    """Row-concatenate shard files in the order given (the caller sorts them lexicographically, as
    upstream does: ``index10`` before ``index2``, index.py:172-175)."""
    ptrs: List[np.ndarray] = []
    idxs: List[np.ndarray] = []
    vals: List[np.ndarray] = []
    n_rows, n_cols, nnz = 0, None, 0
    for f in files:
        indptr, indices, data, shape = _shift_columns(*read_csr_npz(f), shift)
        if n_cols is None:
            n_cols = shape[1]
        elif shape[1] != n_cols:
            raise ValueError(f"{f}: column count {shape[1]} differs from previous shards ({n_cols})")
        ptrs.append(indptr.astype(np.int64)[(1 if ptrs else 0):] + nnz)
        idxs.append(indices)
        vals.append(data)
        n_rows += shape[0]
        nnz += int(indptr[-1])
    idx_dtype = np.int32 if max(nnz, n_cols or 0) < 2**31 - 1 else np.int64
    indptr = np.concatenate(ptrs).astype(idx_dtype)
    indices = np.concatenate(idxs).astype(idx_dtype)
    data = np.concatenate(vals)
    return indptr, indices, data, (n_rows, int(n_cols or 0))


def save_csr_npz(path: str, indptr: np.ndarray, indices: np.ndarray, data: np.ndarray, shape) -> None:
    """Write what ``scipy.sparse.save_npz(path, csr_array(...))`` writes (compressed)."""
    np.savez_compressed(
        path,
        indices=np.ascontiguousarray(indices),
        indptr=np.ascontiguousarray(indptr),
        format=np.array(b"csr"),
        shape=np.asarray(shape, dtype=np.int64),
        data=np.ascontiguousarray(data),
        _is_array=np.array(True),
    )
