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

import ctypes
import os
from concurrent.futures import ThreadPoolExecutor
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


# ---- native loader (csrc/npz.cu through the C ABI) ---------------------------------------------------------------
_NP2VS = {np.dtype(np.int32): 3, np.dtype(np.int64): 4, np.dtype(np.float32): 0, np.dtype(np.float16): 1}
_VS2NP = {0: np.float32, 1: np.float16, 3: np.int32, 4: np.int64, 5: np.uint16, 6: np.uint32, 8: np.float64}


class _NativeShard:
    """One open shard file: member table from the zip central directory + .npy headers (no data read yet)."""

    def __init__(self, path: str):
        from . import _native as nat

        self.nat, self.path = nat, path
        self.h = ctypes.c_void_p()
        nat.check(nat.LIB.vs_npz_open(os.fsencode(path), ctypes.byref(self.h)))

    def info(self, name: str):
        dt, nd, n = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64()
        shape = (ctypes.c_int64 * 4)()
        self.nat.check(self.nat.LIB.vs_npz_member_info(self.h, name.encode(), dt, nd, shape, n))
        return dt.value, tuple(shape[i] for i in range(nd.value)), int(n.value)

    def read_into(self, name: str, out: np.ndarray, skip: int = 0, add: int = 0) -> None:
        assert out.flags.c_contiguous
        self.nat.check(self.nat.LIB.vs_npz_read(self.h, name.encode(), out.ctypes.data_as(ctypes.c_void_p),
                                                _NP2VS[out.dtype], skip, out.size, add))

    def small(self, name: str) -> np.ndarray:
        dt, shape, n = self.info(name)
        out = np.empty(n, dtype=_VS2NP[dt])
        self.read_into(name, out)
        return out.reshape(shape)

    def close(self):
        if self.h:
            self.nat.LIB.vs_npz_close(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass


def shard_row_counts(files: Sequence[str]) -> List[int]:
    """Rows of every shard file, from the .npy header of its ``indptr`` member (nothing is inflated beyond the header)."""
    out = []
    for f in files:
        sh = _NativeShard(f)
        try:
            out.append(sh.info("indptr")[2] - 1)
        finally:
            sh.close()
    return out


def load_csr_shards_native(files: Sequence[str], fp16: bool = False, threads: int | None = None):
    """Row-concatenation of shard files like :func:`load_csr_shards` (``shift == 0``), but every big member of every
    shard is inflated by its own thread straight into its slice of the final arrays (ctypes releases the GIL), with
    the int64 -> int32 narrowing, the row-pointer offset of the shard and the optional ``astype(float16)`` of the
    reference loader (index.py:176) applied on the fly.  Peak host memory = the final arrays.
    Returns (indptr, indices, data, shape) as numpy arrays."""
    shards = [_NativeShard(f) for f in files]
    try:
        metas = []
        n_rows, n_cols, nnz = 0, None, 0
        data_np = None
        for sh, f in zip(shards, files):
            with np.load(f, allow_pickle=False) as z:   # the three tiny members: numpy is fine
                fmt = z["format"].item()
                shape = tuple(int(x) for x in z["shape"])
            fmt = fmt.decode("ascii") if isinstance(fmt, bytes) else str(fmt)
            if fmt != "csr":
                raise ValueError(f"{f}: expected a CSR .npz, found format {fmt!r}")
            (_, _, n_ptr), (_, _, n_idx), (ddt, _, n_dat) = sh.info("indptr"), sh.info("indices"), sh.info("data")
            if n_ptr != shape[0] + 1 or n_idx != n_dat:
                raise ValueError(f"{f}: inconsistent CSR members")
            if n_cols is None:
                n_cols = shape[1]
            elif shape[1] != n_cols:
                raise ValueError(f"{f}: column count {shape[1]} differs from previous shards ({n_cols})")
            if ddt not in (0, 1, 8):
                raise ValueError(f"{f}: data must be float64, float32 or float16")
            # float64 (scipy's default dtype) is narrowed to float32 while inflating: the engine searches f32 / f16 / bf16
            ddt_np = np.float32 if ddt == 8 else _VS2NP[ddt]
            data_np = ddt_np if data_np is None else np.result_type(data_np, ddt_np)
            metas.append((n_rows, nnz, shape[0], n_idx))
            n_rows += shape[0]
            nnz += n_idx
        idx_dtype = np.int32 if max(nnz, n_cols or 0) < 2**31 - 1 else np.int64
        indptr = np.empty(n_rows + 1, dtype=idx_dtype)
        indices = np.empty(nnz, dtype=idx_dtype)
        data = np.empty(nnz, dtype=np.float16 if fp16 else (data_np or np.float32))
        indptr[0] = 0
        jobs = []
        for sh, (r0, z0, rows, cnt) in zip(shards, metas):
            jobs.append((sh, "indices", indices[z0:z0 + cnt], 0, 0))
            jobs.append((sh, "data", data[z0:z0 + cnt], 0, 0))
            jobs.append((sh, "indptr", indptr[r0 + 1:r0 + 1 + rows], 1, z0))   # drop each shard's leading 0, add its offset
        jobs.sort(key=lambda j: -j[2].size)
        with ThreadPoolExecutor(max_workers=threads or min(len(jobs), os.cpu_count() or 1)) as pool:
            list(pool.map(lambda j: j[0].read_into(j[1], j[2], skip=j[3], add=j[4]), jobs))
        return indptr, indices, data, (n_rows, int(n_cols or 0))
    finally:
        for sh in shards:
            sh.close()


def save_csr_npz_native(path: str, indptr: np.ndarray, indices: np.ndarray, data: np.ndarray, shape,
                        threads: int | None = None, level: int = 6) -> None:
    """Same file as :func:`save_csr_npz` (members, order, dtypes, deflate), written by ``vs_npz_write``: every member
    is deflated by a thread pool (independent 4 MB blocks concatenated into one valid stream), so saving a 21M-row
    index is no longer one zlib thread."""
    import io

    from . import _native as nat

    path = os.fspath(path)
    if not path.endswith(".npz"):
        path += ".npz"   # numpy / scipy append the suffix too
    arrays = [("indices", np.ascontiguousarray(indices)), ("indptr", np.ascontiguousarray(indptr)),
              ("format", np.array(b"csr")), ("shape", np.asarray(shape, dtype=np.int64)),
              ("data", np.ascontiguousarray(data)), ("_is_array", np.array(True))]
    keep, members = [], (nat.NpzMemberIn * len(arrays))()
    for i, (name, arr) in enumerate(arrays):
        buf = io.BytesIO()
        np.lib.format.write_array_header_1_0(buf, np.lib.format.header_data_from_array_1_0(arr))
        hdr = buf.getvalue()
        keep.append((hdr, arr))
        members[i].name = name.encode()
        members[i].header = ctypes.cast(ctypes.c_char_p(hdr), ctypes.c_void_p)
        members[i].header_bytes = len(hdr)
        members[i].data = arr.ctypes.data_as(ctypes.c_void_p) if arr.nbytes else None
        members[i].data_bytes = arr.nbytes
    nat.check(nat.LIB.vs_npz_write(os.fsencode(path), members, len(arrays), level, threads or min(16, os.cpu_count() or 1)))


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
