"""ctypes binding of oracle/oracle.c (TEST INFRASTRUCTURE ONLY; see oracle.c header)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build() -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _prep(crow, col, val, q):
    crow = np.ascontiguousarray(crow, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int64)
    val = None if val is None else np.ascontiguousarray(val, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    return crow, col, val, q


def csr_scores(crow, col, val, shape, q):
    crow, col, val, q = _prep(crow, col, val, q)
    N, V = int(shape[0]), int(shape[1])
    B = q.shape[0]
    out = np.empty((B, N), dtype=np.float32)
    rc = lib().oracle_csr_scores(_p(crow), _p(col), _p(val), ctypes.c_int64(N), ctypes.c_int64(V),
                                 _p(q), ctypes.c_int64(B), _p(out))
    assert rc == 0
    return out


def topk(scores, k):
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    B, N = scores.shape
    ids = np.empty((B, k), dtype=np.int64)
    sc = np.empty((B, k), dtype=np.float32)
    rc = lib().oracle_topk(_p(scores), ctypes.c_int64(B), ctypes.c_int64(N), ctypes.c_int(k), _p(ids), _p(sc))
    if rc == -1:
        raise RuntimeError("selected index k out of range")
    assert rc == 0
    return ids, sc


def csr_search(crow, col, val, shape, q, k):
    crow, col, val, q = _prep(crow, col, val, q)
    N, V = int(shape[0]), int(shape[1])
    B = q.shape[0]
    ids = np.empty((B, k), dtype=np.int64)
    sc = np.empty((B, k), dtype=np.float32)
    rc = lib().oracle_csr_search(_p(crow), _p(col), _p(val), ctypes.c_int64(N), ctypes.c_int64(V),
                                 _p(q), ctypes.c_int64(B), ctypes.c_int(k), _p(ids), _p(sc))
    if rc == -1:
        raise RuntimeError("selected index k out of range")
    assert rc == 0
    return ids, sc


def dense_scores(x, q):
    x = np.ascontiguousarray(x, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    N, D = x.shape
    B = q.shape[0]
    out = np.empty((B, N), dtype=np.float32)
    rc = lib().oracle_dense_scores(_p(x), ctypes.c_int64(N), ctypes.c_int64(D), _p(q), ctypes.c_int64(B), _p(out))
    assert rc == 0
    return out


def num_threads() -> int:
    return int(lib().oracle_num_threads())
