"""ctypes binding of the C ABI in include/vsearch_b200.h.

There is no CPU fallback and no alternative backend: if the CUDA library has not been built
(``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C vsearch_b200/csrc``) importing
this module raises, and so does every search.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# VSEARCH_B200_LIB: an experiment build of the same library (csrc/Makefile: BUILD= OUT= EXTRA=); never a different backend
LIB_PATH = os.environ.get("VSEARCH_B200_LIB") or os.path.join(_HERE, "lib", "libvsearch_b200.so")

VS_OK, VS_ERR_INVALID, VS_ERR_UNSUPPORTED, VS_ERR_CUDA, VS_ERR_NOMEM = 0, 1, 2, 3, 4
VS_F32, VS_F16, VS_BF16, VS_I32, VS_I64, VS_U16, VS_U32, VS_NONE, VS_F64 = range(9)
VS_MODE_AUTO, VS_MODE_SCAN, VS_MODE_INVERTED = 0, 1, 2
VS_MAX_K = 2048
MODES = {"auto": VS_MODE_AUTO, "scan": VS_MODE_SCAN, "inverted": VS_MODE_INVERTED}

# every symbol include/vsearch_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "vs_last_error", "vs_abi_version", "vs_index_create_csr", "vs_index_create_dense", "vs_index_destroy",
    "vs_index_info", "vs_index_export_csr", "vs_search_workspace_bytes", "vs_search", "vs_search_keys",
    "vs_scores", "vs_merge_keys", "vs_kernel_timer", "vs_index_last_mode",
    "vs_npz_open", "vs_npz_close", "vs_npz_member_info", "vs_npz_read", "vs_bot_from_tokens", "vs_score_rows", "vs_npz_write", "vs_sparsify_topk",
    "vs_debug_scan_profile", "vs_debug_gather_wavefronts", "vs_search_sparse", "vs_score_rows_workspace_bytes",
    "vs_dense_to_csr", "vs_index_load_npz", "vs_search_dense_step",
]


class NpzMemberIn(ctypes.Structure):
    _fields_ = [("name", c_char_p), ("header", c_void_p), ("header_bytes", c_int64), ("data", c_void_p),
                ("data_bytes", c_int64)]


class NativeLibraryMissing(ImportError):
    pass


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: the sm_100a CUDA library is not built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)."
        )
    lib = ctypes.CDLL(LIB_PATH)
    lib.vs_last_error.restype = c_char_p
    lib.vs_abi_version.restype = c_int
    lib.vs_index_create_csr.argtypes = [c_int, c_int64, c_int64, c_int64, c_void_p, c_int, c_void_p, c_int,
                                        c_void_p, c_int, c_int, c_void_p, POINTER(c_void_p)]
    lib.vs_index_create_dense.argtypes = [c_int, c_int64, c_int64, c_void_p, c_int, c_int64, c_int, c_void_p,
                                          POINTER(c_void_p)]
    lib.vs_index_destroy.argtypes = [c_void_p]
    lib.vs_index_info.argtypes = [c_void_p, POINTER(c_int64), POINTER(c_int64), POINTER(c_int64), POINTER(c_int),
                                  POINTER(c_int), POINTER(c_int64), POINTER(c_int64)]
    lib.vs_index_export_csr.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.vs_search_workspace_bytes.argtypes = [c_void_p, c_int64, c_int]
    lib.vs_search_workspace_bytes.restype = c_size_t
    lib.vs_search.argtypes = [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_int, c_int64,
                              c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.vs_search_keys.argtypes = [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_int, c_int64,
                                   c_void_p, c_void_p, c_size_t, c_void_p]
    lib.vs_search_dense_step.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_int64, c_int,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.vs_search_sparse.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int64,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.vs_score_rows_workspace_bytes.argtypes = [c_void_p, c_int64]
    lib.vs_score_rows_workspace_bytes.restype = c_size_t
    lib.vs_scores.argtypes = [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int, c_void_p, c_void_p, c_size_t,
                              c_void_p]
    lib.vs_merge_keys.argtypes = [c_int, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_int, c_void_p,
                                  c_void_p, c_void_p]
    lib.vs_index_last_mode.argtypes = [c_void_p, POINTER(c_int)]
    lib.vs_debug_scan_profile.argtypes = [c_void_p, c_void_p]
    lib.vs_debug_gather_wavefronts.argtypes = [c_void_p, c_void_p, c_void_p]
    lib.vs_kernel_timer.argtypes = [c_void_p, c_int, POINTER(c_float), POINTER(c_int)]
    lib.vs_score_rows.argtypes = [c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                  c_size_t, c_void_p]
    lib.vs_bot_from_tokens.argtypes = [c_int, c_void_p, c_int, c_int64, c_int64, c_void_p, c_int, c_int, c_int, c_void_p,
                                       c_void_p, c_void_p]
    lib.vs_sparsify_topk.argtypes = [c_int, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_int, c_int, c_void_p]
    lib.vs_dense_to_csr.argtypes = [c_int, c_void_p, c_int, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.vs_index_load_npz.argtypes = [c_int, POINTER(c_char_p), c_int, c_int, c_int, c_int, c_int, c_void_p, POINTER(c_void_p),
                                      POINTER(c_int64)]
    lib.vs_npz_write.argtypes = [c_char_p, c_void_p, c_int, c_int, c_int]
    lib.vs_npz_open.argtypes = [c_char_p, POINTER(c_void_p)]
    lib.vs_npz_close.argtypes = [c_void_p]
    lib.vs_npz_member_info.argtypes = [c_void_p, c_char_p, POINTER(c_int), POINTER(c_int), POINTER(c_int64), POINTER(c_int64)]
    lib.vs_npz_read.argtypes = [c_void_p, c_char_p, c_void_p, c_int, c_int64, c_int64, c_int64]
    for name in SYMBOLS:
        getattr(lib, name)  # AttributeError here = header / library mismatch
    if lib.vs_abi_version() != 2:
        raise ImportError("libvsearch_b200.so ABI version mismatch")
    return lib


LIB = _load()


def last_error() -> str:
    return (LIB.vs_last_error() or b"").decode("utf-8", "replace")


def check(rc: int) -> None:
    """Map a vs_status to the exception the reference would raise at the same point."""
    if rc == VS_OK:
        return
    msg = last_error()
    if rc == VS_ERR_INVALID:
        if "out of range" in msg:
            raise RuntimeError(msg)  # torch: "selected index k out of range" (upstream index.py:92)
        raise ValueError(msg)
    if rc == VS_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == VS_ERR_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)
