#!/usr/bin/env python
"""BASELINE.json config 5 and the config-2 shape variants of SURVEY.md §8(d).

  cfg5        sparse fp32 21,015,324 x 29,523, 256 nnz/row (32.4 GB compact): query nnz {32..768} x batch {1,16,256(,4096)}
              x k {10,100,1000}, `scan` and `inverted` forced + what `auto` picks  -> the scan / inverted-list crossover
  cfg2_768    config 2 (binary, 120 tokens/row) with 768-nnz queries (the reference's default `a=768`, retriever.py:134)
  cfg2_ragged config 2 with row lengths ~ clipped-normal(120, 40) in [16, 256]
  cfg2_86     config 2 at 86 tokens/row (the real BoT density, build_binary_token_index.sh:15)
  cfg2_zipf   config 2 with Zipf(1) token popularity in rows AND queries (heavy-tailed posting lists)

One JSON line per point: ms per call (CUDA events around the public search call, device-resident queries), queries/s,
the scoring kernels' own time (vs_kernel_timer), the mode used.
usage: python scripts/sweep_crossover.py [cfg5 cfg2_768 cfg2_ragged cfg2_86 cfg2_zipf] [--quick]"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vsearch_b200 as vs  # noqa: E402,F401
from vsearch_b200 import _native as nat  # noqa: E402
from vsearch_b200.index import _Engine  # noqa: E402

V = 29523
N21 = 21_015_324
dev = torch.device("cuda:0")
QUICK = "--quick" in sys.argv


def emit(**kw):
    print(json.dumps(kw), flush=True)


def strat_cols(n, m, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    w = V // m
    base = ((torch.arange(m, device=dev, dtype=torch.int64) * V) // m).to(torch.int32)
    out = torch.empty((n, m), dtype=torch.int32, device=dev)
    step = 1 << 20
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        out[lo:hi] = torch.randint(0, w, (hi - lo, m), generator=g, device=dev, dtype=torch.int32) + base[None, :]
    return out


def ragged_cols(n, seed, mean=120.0, std=40.0, lo_len=16, hi_len=256):
    """Row lengths ~ clipped normal; columns stratified per row: col[r,j] = floor(j*V/m_r) + U{0..floor(V/m_r)-1}."""
    g = torch.Generator(device=dev).manual_seed(seed)
    lens = torch.empty(n, device=dev).normal_(mean, std, generator=g).round_().clamp_(lo_len, hi_len).to(torch.int64)
    crow = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens, 0, out=crow[1:])
    nnz = int(crow[-1])
    cols = torch.empty(nnz, dtype=torch.int32, device=dev)
    step = 1 << 19
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        a, b = int(crow[lo]), int(crow[hi])
        row = torch.repeat_interleave(torch.arange(lo, hi, device=dev), lens[lo:hi])
        j = torch.arange(a, b, device=dev) - crow[row]
        m = lens[row]
        w = V // m
        u = (torch.rand(b - a, generator=g, device=dev) * w).to(torch.int64).clamp_(max=w - 1)
        cols[a:b] = ((j * V) // m + u).to(torch.int32)
    return crow, cols


def zipf_cdf():
    p = 1.0 / torch.arange(1, V + 1, dtype=torch.float64, device=dev)
    perm = torch.randperm(V, generator=torch.Generator(device=dev).manual_seed(5), device=dev)  # popularity rank -> token id
    return torch.cumsum(p / p.sum(), 0), perm


def zipf_rows(n, draws, seed):
    """`draws` Zipf(1) token draws per row, duplicates dropped (so rows are sets, like real bag-of-token rows)."""
    g = torch.Generator(device=dev).manual_seed(seed)
    cdf, perm = zipf_cdf()
    lens = torch.empty(n, dtype=torch.int64, device=dev)
    parts = []
    step = 1 << 19
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        u = torch.rand((hi - lo, draws), generator=g, device=dev, dtype=torch.float64)
        tok = perm[torch.searchsorted(cdf, u).clamp_(max=V - 1)].to(torch.int32)
        tok, _ = torch.sort(tok, dim=1)
        keep = torch.ones_like(tok, dtype=torch.bool)
        keep[:, 1:] = tok[:, 1:] != tok[:, :-1]
        lens[lo:hi] = keep.sum(1)
        parts.append(tok[keep])
    crow = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens, 0, out=crow[1:])
    return crow, torch.cat(parts)


def queries(b, nnz, seed=4321, zipf=False):
    g = torch.Generator(device=dev).manual_seed(seed + nnz)
    q = torch.zeros((b, V), device=dev)
    if zipf:
        cdf, perm = zipf_cdf()
        u = torch.rand((b, nnz), generator=g, device=dev, dtype=torch.float64)
        cols = perm[torch.searchsorted(cdf, u).clamp_(max=V - 1)]
    else:
        cols = torch.rand((b, V), generator=g, device=dev).topk(nnz, dim=1).indices
    vals = torch.rand((b, nnz), generator=g, device=dev) * 2.99 + 0.01
    return q.scatter_(1, cols, vals)


def timed(fn, reps, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def point(eng, name, B, qnnz, k, mode, bytes_pass, zipf=False, reps=None):
    q = queries(B, qnnz, zipf=zipf)
    try:
        heavy = B >= 1024
        ms = timed(lambda: eng.search(q, k, mode=mode), reps=reps or (1 if heavy else 3), warm=0 if heavy else 1)
        eng.kernel_timer(reset=True)
        if not heavy:
            eng.search(q, k, mode=mode)
            kms, _ = eng.kernel_timer(reset=True)
        else:
            kms = None
        m = ctypes.c_int()
        nat.check(nat.LIB.vs_index_last_mode(eng.handle, m))
        used = "inverted" if m.value == nat.VS_MODE_INVERTED else "scan"
    except Exception as e:  # noqa: BLE001
        emit(config=name, B=B, qnnz=qnnz, k=k, mode=mode, error=str(e)[:200])
        return
    gbs = B * bytes_pass / (kms * 1e-3) / 1e9 if (used == "scan" and kms) else None
    emit(config=name, B=B, qnnz=qnnz, k=k, mode=mode, mode_used=used, ms_per_call=round(ms, 3), qps=round(B / ms * 1e3, 1),
         kernel_ms=None if kms is None else round(kms, 3), scan_GBps=None if gbs is None else round(gbs, 1))


def cfg5():
    n, m = N21, 256
    cols = strat_cols(n, m, 1234)
    vals = torch.rand(n * m, generator=torch.Generator(device=dev).manual_seed(99), device=dev) * 1.99 + 0.01
    crow = torch.arange(n + 1, device=dev, dtype=torch.int64) * m
    eng = _Engine.from_csr(crow, cols.reshape(-1), vals, (n, V), dev)
    del cols, vals, crow
    torch.cuda.empty_cache()
    bp = n * m * 6 + (n + 1) * 4
    emit(config="cfg5", note="index built", stream_bytes=eng.stream_bytes, bytes_pass=bp,
         mem_GB=round(torch.cuda.memory_allocated() / 1e9, 1))
    ks = (100,) if QUICK else (10, 100, 1000)
    for k in ks:                      # scan cost does not depend on the query's nnz: one column of the table
        for B in (1, 16, 256):
            point(eng, "cfg5", B, 64, k, "scan", bp)
    for qnnz in (32, 64, 128, 256, 512, 768):
        for k in ks:
            for B in (1, 16, 256):
                point(eng, "cfg5", B, qnnz, k, "inverted", bp)
                point(eng, "cfg5", B, qnnz, k, "auto", bp, reps=1)
    if not QUICK:
        point(eng, "cfg5", 4096, 64, 100, "scan", bp)
        for qnnz in (32, 128, 768):
            point(eng, "cfg5", 4096, qnnz, 100, "inverted", bp)
    emit(config="cfg5", note="done", peak_mem_GB=round(torch.cuda.max_memory_allocated() / 1e9, 1))


def binary_case(name, crow, cols, qnnz_list=(64,), zipf=False, batches=(1, 256)):
    n = crow.numel() - 1
    eng = _Engine.from_csr(crow, cols, None, (n, V), dev)
    nnz = int(cols.numel())
    bp = nnz * 2 + (n + 1) * 4
    emit(config=name, note="index built", rows=n, nnz=nnz, mean_row=round(nnz / n, 2), stream_bytes=eng.stream_bytes, bytes_pass=bp)
    del crow, cols
    torch.cuda.empty_cache()
    for qnnz in qnnz_list:
        for B in batches:
            for mode in ("scan", "inverted", "auto"):
                point(eng, name, B, qnnz, 100, mode, bp, zipf=zipf)


def cfg2_768():
    n, m = N21, 120
    binary_case("cfg2_768", torch.arange(n + 1, device=dev, dtype=torch.int64) * m, strat_cols(n, m, 1234).reshape(-1),
                qnnz_list=(64, 768))


def cfg2_ragged():
    crow, cols = ragged_cols(N21, 1234)
    binary_case("cfg2_ragged", crow, cols)


def cfg2_86():
    n, m = N21, 86
    binary_case("cfg2_86", torch.arange(n + 1, device=dev, dtype=torch.int64) * m, strat_cols(n, m, 1234).reshape(-1))


def cfg2_zipf():
    crow, cols = zipf_rows(N21, 160, 1234)
    binary_case("cfg2_zipf", crow, cols, qnnz_list=(64,), zipf=True)


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["cfg2_768", "cfg2_ragged", "cfg2_86", "cfg2_zipf", "cfg5"]
    for w in which:
        globals()[w]()
        torch.cuda.empty_cache()
