#!/usr/bin/env python
"""Secondary measurements for the other BASELINE.json configs (the headline, config 2, is bench.py):
  cfg1  SparseIndex 100k x 29,523 fp32 CSR, ~256 nnz/row, B=64, k=100   (+ reference torch path on CPU and GPU)
  cfg3  SVDR-shape sparse fp32 8,841,823 x 29,523, 256 nnz/row, B=1 and B=256, k=1000
  cfg4  dense 21,015,324 x 768 bf16, B=4096, k=100   (tcgen05 GEMM + fused top-k)
Each line: config, mode, ms per call, queries/s, achieved GB/s or TFLOP/s of the scoring kernels (CUDA events).
usage: python scripts/bench_configs.py [cfg1 cfg3 cfg4 torchgpu] """
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vsearch_b200 as vs  # noqa: E402

V = 29523
dev = torch.device("cuda:0")


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


def queries(b, nnz, seed=4321):
    g = torch.Generator().manual_seed(seed)
    cols = torch.rand(b, V, generator=g).topk(nnz, dim=1).indices
    vals = torch.rand(b, nnz, generator=g) * 2.99 + 0.01
    return torch.zeros(b, V).scatter_(1, cols, vals).to(dev)


def timed(fn, reps=5, warm=2):
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


def torch_path(q, X, k):
    """The reference's own two calls (upstream index.py:89-92) through stock torch: the 'beat that' baseline."""
    with torch.no_grad():
        return torch.matmul(q.to(X.device).type(X.dtype), X.t()).topk(k)


def emit(**kw):
    print(json.dumps(kw), flush=True)


def sparse_case(name, n, m, batches, k, modes=("scan", "inverted", "auto"), qnnz=64):
    cols = strat_cols(n, m, 1234)
    g = torch.Generator(device=dev).manual_seed(99)
    vals = torch.rand(n * m, generator=g, device=dev) * 1.99 + 0.01
    crow = torch.arange(n + 1, device=dev, dtype=torch.int64) * m
    idx = vs.SparseIndex()
    from vsearch_b200.index import _Engine
    idx._engine = _Engine.from_csr(crow, cols.reshape(-1), vals, (n, V), dev)
    idx.device = dev
    eng = idx._engine
    bytes_pass = n * m * 6 + (n + 1) * 4
    for B in batches:
        q = queries(B, qnnz)
        for mode in modes:
            idx.search_mode = mode
            try:
                ms = timed(lambda: eng.search(q, k, mode=mode), reps=3 if B > 16 else 10)
            except Exception as e:  # noqa: BLE001
                emit(config=name, B=B, k=k, mode=mode, error=str(e)[:200])
                continue
            eng.kernel_timer(reset=True)
            eng.search(q, k, mode=mode)
            kms, kn = eng.kernel_timer(reset=True)
            used = idx.last_mode()
            gbs = B * bytes_pass / (kms * 1e-3) / 1e9 if used == "scan" else None
            emit(config=name, n=n, nnz_per_row=m, B=B, k=k, mode=mode, mode_used=used, ms_per_call=ms, qps=B / ms * 1e3,
                 kernel_ms=kms, scan_GBps_algorithmic=gbs, frac_of_6555=(gbs / 6555.8 if gbs else None),
                 stream_bytes=eng.stream_bytes)
    return cols, vals, crow


def cfg1():
    n, m = 100_000, 256
    cols, vals, crow = sparse_case("cfg1", n, m, [64], 100)
    # reference torch path on the same data: GPU (B2 baseline) and CPU (B1 baseline)
    q = queries(64, 64)
    X = torch.sparse_csr_tensor(crow, cols.reshape(-1).to(torch.int64), vals, size=(n, V))
    try:
        ms = timed(lambda: torch_path(q, X, 100), reps=5)
        emit(config="cfg1", impl="reference torch CSR matmul+topk on the SAME B200 (cuSPARSE)", B=64, ms_per_call=ms, qps=64 / ms * 1e3)
    except Exception as e:  # noqa: BLE001
        emit(config="cfg1", impl="reference torch GPU", error=str(e)[:300])
    Xc, qc = X.cpu(), q.cpu()
    torch.set_num_threads(os.cpu_count())
    torch_path(qc, Xc, 100)
    t0 = time.perf_counter()
    for _ in range(3):
        torch_path(qc, Xc, 100)
    dt = (time.perf_counter() - t0) / 3
    emit(config="cfg1", impl=f"reference torch CSR matmul+topk on CPU ({os.cpu_count()} threads)", B=64, ms_per_call=dt * 1e3, qps=64 / dt)


def cfg3():
    sparse_case("cfg3", 8_841_823, 256, [1, 256], 1000)


def cfg4():
    n, d, B, k = 21_015_324, 768, 4096, 100
    g = torch.Generator(device=dev).manual_seed(7)
    x = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
    step = 1 << 20
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        x[lo:hi] = torch.randn((hi - lo, d), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    q = torch.randn((B, d), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    from vsearch_b200.index import _Engine
    eng = _Engine.from_dense(x, dev, torch.bfloat16)
    del x
    torch.cuda.empty_cache()
    from vsearch_b200 import _native as nat
    for Bq in (4096, 512):
        qq = q[:Bq]
        ms = timed(lambda: eng.search(qq, k, score_round=nat.VS_BF16), reps=3, warm=1)
        eng.kernel_timer(reset=True)
        eng.search(qq, k, score_round=nat.VS_BF16)
        kms, kn = eng.kernel_timer(reset=True)
        flops = 2.0 * Bq * n * d
        emit(config="cfg4", n=n, d=d, B=Bq, k=k, ms_per_call=ms, qps=Bq / ms * 1e3, kernel_ms=kms,
             TFLOPs_kernel=flops / (kms * 1e-3) / 1e12, frac_of_1383_sustained=flops / (kms * 1e-3) / 1e12 / 1383.1,
             TFLOPs_call=flops / (ms * 1e-3) / 1e12)
    # torch reference on the same GPU for one shard-sized slice (the full [4096, 21M] score matrix does not fit)
    n_s = 2_626_916
    xs = torch.randn((n_s, d), generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    ms = timed(lambda: torch.matmul(q, xs.t()).topk(k), reps=3, warm=1)
    emit(config="cfg4", impl="reference torch matmul+topk bf16 on the SAME B200, 1/8 shard (2,626,916 rows)", B=B,
         ms_per_call=ms, qps_shard=B / ms * 1e3, TFLOPs=2.0 * B * n_s * d / (ms * 1e-3) / 1e12)


def cfg2_torch():
    """B2 baseline of BASELINE.md: the reference's own torch path (cuSPARSE SpMM + topk) on the SAME B200 for the
    headline config (21M x 120 binary).  The [B, N] score matrix limits the batch (B=64 -> 5.4 GB)."""
    n, m = 21_015_324, 120
    cols = strat_cols(n, m, 1234).reshape(-1).to(torch.int64)
    crow = torch.arange(n + 1, device=dev, dtype=torch.int64) * m
    for dtype in (torch.float32, torch.float16):
        try:
            vals = torch.ones(n * m, device=dev, dtype=dtype)
            X = torch.sparse_csr_tensor(crow, cols, vals, size=(n, V))
            for B in (1, 64):
                q = queries(B, 64)
                ms = timed(lambda: torch_path(q, X, 100), reps=3, warm=1)
                emit(config="cfg2", impl=f"reference torch CSR matmul+topk on the SAME B200 (cuSPARSE), values {dtype}", B=B,
                     ms_per_call=ms, qps=B / ms * 1e3, peak_mem_GB=torch.cuda.max_memory_allocated() / 1e9)
            del X, vals
        except Exception as e:  # noqa: BLE001
            emit(config="cfg2", impl=f"reference torch GPU {dtype}", error=str(e)[:300])
        torch.cuda.empty_cache()


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg1", "cfg3", "cfg4"]
    for w in which:
        {"cfg1": cfg1, "cfg3": cfg3, "cfg4": cfg4, "cfg2_torch": cfg2_torch}[w]()
        torch.cuda.empty_cache()
