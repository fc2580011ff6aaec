#!/usr/bin/env python
"""Scan-kernel experiment driver (GPU box): config-2 shaped binary index of --rows rows, --batch queries through the
passage-major scan; prints achieved GB/s on algorithmic bytes, the per-pass time, and (with --prof) the mean phase
breakdown of a pass from the kernel's own %globaltimer marks (vs_debug_scan_profile).
    python scripts/exp_scan.py --rows 2626916 --batch 256 --prof"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import vsearch_b200 as vs  # noqa: E402
from vsearch_b200 import _native as nat  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=bench.N_TOTAL)
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--qnnz", type=int, default=64)
ap.add_argument("--tokens", type=int, default=bench.TOKENS)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--prof", action="store_true")
ap.add_argument("--mode", default="scan")
ap.add_argument("--check", action="store_true", help="compare the scan's ids with the inverted lists'")
ap.add_argument("--sparse-queries", action="store_true", help="queries as (token, weight) lists (torch sparse CSR), as bench.py sends them")
ap.add_argument("--values", action="store_true", help="a valued fp32 SparseIndex (values U(0.01, 2)) instead of the binary one: config 3 / 5 shapes")
ap.add_argument("--zipf", action="store_true", help="Zipf(1) token popularity in rows (160 draws, duplicates dropped) and queries (scripts/sweep_crossover.py cfg2_zipf)")
args = ap.parse_args()

dev = torch.device("cuda:0")
bench.N_TOTAL = args.rows
if args.zipf:
    sys.argv = sys.argv[:1]
    import sweep_crossover as sw   # noqa: E402  (same directory)

    crow, cols = sw.zipf_rows(args.rows, 160, 1234)
else:
    cols = bench.gen_rows(0, args.rows, dev, tokens=args.tokens)
    crow = torch.arange(args.rows + 1, device=dev, dtype=torch.int64) * args.tokens
torch.cuda.synchronize()
import time  # noqa: E402

t0 = time.perf_counter()
nnz_total = int(cols.numel())
if args.values:
    from vsearch_b200.index import _Engine   # noqa: E402

    vals = torch.rand(nnz_total, generator=torch.Generator(device=dev).manual_seed(99), device=dev) * 1.99 + 0.01
    index = vs.SparseIndex()
    index._engine = _Engine.from_csr(crow, cols.reshape(-1), vals, (args.rows, bench.V), dev)
    index.device = dev
    del vals
else:
    index = vs.BoTIndex.from_token_csr(crow, cols.reshape(-1), (args.rows, bench.V), device=dev)
torch.cuda.synchronize()
build_s = time.perf_counter() - t0
del cols, crow
index.search_mode = args.mode
eng = index._require_engine()
q = sw.queries(args.batch, args.qnnz, zipf=True) if args.zipf else bench.gen_queries(b=args.batch, nnz=args.qnnz).to(dev)
if args.sparse_queries:
    q = q.to_sparse_csr()
n_ctas = torch.cuda.get_device_properties(0).multi_processor_count
for _ in range(2):
    index.search(q, args.k)
torch.cuda.synchronize()
best = None
for _ in range(args.reps):
    eng.kernel_timer(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = index.search(q, args.k)
    e1.record()
    torch.cuda.synchronize()
    kms, kn = eng.kernel_timer(reset=True)
    call_ms = e0.elapsed_time(e1)
    if best is None or kms < best[0]:
        best = (kms, call_ms)
kms, call_ms = best
bytes_pass = nnz_total * (6 if args.values else 2) + (args.rows + 1) * 4
out = {"rows": args.rows, "batch": args.batch, "k": args.k, "mode": index.last_mode(), "build_s": round(build_s, 2),
       "kernel_ms": kms, "call_ms": call_ms, "us_per_pass": kms * 1e3 / args.batch,
       "GBps_algorithmic": args.batch * bytes_pass / (kms * 1e-3) / 1e9, "qps": args.batch / (call_ms * 1e-3),
       "stream_bytes": eng.stream_bytes}
if args.prof and index.last_mode() == "scan":
    buf = torch.zeros(n_ctas * args.batch * 8, dtype=torch.int64, device=dev)
    nat.check(nat.LIB.vs_debug_scan_profile(eng.handle, ctypes.c_void_p(buf.data_ptr())))
    index.search(q, args.k)
    torch.cuda.synchronize()
    nat.check(nat.LIB.vs_debug_scan_profile(eng.handle, None))
    t = buf.view(n_ctas, args.batch, 8).double()
    spans = {"stage_query": (0, 1), "stream_warp0": (1, 4), "wait_slowest_warp": (4, 5), "final_write": (5, 6)}
    ph = {}
    for nme, (a, b_) in spans.items():
        d = (t[:, 1:, b_] - t[:, 1:, a]) / 1e3   # us; skip the first pass (cold)
        ph[nme] = round(float(d.mean()), 2)
    tot = (t[:, 1:, 6] - t[:, 1:, 0]) / 1e3
    ph["pass_total_mean"] = round(float(tot.mean()), 2)
    ph["pass_total_max_cta_mean"] = round(float(tot.mean(dim=1).max()), 2)
    ph["pass_total_min_cta_mean"] = round(float(tot.mean(dim=1).min()), 2)
    out["phases_us"] = ph
if args.prof and index.last_mode() == "inverted":
    buf = torch.zeros(n_ctas * args.batch * 16, dtype=torch.int64, device=dev)
    nat.check(nat.LIB.vs_debug_scan_profile(eng.handle, ctypes.c_void_p(buf.data_ptr())))
    index.search(q, args.k)
    torch.cuda.synchronize()
    nat.check(nat.LIB.vs_debug_scan_profile(eng.handle, None))
    t = buf.view(args.batch, n_ctas, 16).double() / 1e3   # us per (query, CTA)
    names = ["setup", "zero", "accumulate", "first_block_histogram", "block_select", "refresh_compact", "final_write", "total",
             "first_block_select_first_rows", "first_block_select_refreshes", "first_block_select_rest"]
    live = t[:, :, 7].sum(dim=0) > 0   # CTAs that own row blocks (a small index launches fewer than n_ctas)
    out["phases_us_per_query_cta"] = {n: round(float(t[:, live, i].mean()), 2) for i, n in enumerate(names)}
    out["ctas_per_query"] = int(live.sum())
wf = torch.zeros(2, dtype=torch.int64, device=dev)
nat.check(nat.LIB.vs_debug_gather_wavefronts(eng.handle, ctypes.c_void_p(wf.data_ptr()), None))
torch.cuda.synchronize()
out["wavefronts_per_gather"] = round(wf[0].item() / max(wf[1].item(), 1), 4)
if args.check:
    index.search_mode = "inverted"
    r2 = index.search(q, args.k)
    out["ids_equal_inverted"] = bool(torch.equal(res.ids, r2.ids))
    out["scores_equal_inverted"] = bool(torch.equal(res.scores, r2.scores))
print(json.dumps(out))
