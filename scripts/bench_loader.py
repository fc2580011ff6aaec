#!/usr/bin/env python
"""Index load time, shard files -> searchable device index: the reference's loader steps
(SparseIndex.init_index, index.py:163-179: sorted glob -> scipy.sparse.load_npz -> vstack -> torch CSR -> .to(device))
against the native direct-to-device loader (vs_index_load_npz), on the same files.

    python scripts/bench_loader.py [--shards 8] [--rows 2626916] [--tokens 120] [--dir /tmp/vs_loader]

Files: config-2 shaped bag-of-token shards (int32 indices, float32 ones -- scipy cannot load float16 members any more,
SURVEY.md 8c -- int32 row pointers), written once with the native block-parallel writer.  Every phase runs in its own
process so that its peak resident set (ru_maxrss) is its own; one JSON line per phase."""
import argparse
import glob
import json
import os
import resource
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
V = 29523


def rss_gb():
    return resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6   # ru_maxrss is in KB on Linux


def phase_gen(a):
    import numpy as np
    import torch

    import bench
    from vsearch_b200 import npz_io

    os.makedirs(a.dir, exist_ok=True)
    t0 = time.perf_counter()
    total = a.shards * a.rows
    for s in range(a.shards):
        path = os.path.join(a.dir, f"index{s}.npz")
        if os.path.exists(path):
            continue
        cols = bench.gen_rows(s * a.rows, (s + 1) * a.rows, "cuda:0", tokens=a.tokens, n_total=total).cpu().numpy().reshape(-1)
        indptr = (np.arange(a.rows + 1, dtype=np.int64) * a.tokens).astype(np.int32)
        npz_io.save_csr_npz_native(path, indptr, cols, np.ones(cols.size, dtype=np.float32), (a.rows, V), level=6)
    size = sum(os.path.getsize(f) for f in glob.glob(os.path.join(a.dir, "index*.npz")))
    print(json.dumps({"phase": "generate", "shards": a.shards, "rows_per_shard": a.rows, "tokens": a.tokens,
                      "file_bytes": size, "seconds": round(time.perf_counter() - t0, 2), "peak_rss_gb": round(rss_gb(), 2)}))


def _check(index, a):
    import torch

    import bench

    q = bench.gen_queries(b=4, nnz=64)
    res = index.search(q, 10)
    torch.cuda.synchronize()
    return [int(x) for x in res.ids[0, :3].tolist()], int(index._require_engine().n_rows)


def phase_reference(a):
    """The reference's steps with stock scipy / torch, then this engine's build from the device CSR (so that both
    phases end with a searchable index)."""
    import numpy as np
    import scipy.sparse as sp
    import torch

    import vsearch_b200 as vs
    from vsearch_b200.index import _Engine

    torch.cuda.init()
    t0 = time.perf_counter()
    files = sorted(glob.glob(os.path.join(a.dir, "index*.npz")))
    mats = [sp.load_npz(f) for f in files]                                   # index.py:174
    t_load = time.perf_counter() - t0
    m = sp.vstack(mats).tocsr()                                              # index.py:175
    del mats
    t_stack = time.perf_counter() - t0
    x = torch.sparse_csr_tensor(torch.from_numpy(m.indptr), torch.from_numpy(m.indices), torch.from_numpy(m.data), size=m.shape)  # :154
    x = x.to("cuda:0")                                                       # index.py:179
    torch.cuda.synchronize()
    t_dev = time.perf_counter() - t0
    idx = vs.BoTIndex()
    idx._engine = _Engine.from_csr(x.crow_indices(), x.col_indices(), None, tuple(x.shape), torch.device("cuda:0"))
    idx.device = "cuda:0"
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    top, n = _check(idx, a)
    print(json.dumps({"phase": "reference loader steps (scipy load_npz + vstack + torch CSR + .to(device)) + engine build",
                      "seconds_load_npz": round(t_load, 2), "seconds_to_vstack": round(t_stack, 2),
                      "seconds_to_device": round(t_dev, 2), "seconds_total": round(t_all, 2), "peak_rss_gb": round(rss_gb(), 2),
                      "rows": n, "top3_of_query0": top}))


def phase_native(a):
    import torch

    import vsearch_b200 as vs

    torch.cuda.init()
    t0 = time.perf_counter()
    idx = vs.BoTIndex(os.path.join(a.dir, "index*.npz"), fp16=False, device="cuda:0")
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    top, n = _check(idx, a)
    print(json.dumps({"phase": "native loader (vs_index_load_npz: parallel inflate -> pinned staging -> device CSR -> build)",
                      "seconds_total": round(t_all, 2), "peak_rss_gb": round(rss_gb(), 2), "rows": n, "top3_of_query0": top,
                      "binary": idx._engine.kind == 2, "threads": os.cpu_count()}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shards", type=int, default=8)
    ap.add_argument("--rows", type=int, default=2_626_916)
    ap.add_argument("--tokens", type=int, default=120)
    ap.add_argument("--dir", default="/tmp/vs_loader")
    ap.add_argument("--phase", default="all", choices=["all", "generate", "reference", "native"])
    a = ap.parse_args()
    if a.phase == "all":
        for ph in ("generate", "native", "reference"):
            subprocess.check_call([sys.executable, os.path.abspath(__file__), "--phase", ph, "--shards", str(a.shards), "--rows", str(a.rows),
                                   "--tokens", str(a.tokens), "--dir", a.dir])
        return
    {"generate": phase_gen, "reference": phase_reference, "native": phase_native}[a.phase](a)


if __name__ == "__main__":
    main()
