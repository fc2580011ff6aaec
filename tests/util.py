"""Shared helpers for the test-suite: golden loading and seeded synthetic inputs."""
import glob
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
V = 29523


def golden_search_cases(kind=None):
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN, "search_*.npz"))):
        z = np.load(f)
        if kind is None or str(z["kind"]) == kind:
            out.append(os.path.basename(f)[len("search_"):-len(".npz")])
    return out


def load_golden(name):
    z = dict(np.load(os.path.join(GOLDEN, f"search_{name}.npz")))
    z["kind"] = str(z["kind"])
    z["k"] = int(z["k"])
    if z["kind"] == "csr":
        v = int(z["shape"][1])
        q = np.zeros((z["q_idx"].shape[0], v), dtype=np.float32)
        np.put_along_axis(q, z["q_idx"].astype(np.int64), z["q_val"], axis=1)
        if bool(z.get("one_d", False)):
            q = q[0]
        z["q"] = q
    return z


def stratified_csr(n, v, m, seed, grid=True, binary=False, device="cpu", jitter=0):
    """SURVEY.md 8d generator: col[r,j] = floor(j*v/m) + U{0..floor(v/m)-1}; sorted & distinct.

    jitter>0 drops a random suffix of up to `jitter` entries per row (variable lengths)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    w = v // m
    base = (torch.arange(m, dtype=torch.int64) * v) // m
    col = base[None, :] + torch.randint(0, w, (n, m), generator=g)
    if binary:
        val = torch.ones(n, m)
    elif grid:
        val = torch.randint(1, 256, (n, m), generator=g).float() / 64.0
    else:
        val = torch.rand(n, m, generator=g) * 1.99 + 0.01
    if jitter:
        lens = m - torch.randint(0, jitter + 1, (n,), generator=g)
        keep = torch.arange(m)[None, :] < lens[:, None]
        crow = torch.zeros(n + 1, dtype=torch.int64)
        crow[1:] = torch.cumsum(lens, 0)
        col, val = col[keep], val[keep]
    else:
        crow = torch.arange(n + 1, dtype=torch.int64) * m
        col, val = col.reshape(-1), val.reshape(-1)
    return crow.to(device), col.to(device), val.to(device)


def sparse_queries(b, v, nnz, seed, grid=True, neg=False):
    g = torch.Generator(device="cpu").manual_seed(seed)
    q = torch.zeros(b, v)
    for i in range(b):
        idx = torch.randperm(v, generator=g)[:nnz]
        if grid:
            val = torch.randint(1, 193, (nnz,), generator=g).float() / 64.0
        else:
            val = torch.rand(nnz, generator=g) * 2.99 + 0.01
        if neg:
            val = val * (torch.randint(0, 2, (nnz,), generator=g).float() * 2 - 1)
        q[i, idx] = val
    return q
