#!/usr/bin/env python
"""bench.py -- headline benchmark of the index-scoring hot path (BASELINE.json):
queries/sec on the Wikipedia-21M-shape binary bag-of-token index, k=100, 1024 queries per step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path over one batch of B=1024 synthetic queries:
Retriever.retrieve -> Index.search (C ABI: query prep, K2 scan + fused top-k, K6 merge); with N > 1 the
21,015,324 rows are split into N contiguous shards (one rank per GPU), each rank searches its shard and the
k x (score, id) keys are exchanged with ONE all-gather, then merged (strong scaling: the index is fixed).

value  = queries/sec with the query batch already resident in HBM;
e2e    = the same through the public call with HOST (pinned) queries and a device->host read of ids+scores;
roofline = achieved HBM GB/s of the scan kernel on ALGORITHMIC bytes (nnz*2 + (N+1)*4 per pass, one pass per
           query), CUDA events around each launch, vs the measured copy peak in MEASURED_PEAKS.json;
auto_mode = the same step through the default `auto` mode a user gets (for these 64-nnz queries the engine picks
           the K3 inverted lists): value / e2e / roofline of that run, reported next to the headline scan;
sharded_parity (N > 1) = before anything is timed, three seeded cases go through ShardedIndex (local fused top-k ->
           all-gather -> merge) and rank 0 compares ids and scores bit for bit with the CPU oracle on the full index;
           a mismatch ends the run with a non-zero exit code;
cfg3_b1 / cfg3_b256 / cfg4_dense = the other BASELINE.json configs (MS MARCO-shape fp32 sparse index, k=1000, batch 1
           and 256; dense 21M x 768 bf16, 4096 queries), each with its own clocks, roofline and e2e, row-sharded
           through ShardedIndex when N > 1 (--no-extras skips them);
cpu_baseline / --impl reference = the reference's own CPU path (torch CSR matmul + topk, restated in
           oracle/ref_search.py because the reference package cannot be imported/installed -- DESIGN.md) on a
           bounded row sample, linearly extrapolated in N.
Synthetic data: SURVEY.md 8d generator (stratified distinct sorted columns), seeds 1234+block / 4321.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

V = 29523
N_TOTAL = 21_015_324
TOKENS = 120
B = 1024
K = 100
QNNZ = 64
BLOCK_ROWS = 1 << 20
METRIC = "queries/sec on 21M-passage index, k=100; HBM GB/s vs 8 TB/s peak (1/2/4/8 GPU)"
# config 3 (MS MARCO passage collection shape) and config 4 (dense) of BASELINE.json
N_CFG3, NNZ_CFG3, K_CFG3 = 8_841_823, 256, 1000
N_CFG4, D_CFG4, B_CFG4 = 21_015_324, 768, 4096


def gen_rows(lo, hi, device, tokens=TOKENS, v=V, n_total=None, seed=1234):
    """Rows [lo, hi) of a synthetic index as an int32 column matrix [hi-lo, tokens] (stratified distinct sorted
    columns).  Generated in fixed 2^20-row blocks seeded by the block id so any sharding sees the same index."""
    import torch

    n_total = N_TOTAL if n_total is None else n_total
    w = v // tokens
    base = ((torch.arange(tokens, device=device, dtype=torch.int64) * v) // tokens).to(torch.int32)
    parts = []
    for blk in range(lo // BLOCK_ROWS, (hi + BLOCK_ROWS - 1) // BLOCK_ROWS):
        b_lo, b_hi = blk * BLOCK_ROWS, min((blk + 1) * BLOCK_ROWS, n_total)
        g = torch.Generator(device=device).manual_seed(seed + blk)
        r = torch.randint(0, w, (b_hi - b_lo, tokens), generator=g, device=device, dtype=torch.int32)
        s, e = max(lo, b_lo) - b_lo, min(hi, b_hi) - b_lo
        parts.append(r[s:e] + base[None, :])
    return torch.cat(parts, dim=0) if len(parts) > 1 else parts[0]


def gen_block_values(lo, hi, width, device, n_total, seed, kind):
    """Per-row values for rows [lo, hi): `uniform` U(0.01, 2) fp32 [rows, width] or `normal` N(0, 1) bf16, generated in
    the same seeded 2^20-row blocks as gen_rows."""
    import torch

    parts = []
    for blk in range(lo // BLOCK_ROWS, (hi + BLOCK_ROWS - 1) // BLOCK_ROWS):
        b_lo, b_hi = blk * BLOCK_ROWS, min((blk + 1) * BLOCK_ROWS, n_total)
        g = torch.Generator(device=device).manual_seed(seed + blk)
        if kind == "uniform":
            r = torch.rand((b_hi - b_lo, width), generator=g, device=device) * 1.99 + 0.01
        else:
            r = torch.randn((b_hi - b_lo, width), generator=g, device=device).to(torch.bfloat16)
        s, e = max(lo, b_lo) - b_lo, min(hi, b_hi) - b_lo
        parts.append(r[s:e])
    return torch.cat(parts, dim=0) if len(parts) > 1 else parts[0]


def gen_queries(b=None, v=V, nnz=QNNZ, seed=4321):
    """[b, v] fp32 dense-stored queries with `nnz` non-zeros each, U(0.01, 3) (host tensor)."""
    import torch

    b = B if b is None else b
    g = torch.Generator().manual_seed(seed)
    cols = torch.rand(b, v, generator=g).topk(nnz, dim=1).indices
    vals = torch.rand(b, nnz, generator=g) * 2.99 + 0.01
    return torch.zeros(b, v).scatter_(1, cols, vals)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_ms=200):
        self.rows, self.proc, self.gpu, self.period = [], None, gpu_index, period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", str(self.period)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, x in zip(names, r[3:7]) if x.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        return {"hbm": float(m["hbm_gbs"]), "bf16_burst": float(m["bf16_tflops"]),
                "bf16_sustained": float(m["bf16_tflops_sustained"]), "src": "measured (MEASURED_PEAKS.json)"}
    except Exception:  # noqa: BLE001
        return {"hbm": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


def profiled_traffic(mode="scan"):
    """dram__bytes_read.sum + dram__bytes_write.sum per query pass of the dominant kernel, from the committed
    ncu --set full capture of this workload (profiles/traffic.json); (bytes per B-query launch, source) or (None, None).
    It is a constant taken from that capture, not a counter read during this run."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        per_query = t.get(f"cfg2_{mode}_dram_bytes_per_query")
        if not per_query:
            return None, None
        return per_query * B, f"profiles/traffic.json: {t.get(f'cfg2_{mode}_source', 'ncu --set full')}; per query x {B}"
    except Exception:  # noqa: BLE001
        return None, None


def smem_atomic_ceiling(op="atomicAdd(uint32)"):
    """Measured shared-memory atomic-add rate of one B200 (all SMs), from the committed micro-benchmark run: the native
    32-bit integer add K3 issues on a binary index (fixed-point weights), or the fp32 CAS loop of a valued index."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2e_smem_atomics.json")) as f:
            return float(json.load(f)[op]["ops_per_s"]), f"profiles/r2e_smem_atomics.json (scripts/micro/smem_atomics.cu), {op}"
    except Exception:  # noqa: BLE001
        return None, None


# ---------------------------------------------------------------------------------------------- CPU reference
def cpu_reference(steps, warmup, budget_s=25.0, rows=1_000_000, bq=16):
    """The reference's CPU path (index.py:88-94 via oracle/ref_search.py) on a bounded sample:
    `rows` rows of the same synthetic index x `bq` of the same queries, all host threads; q/s is
    extrapolated linearly in N to the 21M-row index (cost of the reference is proportional to nnz * B)."""
    import torch

    from oracle import ref_search

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cols = gen_rows(0, rows, "cpu").to(torch.int64).reshape(-1)
    crow = torch.arange(rows + 1, dtype=torch.int64) * TOKENS
    X = ref_search.torch_csr(crow, cols, torch.ones(cols.numel()), (rows, V))
    q = gen_queries()[:bq]
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        res = ref_search.ref_search(q, X, K)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        if time.perf_counter() - t_start > budget_s and len(times) >= 1:
            break
    assert res.ids.shape == (bq, K)
    t = statistics.median(times)
    qps = bq / (t * (N_TOTAL / rows))
    sample = (f"{rows} of {N_TOTAL} rows x {bq} of {B} queries per step, torch {torch.__version__} CSR matmul+topk "
              f"fp32 (reference index.py:91-92), median of {len(times)} steps, q/s extrapolated linearly in N")
    return qps, t, cores, sample, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    qps, t, cores, sample, n = cpu_reference(args.steps, args.warmup, budget_s=120.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": n, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: binary bag-of-token index 21,015,324 x 29,523, 120 tokens/row, B=1024, "
                               "64 nnz/query, k=100", "parallelism": "cpu", "sampled": True},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- GPU arm
class Ctx:
    """Per-process state of the GPU arm."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            sys.exit(f"--gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus}")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.peaks = measured_peaks()

    def timed(self, fn, steps, warmup):
        """W untimed + K timed steps, barrier + synchronize on both sides, CUDA events, max over ranks (ms total)."""
        torch, dist = self.torch, self.dist
        for _ in range(warmup):
            fn()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if steps == 0:
            return 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def kernel_ms(self, eng):
        """Mean scoring-kernel time per timed call on this rank's stream (CUDA events inside the ABI), max over ranks."""
        torch, dist = self.torch, self.dist
        ms, n = eng.kernel_timer(reset=True)
        t = torch.tensor([ms / max(n, 1), float(n)], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0].item()), int(t[1].item())

    def measure(self, step_resident, step_e2e, eng, nq, steps, warmup, e2e_steps=None):
        """One timed configuration -> value (queries resident), e2e (host in / host out), clocks, kernel time."""
        sampler = ClockSampler(self.local)
        self.timed(step_resident, 0, warmup)
        eng.kernel_timer(reset=True)
        sampler.start()
        ms_total = self.timed(step_resident, steps, 0)
        clocks = sampler.stop()
        kern_ms, kern_n = self.kernel_ms(eng)
        ms_step = ms_total / steps
        out = {"value": nq / (ms_step * 1e-3), "ms_per_step": ms_step, "steps": steps, "clocks": clocks,
               "kernel_ms_per_launch": kern_ms, "kernel_launches_per_step": kern_n / steps,
               "kernel_share_of_step": kern_ms * kern_n / ms_total if ms_total else None}
        if step_e2e is not None:
            e2e_steps = e2e_steps or steps
            ms_e2e = self.timed(step_e2e, e2e_steps, min(warmup, 2)) / e2e_steps
            out["e2e"] = {"value": nq / (ms_e2e * 1e-3), "unit": "queries/s", "ms_per_step": ms_e2e}
        return out


def sharded_parity(ctx):
    """N > 1: the row-sharded path against the CPU oracle on the full index (the cases of tests/nccl_parity.py), before
    anything is timed.  Returns 'ok' on every rank or raises SystemExit(3)."""
    import vsearch_b200 as vs
    from oracle import ref_search
    from tests.util import sparse_queries, stratified_csr

    torch, dist = ctx.torch, ctx.dist
    fails = []
    for binary, n, m, k, mode in [(True, 300_001, 60, 100, "scan"), (False, 120_000, 128, 1000, "scan"),
                                  (True, 300_001, 60, 100, "inverted")]:
        crow, col, val = stratified_csr(n, V, m, seed=5, grid=True, binary=binary, jitter=11)
        lo, hi = vs.row_partition(n, ctx.world, ctx.rank)
        c = crow[lo:hi + 1] - crow[lo]
        sl = slice(int(crow[lo]), int(crow[hi]))
        idx = (vs.BoTIndex if binary else vs.SparseIndex)()
        idx.vector = ref_search.torch_csr(c, col[sl], val[sl], (hi - lo, V))
        idx.move_to_device(ctx.dev)
        idx.search_mode = mode
        q = sparse_queries(9, V, 200, seed=3)
        if binary:
            q = (q != 0).float()  # heavy ties across shard boundaries
        res = vs.ShardedIndex(idx, lo, n).search(q, k)
        torch.cuda.synchronize()
        if ctx.rank == 0:
            X = ref_search.torch_csr(crow, col, val, (n, V))
            msg = ref_search.compare_results(res, ref_search.ref_scores(q, X), k, exact=True)
            if msg is not None:
                fails.append(f"binary={binary} n={n} k={k} mode={mode}: {msg}")
        del idx
    # dense index, thresholds pooled between the ranks (vs_search_dense_step), and the one-all-gather path beside it
    for steps in (True, False):
        n, d, B, k = 1_000_003, 64, 6, 100
        g = torch.Generator().manual_seed(21)
        x = torch.randint(-8, 9, (n, d), generator=g).float() / 4.0
        q = torch.randint(-8, 9, (B, d), generator=g).float() / 4.0
        lo, hi = vs.row_partition(n, ctx.world, ctx.rank)
        idx = vs.Index()
        idx.vector = x[lo:hi].to(torch.bfloat16)
        idx.move_to_device(ctx.dev)
        sh = vs.ShardedIndex(idx, lo, n)
        sh.dense_steps = steps
        res = sh.search(q, k)
        torch.cuda.synchronize()
        if ctx.rank == 0:
            canon = ref_search.canonical_topk(ref_search.quantize_like(ref_search.ref_scores(q, x), torch.bfloat16), k)
            if not (torch.equal(res.ids.cpu(), canon.ids) and torch.equal(res.scores.float().cpu(), canon.scores)):
                fails.append(f"dense n={n} k={k} stepwise={steps}: ids / scores differ from the reference's")
        del idx, sh
    flag = torch.tensor([len(fails)], device=ctx.dev)
    dist.broadcast(flag, 0)
    if int(flag.item()):
        if ctx.rank == 0:
            print(json.dumps({"sharded_parity": "FAILED", "detail": fails}), flush=True)
        dist.destroy_process_group()
        raise SystemExit(3)
    torch.cuda.empty_cache()
    return "ok"


def run_cfg2(ctx):
    """Headline: config 2.  Returns (line fields, cleanup)."""
    import vsearch_b200 as vs

    torch, args, dev, world = ctx.torch, ctx.args, ctx.dev, ctx.world
    lo, hi = vs.row_partition(N_TOTAL, world, ctx.rank)
    t0 = time.perf_counter()
    cols = gen_rows(lo, hi, dev)
    crow = torch.arange(hi - lo + 1, device=dev, dtype=torch.int64) * TOKENS
    index = vs.BoTIndex.from_token_csr(crow, cols.reshape(-1), (hi - lo, V), device=dev)
    del cols, crow
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    eng = index._require_engine()
    sharded = vs.ShardedIndex(index, lo, N_TOTAL) if world > 1 else None
    retr = vs.Retriever(device=dev)
    retr.index = index

    # queries travel as (token, weight) lists -- a torch sparse CSR tensor, what the reference's sparsifier emits
    # (utils/sparse.py:8-19) -- unless --dense-queries asks for the [B, V] fp32 rows of round 1
    q_host = gen_queries(nnz=args.qnnz)
    q_host = q_host.pin_memory() if args.dense_queries else q_host.to_sparse_csr()
    q_dev = q_host.to(dev)
    h2d_bytes = (q_host.numel() * 4 if args.dense_queries else
                 q_host.crow_indices().numel() * 8 + q_host.col_indices().numel() * 8 + q_host.values().numel() * 4)
    ids_host = torch.empty((B, K), dtype=torch.int64).pin_memory()
    sc_host = torch.empty((B, K), dtype=torch.float32).pin_memory()

    def step_resident():
        return sharded.search(q_dev, K) if sharded else retr.retrieve(q_dev, k=K)

    def step_e2e():
        res = sharded.search(q_host, K) if sharded else retr.retrieve(q_host, k=K)  # H2D of the batch inside
        ids_host.copy_(res.ids, non_blocking=True)
        sc_host.copy_(res.scores, non_blocking=True)
        return res

    n_loc = hi - lo

    def measure(mode, steps, warmup):
        index.search_mode = mode
        m = ctx.measure(step_resident, step_e2e, eng, B, steps, warmup)
        used = index.last_mode()
        kern_ms = m["kernel_ms_per_launch"]
        if used == "scan":   # nnz * b_col + (N+1) * b_ptr, binary: b_val = 0; one pass per query (Q_tile = 1)
            bytes_pass = n_loc * TOKENS * 2 + (n_loc + 1) * 4
            kernel = "vs::scan_bin_kernel<0, 0>"
        else:                # K3: postings of the query's tokens (uint16 block-local row ids; binary: no values)
            bytes_pass = int(args.qnnz * (n_loc * TOKENS / V)) * 2
            kernel = "vs::inv_search_kernel"
        # our kernels per step: [prep_query (dense rows only)] + scan + merge; auto / inverted add the list kernel (extract
        # or lists-from-CSR), the fixed-point kernel (binary index), the decision kernel and the inverted-list kernel (the
        # loser of the two scoring kernels exits at once); N > 1 adds the merge of the gathered keys
        launches = steps * ((2 if mode == "scan" else 6) + (1 if args.dense_queries else 0) + (1 if world > 1 else 0))
        achieved = B * bytes_pass / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else None
        traffic, traffic_src = profiled_traffic(used) if world == 1 else (None, None)
        peak = ctx.peaks["hbm"]
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                    "kernel": kernel, "peak_source": ctx.peaks["src"] + " hbm_gbs",
                    "algorithmic_bytes_per_launch": B * bytes_pass,
                    "streamed_bytes_per_launch": B * eng.stream_bytes if used == "scan" else None,
                    "kernel_ms_per_step": kern_ms, "steps_timed": steps,
                    "kernel_share_of_step": m["kernel_share_of_step"],
                    "frac_of_8TBps": (achieved / 8000.0) if achieved else None, "q_tile": 1}
        if used != "scan":
            # K3's work is shared-memory read-modify-writes, not HBM bytes: postings/s against the measured ceiling of the
            # operation it issues on this binary index (ATOMS.ADD on fixed-point weights; the fp32 CAS loop of valued
            # indices peaks at a third of that; scripts/micro/smem_atomics.cu).  The kernel is latency- and
            # barrier-bound well below either (profiles/README.md)
            postings = args.qnnz * (n_loc * TOKENS / V)
            ceil_ops, ceil_src = smem_atomic_ceiling()
            pps = B * postings / (kern_ms * 1e-3) if kern_ms > 0 else None
            roofline["shared_atomics"] = {"achieved_postings_per_s": pps, "ceiling_ops_per_s": ceil_ops,
                                          "frac": (pps / ceil_ops) if (pps and ceil_ops) else None, "ceiling_source": ceil_src}
        m["e2e"].update({"h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": B * K * 12})
        return {"mode": mode, "mode_used": used, "value": m["value"], "ms_per_step": m["ms_per_step"], "e2e": m["e2e"],
                "gpu_launches": launches, "clocks": m["clocks"], "roofline": roofline}

    # headline = the passage-major scan (the HBM-roofline kernel the metric names); the same run also reports the
    # `auto` mode a user gets by default (token-major inverted lists for these 64-nnz queries)
    main_res = measure(args.mode, args.steps, args.warmup)
    extra = measure("auto", args.steps, args.warmup) if (args.mode == "scan" and not args.no_auto) else None
    stream_gb = eng.stream_bytes / 1e9
    del index, eng, sharded, retr, q_dev
    torch.cuda.empty_cache()
    return main_res, extra, build_s, stream_gb


def sparse_extra(ctx, name, batches):
    """Config 3: MS MARCO-shape fp32 sparse index (8,841,823 x 29,523, 256 nnz/row), k=1000, through `auto` (what a
    user gets) and through the scan (the HBM-roofline kernel), per batch size."""
    import vsearch_b200 as vs
    from vsearch_b200.index import _Engine

    torch, dev, world = ctx.torch, ctx.dev, ctx.world
    lo, hi = vs.row_partition(N_CFG3, world, ctx.rank)
    n_loc = hi - lo
    t0 = time.perf_counter()
    cols = gen_rows(lo, hi, dev, tokens=NNZ_CFG3, n_total=N_CFG3, seed=3234)
    vals = gen_block_values(lo, hi, NNZ_CFG3, dev, N_CFG3, 5234, "uniform")
    crow = torch.arange(n_loc + 1, device=dev, dtype=torch.int64) * NNZ_CFG3
    index = vs.SparseIndex()
    index._engine = _Engine.from_csr(crow, cols.reshape(-1), vals.reshape(-1), (n_loc, V), dev)
    index.device = dev
    del cols, vals, crow
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    eng = index._engine
    sharded = vs.ShardedIndex(index, lo, N_CFG3) if world > 1 else None
    bytes_pass = n_loc * NNZ_CFG3 * 6 + (n_loc + 1) * 4   # uint16 column + fp32 value per entry, uint32 row pointers
    out = {}
    for bq in batches:
        q_host = gen_queries(b=bq, nnz=QNNZ, seed=777)
        q_host = q_host.pin_memory() if ctx.args.dense_queries else q_host.to_sparse_csr()
        q_dev = q_host.to(dev)
        h2d_bytes = (q_host.numel() * 4 if ctx.args.dense_queries else
                     q_host.crow_indices().numel() * 8 + q_host.col_indices().numel() * 8 + q_host.values().numel() * 4)
        ids_host = torch.empty((bq, K_CFG3), dtype=torch.int64).pin_memory()
        sc_host = torch.empty((bq, K_CFG3), dtype=torch.float32).pin_memory()

        def step_resident():
            return sharded.search(q_dev, K_CFG3) if sharded else index.search(q_dev, K_CFG3)

        def step_e2e():
            res = sharded.search(q_host, K_CFG3) if sharded else index.search(q_host, K_CFG3)
            ids_host.copy_(res.ids, non_blocking=True)
            sc_host.copy_(res.scores, non_blocking=True)

        res = {"workload": f"cfg3: sparse fp32 index {N_CFG3:,} x {V:,}, {NNZ_CFG3} nnz/row, B={bq}, {QNNZ} nnz/query, "
                           f"k={K_CFG3}", "parallelism": f"row-shard x{world}" if world > 1 else "single GPU",
               "index_build_s": round(build_s, 2)}
        for mode in ("auto", "scan"):
            index.search_mode = mode
            steps = (20 if bq == 1 else (10 if mode == "auto" else 2))
            m = ctx.measure(step_resident, step_e2e, eng, bq, steps, 2, e2e_steps=max(2, steps // 2))
            used = index.last_mode()
            m["mode_used"] = used
            m["e2e"].update({"h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": bq * K_CFG3 * 12})
            if used == "scan":
                kms, launches = m["kernel_ms_per_launch"], max(1.0, m["kernel_launches_per_step"])   # the ABI cuts large batches
                ach = bq * bytes_pass / launches / (kms * 1e-3) / 1e9 if kms > 0 else None
                m["roofline"] = {"bound": "hbm", "achieved": ach, "peak": ctx.peaks["hbm"], "unit": "GB/s",
                                 "frac": ach / ctx.peaks["hbm"] if ach else None, "traffic": None,
                                 "kernel": "vs::scan_topk_kernel<1, 2, 0, 0, 0>", "algorithmic_bytes_per_launch": bq * bytes_pass / launches,
                                 "peak_source": ctx.peaks["src"] + " hbm_gbs"}
            res[mode] = m
        out[f"{name}_b{bq}"] = res
        del q_dev
    del index, eng, sharded
    torch.cuda.empty_cache()
    return out


def dense_extra(ctx):
    """Config 4: dense bf16 index 21,015,324 x 768, 4096 queries, k=100 (tcgen05 GEMM with the top-k fused in)."""
    import vsearch_b200 as vs
    from vsearch_b200.index import _Engine

    torch, dev, world = ctx.torch, ctx.dev, ctx.world
    lo, hi = vs.row_partition(N_CFG4, world, ctx.rank)
    n_loc = hi - lo
    t0 = time.perf_counter()
    x = torch.empty((n_loc, D_CFG4), dtype=torch.bfloat16, device=dev)
    b_lo = lo
    while b_lo < hi:     # block-seeded like the sparse generators, one 2^20-row block at a time
        b_hi = min(hi, (b_lo // BLOCK_ROWS + 1) * BLOCK_ROWS)
        x[b_lo - lo:b_hi - lo] = gen_block_values(b_lo, b_hi, D_CFG4, dev, N_CFG4, 7234, "normal")
        b_lo = b_hi
    index = vs.Index(fp16=False)
    index._vector = None
    index._engine = _Engine.from_dense(x, dev, torch.bfloat16)
    index.device = dev
    index._logical_dtype = torch.bfloat16
    del x
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    eng = index._engine
    sharded = vs.ShardedIndex(index, lo, N_CFG4) if world > 1 else None
    g = torch.Generator().manual_seed(4321)
    q_host = torch.randn((B_CFG4, D_CFG4), generator=g).pin_memory()
    q_dev = q_host.to(dev)
    ids_host = torch.empty((B_CFG4, K), dtype=torch.int64).pin_memory()
    sc_host = torch.empty((B_CFG4, K), dtype=torch.float32).pin_memory()

    def step_resident():
        return sharded.search(q_dev, K) if sharded else index.search(q_dev, K)

    def step_e2e():
        res = sharded.search(q_host, K) if sharded else index.search(q_host, K)
        ids_host.copy_(res.ids, non_blocking=True)
        sc_host.copy_(res.scores.float(), non_blocking=True)

    m = ctx.measure(step_resident, step_e2e, eng, B_CFG4, 5, 2, e2e_steps=3)
    flops = 2.0 * B_CFG4 * n_loc * D_CFG4
    call_tf = flops / (m["ms_per_step"] * 1e-3) / 1e12
    kern_tf = flops / (m["kernel_ms_per_launch"] * m["kernel_launches_per_step"] * 1e-3) / 1e12 if m["kernel_ms_per_launch"] else None
    m["e2e"].update({"h2d_bytes_per_step": q_host.numel() * 4, "d2h_bytes_per_step": B_CFG4 * K * 12})
    m["roofline"] = {"bound": "tensor", "achieved": kern_tf, "peak": ctx.peaks["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": kern_tf / ctx.peaks["bf16_sustained"] if kern_tf else None,
                     "frac_of_burst": kern_tf / ctx.peaks["bf16_burst"] if kern_tf else None,
                     "peak_burst": ctx.peaks["bf16_burst"], "traffic": None, "kernel": "vs::dense_topk_pair_kernel",
                     "algorithmic_flops_per_step_per_gpu": flops, "whole_call_tflops_per_gpu": call_tf,
                     "note": "achieved = 2*B*N_shard*D / device time of the sweeps of one call (sample + filtered sweeps, "
                             "CUDA events inside the ABI); peak = sustained cuBLAS bf16 rate, frac_of_burst vs the burst rate",
                     "peak_source": ctx.peaks["src"]}
    if sharded is not None:   # the same through the one-all-gather path (every rank finds its thresholds alone)
        sharded.dense_steps = False
        m1 = ctx.measure(step_resident, step_e2e, eng, B_CFG4, 5, 2, e2e_steps=2)
        sharded.dense_steps = True
        k1 = flops / (m1["kernel_ms_per_launch"] * m1["kernel_launches_per_step"] * 1e-3) / 1e12 if m1["kernel_ms_per_launch"] else None
        m["without_threshold_sharing"] = {"value": m1["value"], "ms_per_step": m1["ms_per_step"], "sweeps_tflops_per_gpu": k1,
                                          "whole_call_tflops_per_gpu": flops / (m1["ms_per_step"] * 1e-3) / 1e12}
        m["roofline"]["note"] += "; N > 1: thresholds pooled between the ranks after every sweep (vs_search_dense_step), the " \
                                 "timed sweeps include the two key all-gathers in between"
    m["workload"] = f"cfg4: dense bf16 index {N_CFG4:,} x {D_CFG4}, B={B_CFG4}, k={K}"
    m["parallelism"] = f"row-shard x{world}" if world > 1 else "single GPU"
    m["index_build_s"] = round(build_s, 2)
    del index, eng, sharded, q_dev
    torch.cuda.empty_cache()
    return {"cfg4_dense": m}


def run_gpu(args):
    ctx = Ctx(args)
    torch, dist, world, rank = ctx.torch, ctx.dist, ctx.world, ctx.rank
    parity = sharded_parity(ctx) if world > 1 else None
    main_res, extra, build_s, stream_gb = run_cfg2(ctx)
    extras = {}
    if not args.no_extras:
        for fn in (lambda: sparse_extra(ctx, "cfg3", [1, 256]), lambda: dense_extra(ctx)):
            try:
                extras.update(fn())
            except Exception as e:  # noqa: BLE001 -- an extra must never take the headline line down with it
                extras.setdefault("errors", []).append(f"{type(e).__name__}: {e}"[:300])
                torch.cuda.empty_cache()
    used_mode = main_res["mode_used"]
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cqps, ct, cores, sample, _ = cpu_reference(3, 1, budget_s=25.0)
            cpu = {"value": cqps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample}
        line = {
            "metric": METRIC, "value": main_res["value"], "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"cfg2: binary bag-of-token index {N_TOTAL:,} x 29,523, 120 tokens/row, B={B}, "
                                   f"{args.qnnz} nnz/query, k=100",
                       "parallelism": f"row-shard x{world}" if world > 1 else "single GPU",
                       "mode": args.mode, "mode_used": used_mode,
                       "queries": "dense [B, V] fp32 rows" if args.dense_queries else
                                  f"sparse (token, weight) lists, {args.qnnz} per query (torch sparse CSR -> vs_search_sparse)",
                       "l2": "inputs larger than L2 (shard streams "
                                                f"{stream_gb:.2f} GB per query pass; L2 is 126 MB)",
                       "index_build_s": round(build_s, 2)},
            "clocks": main_res["clocks"],
            "e2e": main_res["e2e"],
            "gpu_launches": main_res["gpu_launches"],
            "roofline": main_res["roofline"],
        }
        if parity:
            line["sharded_parity"] = parity
        if extra:
            line["auto_mode"] = {k: extra[k] for k in ("mode_used", "value", "ms_per_step", "e2e", "gpu_launches", "clocks", "roofline")}
        line.update(extras)
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    global B, N_TOTAL, N_CFG3, N_CFG4
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="scan", choices=["auto", "scan", "inverted"],
                    help="kernel family of the headline line (default: the passage-major scan)")
    ap.add_argument("--no-auto", action="store_true", help="skip the extra `auto`-mode measurement")
    ap.add_argument("--no-extras", action="store_true", help="skip the cfg3 / cfg4 measurements")
    ap.add_argument("--qnnz", type=int, default=QNNZ)
    ap.add_argument("--dense-queries", action="store_true", help="queries as dense [B, V] fp32 rows instead of sparse (token, weight) lists")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=B, help="queries per step (default: the config's 1024; smaller only for profiling)")
    ap.add_argument("--rows", type=int, default=N_TOTAL, help="index rows (default: the config's 21,015,324; smaller only for profiling)")
    ap.add_argument("--extras-rows", type=int, default=0, help="shrink the cfg3 / cfg4 indices to this many rows (smoke runs only)")
    args = ap.parse_args()
    B, N_TOTAL = args.batch, args.rows
    if args.extras_rows:
        N_CFG3 = N_CFG4 = args.extras_rows
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
