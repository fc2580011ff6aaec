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


def gen_rows(lo, hi, device, tokens=TOKENS, v=V):
    """Rows [lo, hi) of the synthetic bag-of-token index as an int32 column matrix [hi-lo, tokens].
    Generated in fixed 2^20-row blocks seeded by the block id so any sharding sees the same index."""
    import torch

    w = v // tokens
    base = ((torch.arange(tokens, device=device, dtype=torch.int64) * v) // tokens).to(torch.int32)
    parts = []
    for blk in range(lo // BLOCK_ROWS, (hi + BLOCK_ROWS - 1) // BLOCK_ROWS):
        b_lo, b_hi = blk * BLOCK_ROWS, min((blk + 1) * BLOCK_ROWS, N_TOTAL)
        g = torch.Generator(device=device).manual_seed(1234 + blk)
        r = torch.randint(0, w, (b_hi - b_lo, tokens), generator=g, device=device, dtype=torch.int32)
        s, e = max(lo, b_lo) - b_lo, min(hi, b_hi) - b_lo
        parts.append(r[s:e] + base[None, :])
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

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
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


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def profiled_traffic(mode="scan"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    ncu --set full capture of this workload (profiles/traffic.json), scaled to this batch size; else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        per_query = t.get(f"cfg2_{mode}_dram_bytes_per_query")
        return per_query * B if per_query else None
    except Exception:  # noqa: BLE001
        return None


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
def run_gpu(args):
    import torch
    import torch.distributed as dist

    import vsearch_b200 as vs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit(f"--gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- build this rank's shard of the synthetic index directly on the device
    lo, hi = vs.row_partition(N_TOTAL, world, rank)
    t0 = time.perf_counter()
    cols = gen_rows(lo, hi, dev)
    crow = torch.arange(hi - lo + 1, device=dev, dtype=torch.int64) * TOKENS
    index = vs.BoTIndex.from_token_csr(crow, cols.reshape(-1), (hi - lo, V), device=dev)
    index.search_mode = args.mode
    del cols, crow
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    eng = index._require_engine()
    sharded = vs.ShardedIndex(index, lo, N_TOTAL) if world > 1 else None
    retr = vs.Retriever(device=dev)
    retr.index = index

    q_host = gen_queries(nnz=args.qnnz).pin_memory()
    q_dev = q_host.to(dev)
    ids_host = torch.empty((B, K), dtype=torch.int64).pin_memory()
    sc_host = torch.empty((B, K), dtype=torch.float32).pin_memory()

    def step_resident():
        return sharded.search(q_dev, K) if sharded else retr.retrieve(q_dev, k=K)

    def step_e2e():
        res = sharded.search(q_host, K) if sharded else retr.retrieve(q_host, k=K)  # H2D of the batch inside
        ids_host.copy_(res.ids, non_blocking=True)
        sc_host.copy_(res.scores, non_blocking=True)
        return res

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    peak, peak_src = measured_peak()
    n_loc = hi - lo

    def measure(mode, steps, warmup):
        """One timed configuration: value (queries resident), e2e (host in / host out), roofline of the dominant
        kernel of THAT mode (CUDA events around every scoring launch inside the ABI)."""
        index.search_mode = mode
        sampler = ClockSampler(local)
        timed(step_resident, 0, warmup)
        eng.kernel_timer(reset=True)
        sampler.start()
        ms_total = timed(step_resident, steps, 0)
        clocks = sampler.stop()
        kern_ms, kern_n = eng.kernel_timer(reset=True)
        ms_step = ms_total / steps
        ms_e2e = timed(step_e2e, steps, min(warmup, 2)) / steps
        used = index.last_mode()
        if used == "scan":   # nnz * b_col + (N+1) * b_ptr, binary: b_val = 0; one pass per query (Q_tile = 1)
            bytes_pass = n_loc * TOKENS * 2 + (n_loc + 1) * 4
            kernel, launches = "vs::scan_topk_kernel<0, 3, 1, 0, 0>", steps * (3 if world == 1 else 4)
        else:                # K3: postings of the query's tokens (uint16 block-local row ids; binary: no values)
            bytes_pass = int(args.qnnz * (n_loc * TOKENS / V)) * 2
            kernel, launches = "vs::inv_search_kernel", steps * (4 if world == 1 else 5)
        achieved = B * bytes_pass / (kern_ms / max(kern_n, 1) * 1e-3) / 1e9 if kern_ms > 0 else None
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": (achieved / peak) if achieved else None,
                    "traffic": profiled_traffic(used) if world == 1 else None, "kernel": kernel, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": B * bytes_pass,
                    "streamed_bytes_per_launch": B * eng.stream_bytes if used == "scan" else None,
                    "kernel_ms_per_step": kern_ms / max(kern_n, 1), "steps_timed": kern_n,
                    "kernel_share_of_step": (kern_ms / ms_total) if ms_total else None,
                    "frac_of_8TBps": (achieved / 8000.0) if achieved else None}
        return {"mode": mode, "mode_used": used, "value": B / (ms_step * 1e-3), "ms_per_step": ms_step,
                "e2e": {"value": B / (ms_e2e * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": q_host.numel() * 4,
                        "d2h_bytes_per_step": B * K * 12, "ms_per_step": ms_e2e},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline}

    # headline = the passage-major scan (the HBM-roofline kernel the metric names); the same run also reports the
    # `auto` mode a user gets by default (token-major inverted lists for these 64-nnz queries)
    main_res = measure(args.mode, args.steps, args.warmup)
    extra = None
    if args.mode == "scan" and not args.no_auto:
        extra = measure("auto", args.steps, args.warmup)
    used_mode = main_res["mode_used"]
    qps, ms_step, clocks, roofline = main_res["value"], main_res["ms_per_step"], main_res["clocks"], main_res["roofline"]

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cqps, ct, cores, sample, _ = cpu_reference(3, 1, budget_s=25.0)
            cpu = {"value": cqps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample}
        line = {
            "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"cfg2: binary bag-of-token index {N_TOTAL:,} x 29,523, 120 tokens/row, B={B}, "
                                   f"{args.qnnz} nnz/query, k=100",
                       "parallelism": f"row-shard x{world}" if world > 1 else "single GPU",
                       "mode": args.mode, "mode_used": used_mode, "l2": "inputs larger than L2 (shard streams "
                                                f"{eng.stream_bytes / 1e9:.2f} GB per query pass; L2 is 126 MB)",
                       "index_build_s": round(build_s, 2)},
            "clocks": clocks,
            "e2e": main_res["e2e"],
            "gpu_launches": main_res["gpu_launches"],
            "roofline": roofline,
        }
        if extra:
            line["auto_mode"] = {k: extra[k] for k in ("mode_used", "value", "ms_per_step", "e2e", "gpu_launches", "roofline")}
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    global B, N_TOTAL
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="scan", choices=["auto", "scan", "inverted"],
                    help="kernel family of the headline line (default: the passage-major scan)")
    ap.add_argument("--no-auto", action="store_true", help="skip the extra `auto`-mode measurement")
    ap.add_argument("--qnnz", type=int, default=QNNZ)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=B, help="queries per step (default: the config's 1024; smaller only for profiling)")
    ap.add_argument("--rows", type=int, default=N_TOTAL, help="index rows (default: the config's 21,015,324; smaller only for profiling)")
    args = ap.parse_args()
    B, N_TOTAL = args.batch, args.rows
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
