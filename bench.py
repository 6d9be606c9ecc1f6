#!/usr/bin/env python
"""bench.py -- tokens/s of the SCONE input-embedding lookup on B200, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload config2|config1|config3]

One "step" = one pass of the fused hot path (longest f-gram match + row gather + dequant + fallback) over
one [B, L] batch of synthetic token ids.  Default workload = BASELINE.json configs[1]:
GPT-2-medium shape, 1 M f-grams, max_n = 4, INT8 cache, batch 64 x 1024, bf16 output.

Printed JSON line (rank 0): `value` = whole-job tokens/s with inputs resident in HBM (K steps replayed as
one CUDA graph, timed with CUDA events, max over ranks); `e2e` = the same metric through
EmbeddingCache.lookup() with the ids coming from pinned HOST memory and the match result read back to the
host every step; `roofline` = algorithmic bytes / kernel time against MEASURED_PEAKS.json; `cpu_baseline` =
the oracle's Python port of the reference path timed on this box's cores.

`--impl reference` times that CPU port alone (all cores, fork pool over batch rows) -- the reference is pure
Python and cannot travel to the GPU box, so the port (oracle/py_oracle.py, pinned to fixtures generated from
the unmodified reference) stands in for it.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: N f-grams, D, V, max_n, quant, B, L   (BASELINE.json configs)
    "config1": dict(N=100_000, D=768, V=50_257, max_n=3, quant="fp16", B=8, L=512,
                    desc="GPT-2 small, small-100k f-grams, max_n=3, FP16 cache, batch 8x512"),
    "config2": dict(N=1_000_000, D=1024, V=50_257, max_n=4, quant="int8", B=64, L=1024,
                    desc="GPT-2 medium, medium-1m f-grams, max_n=4, INT8 cache, batch 64x1024"),
    "config3": dict(N=10_000_000, D=4096, V=128_000, max_n=5, quant="int4", B=256, L=2048,
                    desc="hidden 4096, 10M f-grams, max_n=5, INT4 g128 cache, batch 256x2048"),
    # N here is PER GPU (12.5 M rows x 8 192 B = 102 GB per GPU; 100 M f-grams at 8 GPUs)
    "config4": dict(N=12_500_000, D=4096, V=128_000, max_n=5, quant="fp16", B=256, L=2048, tier="sharded",
                    desc="hidden 4096, 100M f-grams (12.5M per GPU), FP16 cache row-sharded by id % W via NCCL all-to-all, batch 256x2048 per GPU"),
    # config 3's table row-sharded instead of replicated ("8-GPU run: replicas and sharded variant for comparison")
    "config3s": dict(N=1_250_000, D=4096, V=128_000, max_n=5, quant="int4", B=256, L=2048, tier="sharded",
                     desc="hidden 4096, 10M f-grams (1.25M per GPU), INT4 g128 cache row-sharded by id % W, batch 256x2048 per GPU"),
    # N is capped by the box's host RAM (pinned); 200 M rows need 416 GB
    "config5": dict(N=200_000_000, D=2048, V=128_000, max_n=5, quant="int8", B=256, L=2048, tier="host",
                    desc="hidden 2048, 200M f-grams, INT8 cache in pinned host RAM read zero-copy (TMA bulk over the host link), batch 256x2048"),
}
N_BATCHES = 8          # distinct id batches rotated through the timed steps (rows touched >> L2)
METRIC = "tokens/sec embedded"


def row_bytes_algorithmic(quant: str, D: int, group: int = 128) -> int:
    return {"fp16": 2 * D, "int8": D + 4, "int4": D // 2 + 2 * D // group}[quant]


def bytes_per_token(w, hit: float, probes: float, slot_bytes: int = 32) -> float:
    """id in + probed slots + hit row / fallback row + output + (id, len) out.  `slot_bytes` is the slot
    format the index actually built (32, or 16 for the compact format; the survey's own figure is 16)."""
    D = w["D"]
    return 8 + slot_bytes * probes + hit * row_bytes_algorithmic(w["quant"], D) + (1 - hit) * 2 * D + 2 * D + 5


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.f, self.p = gpu_index, None, None

    def __enter__(self):
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None
        return self

    def __exit__(self, *a):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except Exception:
                self.p.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.f is None:
            return out
        try:
            self.f.flush()
            rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
            os.unlink(self.f.name)
        except Exception:
            return out
        sm = sorted(float(r[1]) for r in rows if r[1].strip().replace(".", "").isdigit())
        if sm:
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_max_mhz"] = float(rows[0][2])
        out["samples"] = len(rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out["reasons"] = [n for k, n in enumerate(names) if any("Active" in r[5 + k] and "Not" not in r[5 + k] for r in rows)]
        return out


# ------------------------------------------------------------------------------------------------------------
# CPU arms (the ONLY place bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_rows(rows):
    """fork-pool worker: oracle port of the reference path on some batch rows."""
    from oracle import py_oracle as po
    s = _CPU
    out, fid, ml = po.embed_forward(s["g2i"], s["max_n"], s["table"], s["base_bits"], s["ids"][rows], "bf16")
    return int(out.shape[0] * out.shape[1]), int((fid >= 0).sum())


def cpu_setup(w, seed_rank=0):
    """Same construction as the GPU arm (same generators, run on the CPU device) at min(N, 2M) f-grams: Python
    dict probes are size-independent and the dict costs ~360 B / f-gram."""
    import numpy as np
    import torch
    from oracle import py_oracle as po
    from scone_b200.utils import synthetic as S
    N = min(w["N"], 2_000_000)
    toks, lens, longest = S.make_vocab_device(N, w["max_n"], w["V"], seed=0, device="cpu", return_longest=True)
    ids = S.make_stream_device(toks, lens, w["B"], w["L"], w["V"], seed=100 + seed_rank, p_plant=1.0, pick_ids=longest).numpy()
    toks, lens = toks.numpy(), lens.numpy()
    g2i = {tuple(r[:n]): i for i, (r, n) in enumerate(zip(toks.tolist(), lens.tolist()))}
    rng = np.random.default_rng(2)
    # one quantised 65 536-row block tiled over the table: same footprint and gather pattern as N independent rows
    D = w["D"]
    blk = po.OracleTable.from_fp32(rng.standard_normal((min(N, 65536), D), dtype=np.float32) * np.float32(0.02), w["quant"])
    reps = (N + 65535) // 65536
    payload = np.tile(blk.payload, (reps, 1))[:N]
    scales = None if blk.scales is None else np.tile(blk.scales, (reps,) + (1,) * (blk.scales.ndim - 1))[:N]
    table = po.OracleTable(w["quant"], D, payload, scales)
    base_bits = po.cast_bits(rng.standard_normal((w["V"], D), dtype=np.float32) * np.float32(0.02), "bf16")
    _CPU.update(toks=toks, lens=lens)
    _CPU.update(g2i=g2i, max_n=w["max_n"], table=table, base_bits=base_bits, ids=ids, N=N)
    return N


def cpu_time_rows(rows_per_step, cores):
    """One CPU step over `rows_per_step` batch rows; returns (seconds, tokens)."""
    import multiprocessing as mp
    import numpy as np
    B = _CPU["ids"].shape[0]
    rows = np.arange(rows_per_step) % B
    t0 = time.perf_counter()
    if cores == 1:
        tok, _ = _cpu_rows(rows)
    else:
        parts = [p for p in np.array_split(rows, cores) if len(p)]
        tok = sum(t for t, _ in _CPU["pool"].map(_cpu_rows, parts))
    return time.perf_counter() - t0, tok


def run_reference(args, w, rank, world):
    """--impl reference: the CPU port, all host cores, bounded sample per step."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    N = cpu_setup(w)
    _CPU["pool"] = mp.get_context("fork").Pool(cores) if cores > 1 else None
    B = w["B"]
    # size the per-step sample so the whole run stays within ~2 minutes
    dt, tok = cpu_time_rows(min(B, cores), cores)
    per_row = dt / max(1, min(B, cores))
    budget = 120.0 / max(1, args.steps + args.warmup)
    rows_per_step = int(max(1, min(B, budget / max(per_row, 1e-9))))
    for _ in range(args.warmup):
        cpu_time_rows(rows_per_step, cores)
    t_total, tok_total = 0.0, 0
    for _ in range(args.steps):
        dt, tok = cpu_time_rows(rows_per_step, cores)
        t_total += dt
        tok_total += tok
    if _CPU["pool"] is not None:
        _CPU["pool"].close()
    value = tok_total / t_total
    sample = f"{rows_per_step} of {B} batch rows x {w['L']} tokens per step; vocabulary {N} of {w['N']} f-grams (Python dict)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int8->bf16" if w["quant"] == "int8" else f"{w['quant']}->bf16", "data": "synthetic",
        "config": {"workload": w["desc"], "batch": [w["B"], w["L"]], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample,
                         "what": "oracle/py_oracle.py embed_forward: dict probes n=max_n..1 + numpy gather/dequant/cast"},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------

def run_ours(args, w, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist

    import scone_b200 as sb
    from scone_b200 import _lib
    from scone_b200.utils import synthetic as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- scone_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B, L, D, N, V = w["B"], w["L"], w["D"], w["N"], w["V"]
    T = B * L
    # ---- build (untimed): vocabulary, index, table, fallback rows, rotating id batches -----------------------
    toks, lens, longest = S.make_vocab_device(N, w["max_n"], V, seed=0, device=dev, return_longest=True)
    index = sb.FGramIndex(toks, lens)
    table = sb.CacheTable(N, D, w["quant"], device=dev)
    S.fill_table_device(table, seed=2)
    base = S.make_base_device(V, D, torch.bfloat16, seed=3, device=dev)
    if args.id_dist == "zipf":
        # secondary workload: planted f-gram ids follow a Zipf-like law instead of being uniform over the
        # table, so hot rows hit L2 and the algorithmic GB/s exceeds the DRAM GB/s
        gen = torch.Generator(device=dev)
        gen.manual_seed(7)
        ranks = S.zipf_like_torch(gen, (8 * T,), longest.numel(), dev)
        longest = longest[ranks]
    batches = [S.make_stream_device(toks, lens, B, L, V, seed=100 + rank * N_BATCHES + k, p_plant=1.0, pick_ids=longest)
               for k in range(N_BATCHES)]
    del toks, lens
    out = torch.empty((B, L, D), dtype=torch.bfloat16, device=dev)
    out_id = torch.empty((B, L), dtype=torch.int32, device=dev)
    out_len = torch.empty((B, L), dtype=torch.uint8, device=dev)
    status = torch.zeros((1,), dtype=torch.int32, device=dev)

    def step(k):
        sb.embed_forward(index, table, base, batches[k % N_BATCHES], out=out, status=status, out_id=out_id, out_len=out_len)

    # hit rate / probe count of the workload (for the algorithmic byte count)
    hits = 0
    for k in range(N_BATCHES):
        step(k)
        hits += int((out_id >= 0).sum().item())
    hit = hits / (N_BATCHES * T)
    probes = bin(index.len_mask).count("1")

    # ---- device-resident timing: K steps captured into one CUDA graph, replayed once ---------------------------
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        for k in range(max(3, args.warmup)):
            step(k)
        stream.synchronize()
        launches0 = _lib.launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            for k in range(args.steps):
                step(k)
        gpu_launches = _lib.launch_count() - launches0
        graph.replay()                       # one untimed replay (graph upload)
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        graph.replay()
        e1.record(stream)
        stream.synchronize()
        reps = max(1, min(400, int(0.4 / max(1e-5, e0.elapsed_time(e1) * 1e-3))))   # keep the GPU busy ~0.4 s either side for the clock samples
        barrier()
        with ClockSampler(local_rank) as clocks:
            for _ in range(reps):
                graph.replay()
            stream.synchronize()
            e0.record(stream)
            graph.replay()
            e1.record(stream)
            stream.synchronize()
            barrier()
            ms = e0.elapsed_time(e1)
            repeats = []
            for _ in range(4):                       # variance only; `value` is the ONE timed replay above
                r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                r0.record(stream)
                graph.replay()
                r1.record(stream)
                stream.synchronize()
                repeats.append(r0.elapsed_time(r1) / args.steps)
            for _ in range(reps):
                graph.replay()
            stream.synchronize()
        clk = clocks.summary()
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    assert int(status.item()) == 0
    value = world * T * args.steps / (ms * 1e-3)
    kernel_ms = ms / args.steps

    # ---- e2e: host ids -> EmbeddingCache-level call -> match result back on the host, every step ---------------
    h_ids = [b.cpu().pin_memory() for b in batches]
    h_id = torch.empty((B, L), dtype=torch.int32).pin_memory()
    h_len = torch.empty((B, L), dtype=torch.uint8).pin_memory()
    d_ids = torch.empty((B, L), dtype=torch.int64, device=dev)
    h_emb = torch.empty((B, L, D), dtype=torch.bfloat16).pin_memory() if args.e2e_embeds_to_host else None

    def e2e_step(k, embeds_to_host=False):
        d_ids.copy_(h_ids[k % N_BATCHES], non_blocking=True)
        sb.embed_forward(index, table, base, d_ids, out=out, status=status, out_id=out_id, out_len=out_len)
        h_id.copy_(out_id, non_blocking=True)
        h_len.copy_(out_len, non_blocking=True)
        if embeds_to_host:
            h_emb.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()            # the caller has its results

    def time_e2e(embeds_to_host):
        for k in range(max(3, args.warmup)):
            e2e_step(k, embeds_to_host)
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            e2e_step(k, embeds_to_host)
        dt = time.perf_counter() - t0
        barrier()
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return world * T * args.steps / dt

    sync_value = time_e2e(False)

    # the throughput form of the same call: scone_b200.HostPipeline double-buffers the batches, so the H2D copy of batch
    # k+1 and the D2H read of batch k-1 run under batch k's kernel.  Every step still copies its own ids from pinned host
    # memory and lands its own match result in pinned host memory inside the timed region.
    def time_pipelined():
        pipe = sb.HostPipeline(index, table, base, (B, L))
        for k in range(max(3, args.warmup)):
            pipe.submit(h_ids[k % N_BATCHES])
        pipe.flush()
        barrier()
        t0 = time.perf_counter()
        n_done = 0
        for k in range(args.steps):
            if pipe.submit(h_ids[k % N_BATCHES]) is not None:
                n_done += 1
        n_done += len(pipe.flush())
        dt = time.perf_counter() - t0
        barrier()
        assert n_done == args.steps
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return world * T * args.steps / dt

    pipe_value = time_pipelined()
    e2e = {"value": max(pipe_value, sync_value), "unit": "tokens/s", "h2d_bytes_per_step": T * 8, "d2h_bytes_per_step": T * 5,
           "mode": "pipelined (scone_b200.HostPipeline, 4 slots, up to 3 batches in flight)" if pipe_value >= sync_value else "synchronous call per step",
           "pipelined": pipe_value, "synchronous": sync_value,
           "note": "ids from pinned host memory; fgram_id + match_len read back to pinned host memory every step; the embeddings "
                   "stay in HBM for the transformer, as with the reference's get_embeddings(ids, device). `synchronous` = one "
                   "blocking lookup per step (copy in, kernel, copy out, stream sync)"}
    if args.e2e_embeds_to_host:
        e2e["embeds_to_host"] = {"value": time_e2e(True), "unit": "tokens/s", "d2h_bytes_per_step": T * 5 + T * D * 2}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline ---------------------------------------------------------------------------------------------
    bpt = bytes_per_token(w, hit, probes, slot_bytes=index.slot_bytes)
    peak, peak_src = measured_peak_hbm()
    achieved = bpt * T / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel": "embed_bulk_kernel (fused match+gather+dequant+fallback)", "slot_bytes": index.slot_bytes,
                "bytes_per_token": bpt, "hit_rate": hit, "probes_per_token": probes, "kernel_ms": kernel_ms}
    prof = os.path.join(ROOT, "profiles", f"traffic_{args.workload}.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof))["dram_bytes_per_launch"]
        except Exception:
            pass

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload -----------------------------------
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        Ncpu = cpu_setup(w)
        cpu_time_rows(1, 1)
        dt, tok, passes = 0.0, 0, 0
        while dt < 10.0 and passes < 64:             # ~10 s of single-thread CPU work over whole batches
            d1, t1 = cpu_time_rows(B, 1)
            dt, tok, passes = dt + d1, tok + t1, passes + 1
        cpu_baseline = {"value": tok / dt, "unit": "tokens/s", "cores": 1, "kind": "port",
                        "sample": f"{passes} passes over the {B} x {L} batch ({dt:.1f} s); vocabulary {Ncpu} of {N} f-grams (Python dict)",
                        "what": "oracle/py_oracle.py embed_forward (Python port of the reference path), single thread"}
        try:
            from oracle.c_oracle import COracleIndex
            from scone_b200.utils.synthetic import pack_table_numpy
            tb = _CPU["table"]
            cix = COracleIndex(_CPU["toks"], _CPU["lens"])
            packed, stride, soff = pack_table_numpy(w["quant"], tb.payload, tb.scales)
            cores = os.cpu_count() or 1
            obuf = np.empty((B, L, D), np.uint16)
            cix.embed(w["quant"], D, 128, packed, stride, packed[:, soff:] if soff else None, stride, _CPU["base_bits"],
                      _CPU["ids"][:1], "bf16", nthreads=1)
            t0 = time.perf_counter()
            cix.embed(w["quant"], D, 128, packed, stride, packed[:, soff:] if soff else None, stride, _CPU["base_bits"],
                      _CPU["ids"], "bf16", nthreads=cores, out=obuf)
            dtc = time.perf_counter() - t0
            cpu_baseline["c_port"] = {"value": T / dtc, "unit": "tokens/s", "cores": cores,
                                      "what": f"oracle/c_oracle.c, {cores} pthreads, one full batch"}
        except Exception as e:  # the C number is informational
            cpu_baseline["c_port"] = {"error": repr(e)}

    line = {
        "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": kernel_ms, "ms_per_step_repeats": repeats, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": f"{w['quant']}->bf16", "data": "synthetic",
        "config": {"workload": w["desc"], "f_grams": N, "dim": D, "max_n": w["max_n"], "quant": w["quant"], "batch": [B, L],
                   "per_gpu_batch": [B, L], "parallelism": f"replicas x{world} (table fits one GPU; no data-path collective)",
                   "hit_rate": hit, "id_dist": args.id_dist, "index_bytes": index.bytes, "index_slot_bytes": index.slot_bytes,
                   "table_bytes": table.bytes,
                   "l2": f"inputs > L2: {N_BATCHES} rotating id batches gather rows uniformly from a {table.bytes / 1e9:.2f} GB "
                         f"table and each step writes {T * D * 2 / 1e6:.0f} MB of output; no explicit flush",
                   "timing": "K steps captured in one CUDA graph, CUDA events on the launching stream, max over ranks"},
        "e2e": e2e, "gpu_launches": int(gpu_launches), "clocks": clk, "roofline": roofline,
    }
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _tile_fill(table, block_rows=65536, seed=2):
    """Fill a (possibly huge) table by tiling one quantised random block: same footprint / access pattern as N independent
    rows at a fraction of the generation time (benchmark data only; parity tests use real tables)."""
    import torch
    import scone_b200 as sb
    from scone_b200.utils import synthetic as S
    blk = sb.CacheTable(min(block_rows, table.num_rows), table.dim, table.quant, table.group, device=table.device)
    S.fill_table_device(blk, seed=seed)
    src = blk.storage
    for s0 in range(0, table.num_rows, src.shape[0]):
        k = min(src.shape[0], table.num_rows - s0)
        table.storage[s0:s0 + k].copy_(src[:k], non_blocking=True)
    torch.cuda.synchronize()


def _timed_steps(step, steps, warmup, barrier, dev, world):
    import torch
    import torch.distributed as dist
    for k in range(max(3, warmup)):
        step(k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for k in range(steps):
        step(k)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def run_sharded(args, w, rank, local_rank, world):
    """config 4: table row-sharded over the ranks, ids out / packed rows back over NVLink (scone_b200/sharded.py)."""
    import torch
    import torch.distributed as dist
    import scone_b200 as sb
    from scone_b200 import _lib, sharded
    from scone_b200.utils import synthetic as S
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    B, L, D, V = w["B"], w["L"], w["D"], w["V"]
    T = B * L
    rows_per_gpu = args.rows_per_gpu or w["N"]
    N = rows_per_gpu * world
    toks, lens, longest = S.make_vocab_device(N, w["max_n"], V, seed=0, device=dev, return_longest=True)
    index = sb.FGramIndex(toks, lens)
    mode = args.sharded_mode
    if mode == "peer":
        # the shard lives in symmetric memory mapped by every rank; the exchange happens inside the fused kernel
        ptab = sharded.PeerShardedTable(N, D, w["quant"], device=dev)
        table = ptab.local
    else:
        table = sb.CacheTable(sharded.shard_rows(N, rank, world), D, w["quant"], device=dev)
    _tile_fill(table)
    base = S.make_base_device(V, D, torch.bfloat16, seed=3, device=dev)
    batches = [S.make_stream_device(toks, lens, B, L, V, seed=100 + rank * N_BATCHES + k, p_plant=1.0, pick_ids=longest) for k in range(4)]
    del toks, lens, longest
    torch.cuda.empty_cache()
    out = torch.empty((B, L, D), dtype=torch.bfloat16, device=dev)
    fid, _ = index.lookup(batches[0])
    hit = float((fid >= 0).float().mean().item())
    n_hits = int((fid >= 0).sum().item())
    remote = int(((fid >= 0) & (fid % world != rank)).sum().item())
    if mode == "peer":
        ptab.publish()
        out_id = torch.empty((B, L), dtype=torch.int32, device=dev)
        out_len = torch.empty((B, L), dtype=torch.uint8, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)

        def step(k):
            sharded.embed_forward_sharded(index, ptab, base, batches[k % 4], out=out, status=status, out_id=out_id, out_len=out_len)
        launches_per_step = 1
    else:
        cache = sharded.ShardedEmbeddingCache(sharded.CudaOps(index, table, base))
        l0 = _lib.launch_count()
        cache.lookup(batches[0], out=out)
        launches_per_step = _lib.launch_count() - l0

        def step(k):
            cache.lookup(batches[k % 4], out=out)
    with ClockSampler(local_rank) as clocks:
        ms = _timed_steps(step, args.steps, args.warmup, barrier, dev, world)
    clk = clocks.summary()
    if rank == 0:
        value = world * T * args.steps / (ms * 1e-3)
        step_s = ms * 1e-3 / args.steps
        nv_in = remote * (table.row_stride + 4) / step_s / 1e9
        bpt = bytes_per_token(w, hit, bin(index.len_mask).count("1"))
        if mode != "peer":
            bpt += hit * 2 * table.row_stride                    # NCCL variant: served copy is read + written once more
        peak, peak_src = measured_peak_hbm()
        line = {
            "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": f"{w['quant']}->bf16", "data": "synthetic",
            "config": {"workload": w["desc"], "f_grams": N, "rows_per_gpu": rows_per_gpu, "dim": D, "max_n": w["max_n"], "quant": w["quant"],
                       "per_gpu_batch": [B, L], "parallelism": (f"row-sharded x{world}: index replicated, rows pulled from peer memory over NVLink inside the fused kernel (TMA bulk)"
                                       if mode == "peer" else f"row-sharded x{world}: index replicated, 2 NCCL all-to-alls per step"),
                       "hit_rate": hit, "remote_fraction": remote / max(1, n_hits), "index_bytes": index.bytes,
                       "table_bytes_per_gpu": table.bytes, "l2": "inputs > L2 (rows gathered uniformly from the shard)",
                       "sharded_mode": mode,
                       "timing": "eager steps, CUDA events, max over ranks"},
            "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8 * world * 2,
                    "note": "same call; ids resident on the device, only the all-to-all split sizes cross to the host"},
            "gpu_launches": int(launches_per_step * args.steps), "clocks": clk,
            "roofline": {"bound": "hbm", "achieved": bpt * T / step_s / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": bpt * T / step_s / 1e9 / peak, "traffic": None, "peak_source": peak_src, "bytes_per_token": bpt,
                         "note": "per GPU; this tier is NVLink-bound, see nvlink"},
            "nvlink": {"achieved_in_GBps_per_gpu": nv_in, "peak": 770.0, "frac": nv_in / 770.0,
                       "peak_source": "B200_PROFILING.md measured peer copy, per direction per GPU",
                       "bytes_in_per_step_per_gpu": remote * (table.row_stride + 4)},
        }
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def run_host(args, w, rank, local_rank, world):
    """config 5: the table lives in pinned host RAM and is read zero-copy by the same kernels."""
    import torch
    import scone_b200 as sb
    from scone_b200 import _lib
    from scone_b200.utils import synthetic as S
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    B, L, D, V = w["B"], w["L"], w["D"], w["V"]
    T = B * L
    stride, _ = sb.table_layout(w["quant"], D)
    total_ram = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES")
    budget = int(total_ram * (args.host_fraction or 0.5))
    N = int(min(w["N"], args.rows_per_gpu or (budget // stride)))
    t0 = time.perf_counter()
    table = sb.CacheTable(N, D, w["quant"], device=dev, tier="host")
    pin_s = time.perf_counter() - t0
    blk = sb.CacheTable(min(1 << 20, N), D, w["quant"], device=dev)
    S.fill_table_device(blk, seed=2)
    for s0 in range(0, N, blk.num_rows):
        k = min(blk.num_rows, N - s0)
        table.storage[s0:s0 + k].copy_(blk.storage[:k], non_blocking=True)
    torch.cuda.synchronize()
    del blk
    toks, lens, longest = S.make_vocab_device(N, w["max_n"], V, seed=0, device=dev, return_longest=True)
    index = sb.FGramIndex(toks, lens)
    base = S.make_base_device(V, D, torch.bfloat16, seed=3, device=dev)
    batches = [S.make_stream_device(toks, lens, B, L, V, seed=100 + k, p_plant=1.0, pick_ids=longest) for k in range(4)]
    del toks, lens, longest
    out = torch.empty((B, L, D), dtype=torch.bfloat16, device=dev)
    out_id = torch.empty((B, L), dtype=torch.int32, device=dev)
    out_len = torch.empty((B, L), dtype=torch.uint8, device=dev)

    def step(k):
        sb.embed_forward(index, table, base, batches[k % 4], out=out, out_id=out_id, out_len=out_len)

    step(0)
    hit = float((out_id >= 0).float().mean().item())
    l0 = _lib.launch_count()
    with ClockSampler(local_rank) as clocks:
        ms = _timed_steps(step, args.steps, args.warmup, torch.cuda.synchronize, dev, 1)
    clk = clocks.summary()
    step_s = ms * 1e-3 / args.steps
    link = hit * T * stride / step_s / 1e9
    # the staged variant of the same tier: host threads gather rows into pinned staging, cudaMemcpyAsync on a side stream
    from scone_b200.offload import StagedHostLookup
    staged = StagedHostLookup(index, table, base, micro_batches=8, max_positions=T)
    t0 = time.perf_counter()
    n_st = max(2, args.steps // 3)
    staged.lookup(batches[0], out=out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(n_st):
        staged.lookup(batches[k % 4], out=out)
    torch.cuda.synchronize()
    st_s = (time.perf_counter() - t0) / n_st
    staged_info = {"value": T / st_s, "unit": "tokens/s", "host_link_GBps": hit * T * stride / st_s / 1e9, "micro_batches": 8,
                   "host_threads": staged.threads, "ms_per_step": st_s * 1e3}
    line = {
        "metric": METRIC, "value": T / step_s, "unit": "tokens/s", "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": f"{w['quant']}->bf16", "data": "synthetic",
        "config": {"workload": w["desc"], "f_grams": N, "f_grams_named": w["N"], "dim": D, "max_n": w["max_n"], "quant": w["quant"],
                   "batch": [B, L], "hit_rate": hit, "host_ram_bytes": total_ram, "table_bytes_host": table.bytes, "pin_seconds": pin_s,
                   "scaled": f"N scaled from {w['N']} to {N} rows: the box has {total_ram / 1e9:.0f} GB of host RAM" if N < w["N"] else None,
                   "tier": "zero-copy: cp.async.bulk straight from pinned host memory into shared memory",
                   "l2": "rows come over the host link, never cached", "timing": "eager steps, CUDA events"},
        "e2e": {"value": T / step_s, "unit": "tokens/s", "h2d_bytes_per_step": int(hit * T * stride), "d2h_bytes_per_step": 0,
                "note": "the row bytes cross the host link inside the kernel (zero-copy)"},
        "gpu_launches": int(_lib.launch_count() - l0), "clocks": clk,
        "roofline": {"bound": "hbm", "achieved": link, "peak": 64.0, "unit": "GB/s", "frac": link / 64.0, "traffic": None,
                     "peak_source": "nominal PCIe Gen5 x16 per direction (host-link bound, not HBM)", "host_link_GBps": link},
        "staged": staged_info,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--id-dist", default="uniform", choices=["uniform", "zipf"],
                    help="distribution of the planted f-gram ids over the table: uniform (primary, worst case for caches) or Zipf-like")
    ap.add_argument("--rows-per-gpu", type=int, default=0, help="config4/5: table rows per GPU (default: the named size / what host RAM allows)")
    ap.add_argument("--sharded-mode", default="peer", choices=["peer", "nccl"], help="config4: peer-direct fused kernel or NCCL all-to-all")
    ap.add_argument("--host-fraction", type=float, default=0.0, help="config5: fraction of host RAM to pin (default 0.5)")
    ap.add_argument("--e2e-embeds-to-host", action="store_true", help="also time e2e with the embeddings copied to pinned host memory")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it so that one process drives each GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w, rank, world)
    elif w.get("tier") == "sharded":
        run_sharded(args, w, rank, local_rank, world)
    elif w.get("tier") == "host":
        run_host(args, w, rank, local_rank, world)
    else:
        run_ours(args, w, rank, local_rank, world)


if __name__ == "__main__":
    main()
