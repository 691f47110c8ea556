#!/usr/bin/env python
"""bench.py — BM25 top-10 queries/sec at 10 M docs on 1/2/4/8 B200 (BASELINE.json metric).

Workload (config C2 of BASELINE.md §4): synthetic 10 M-doc Zipfian corpus (uniform 100..300 tokens
per doc, 1 M-term vocabulary, seed 20260101), a batch of 4096 two-to-five-term OR queries
(seed 20260102), top-10 (internal k = 11), k1 = 0.9, b = 0.4.  The corpus is generated on the GPU
(searchlite_b200.synth) and loaded once into the engine's HBM-resident layout; a "step" is one
pass of the whole 4096-query batch.

  value      queries/sec with the prepared batch already resident in HBM (slg_batch_run)
  e2e        queries/sec through slg_search_batch with HOST query structs in and HOST hits out
             (H2D of the packed batch + all kernels + D2H of the hits inside the timed region)
  roofline   scoring kernel: algorithmic posting bytes (5 B x sum of df over query terms) / its
             CUDA-event time, against the measured HBM copy bandwidth
  cpu_baseline  the oracle ("port" of the reference CPU path) on a bounded sample of the same
             queries on the host cores, with a parity check of the GPU results against it

N > 1 (torchrun): the corpus is split into N contiguous doc-range segments, one per rank; every
rank scores all queries on its segment, one NCCL all-gather exchanges the local top-k and every
rank merges (strong scaling: total corpus fixed).

`--impl reference` times the reference's own CPU algorithm (oracle restatement, faithful mode:
per-query varint decode + doc-length vector + WAND, the reference's default execution) on the
host cores; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--docs", type=int, default=10_000_000)
    ap.add_argument("--vocab", type=int, default=1_000_000)
    ap.add_argument("--queries", type=int, default=4096)
    ap.add_argument("--limit", type=int, default=10)
    ap.add_argument("--execution", default="bm25", choices=["bm25", "wand", "bmw"])
    ap.add_argument("--tile-docs", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--sub-docs", type=int, default=0)
    ap.add_argument("--kernel", default="auto", choices=["auto", "cta", "warp", "reg", "warp-inplace", "auto-inplace"])
    ap.add_argument("--option", action="append", default=[], help="engine residency option name=value (slg_set_option)")
    ap.add_argument("--cpu-sample", type=int, default=256, help="queries in the CPU baseline sample (0 = skip)")
    ap.add_argument("--ref-sample", type=int, default=64, help="queries per step of --impl reference")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-pruned", action="store_true", help="skip the extra timing of the block-max pruned execution")
    return ap.parse_args()


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def build_corpus(args, rank: int, world: int, device):
    """this rank's segment of the C2 corpus (contiguous doc range), generated on `device`"""
    from searchlite_b200 import synth
    from searchlite_b200.shard import shard_ranges
    lo, hi = shard_ranges(args.docs, world)[rank]
    spec = synth.CorpusSpec(n_docs=hi - lo, vocab=args.vocab, seed=20260101, segment_ord=rank, doc_base=lo)
    seg = synth.generate_segment(spec, device)
    qb = synth.generate_queries(args.queries, args.vocab, seed=20260102)
    return seg, qb


def workload_name(args, world: int) -> str:
    return (f"C2: synthetic {args.docs / 1e6:g}M-doc Zipf(s=1) corpus, uniform 100..300 tokens/doc, {args.vocab / 1e6:g}M-term vocab, "
            f"{args.queries} OR queries of 2-5 terms, top-{args.limit} (k={args.limit + 1}), k1=0.9 b=0.4, "
            f"{world} doc-range segment(s)")


def run_reference(args, rank: int, world: int):
    """the reference's CPU algorithm on the host cores (oracle port, faithful per-query decode + WAND)"""
    if rank != 0:
        return
    import torch
    from oracle import slo
    slo.build()
    t0 = time.time()
    device = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    # the reference arm searches the whole corpus as ONE segment per rank-equivalent; at N>1 it still
    # times the same bounded sample over the full corpus split into `world` segments, sequentially,
    # as IndexReader::search does (api/reader.rs:2670)
    from searchlite_b200 import synth
    from searchlite_b200.shard import shard_ranges
    oracles = []
    for r, (lo, hi) in enumerate(shard_ranges(args.docs, world)):
        spec = synth.CorpusSpec(n_docs=hi - lo, vocab=args.vocab, seed=20260101, segment_ord=r, doc_base=lo)
        seg = synth.generate_segment(spec, device).to_host()
        if device.type == "cuda":
            torch.cuda.empty_cache()
        o = slo.OracleIndex(seg)
        o.build_post_image()
        oracles.append(o)
    qb = synth.generate_queries(args.queries, args.vocab, seed=20260102)
    k = args.limit + 1
    threads = slo.max_threads()
    sample = min(args.ref_sample, args.queries)
    setup_s = time.time() - t0
    times = []
    for step in range(args.warmup + args.steps):
        lo = (step * sample) % max(1, args.queries - sample + 1)
        sub = qb.subset(lo, lo + sample)
        t = time.perf_counter()
        lists = [o.search_batch(sub, k, "wand", faithful=True, threads=threads) for o in oracles]
        if len(oracles) > 1:
            for qi in range(sample):
                slo.merge_hits([h[qi, : c[qi]] for h, c in lists], k)
        dt = time.perf_counter() - t
        if step >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    qps = sample / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "bm25_top10_queries_per_sec", "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, world), "execution": "wand (reference default)",
                   "step": f"{sample} queries per step (bounded sample of the 4096-query batch)"},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} queries/step, faithful mode: per-query varint decode of every term list + per-query "
                                   f"doc-length vector + WAND (oracle restatement; the Rust reference cannot be built here)"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "setup_s": round(setup_s, 1),
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, max(world, args.gpus if world == 1 else world))
        return

    import torch
    import torch.distributed as dist
    from searchlite_b200 import GpuIndex
    from searchlite_b200.shard import ShardedSearcher

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    k = args.limit + 1

    t0 = time.time()
    seg, qb = build_corpus(args, rank, world, device)
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    t0 = time.time()
    options = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in args.option}
    gi = GpuIndex(local_rank, tile_docs=args.tile_docs, ctas_per_sm=args.ctas_per_sm, sub_docs=args.sub_docs, kernel=args.kernel,
                  options=options)
    gi.load_segment(seg)
    load_s = time.time() - t0
    n_postings = gi.segment_stats(rank)["n_postings"]
    host_seg = None
    if rank == 0 and world == 1 and args.cpu_sample > 0:
        host_seg = seg.to_host()
    del seg
    torch.cuda.empty_cache()

    stream = torch.cuda.ExternalStream(gi.stream_ptr(), device=device)
    searcher = ShardedSearcher(gi, args.queries, k) if world > 1 else None
    prepared = gi.prepare(qb, k, args.execution)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        prepared.run(sync=True)
        if world > 1:
            return searcher.exchange_and_merge(prepared)
        return None

    # ---- value: batch resident in HBM ----
    for _ in range(args.warmup):
        step_resident()
    barrier()
    c0 = gi.counters()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            step_resident()
        ev1.record(stream)
    barrier()
    clocks = sampler.stop() if sampler else None
    c1 = gi.counters()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    score_launches = c1["score_launches"] - c0["score_launches"]
    score_ms = (c1["score_ms_total"] - c0["score_ms_total"]) / max(score_launches, 1)
    launches = (c1["kernel_launches"] - c0["kernel_launches"])
    posting_count = c1["last_posting_count"]

    # ---- e2e: host buffers in and out through the C ABI ----
    e2e = None
    if not args.no_e2e:
        def step_e2e():
            if world == 1:
                return gi.search_batch(qb, k, args.execution)
            p = gi.prepare(qb, k, args.execution)
            p.run(sync=True)
            out = searcher.exchange_and_merge(p)
            p.free()
            return out
        for _ in range(max(1, args.warmup)):
            step_e2e()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(args.steps):
                last = step_e2e()
            e1.record(stream)
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - w0)
        dev_ms = e0.elapsed_time(e1)
        te = torch.tensor([max(dev_ms, wall_ms)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        cc = gi.counters()
        e2e_ms = float(te.item()) / args.steps
        e2e = {"value": args.queries / (e2e_ms / 1e3), "unit": "queries/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(cc["last_h2d_bytes"]), "d2h_bytes_per_step": int(cc["last_d2h_bytes"])}

    # results of the resident path for the parity check
    prepared.run(sync=True)
    got_h, got_c = (searcher.exchange_and_merge(prepared) if world > 1 else prepared.fetch())

    # ---- the same batch with block-max pruning (configs[1] names it): exact result, fewer postings scored ----
    pruned = None
    if args.execution == "bm25" and not args.no_pruned:
        pp = gi.prepare(qb, k, "bmw")

        def step_pruned():
            pp.run(sync=True)
            return searcher.exchange_and_merge(pp) if world > 1 else None
        for _ in range(args.warmup):
            step_pruned()
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            p0.record(stream)
            for _ in range(args.steps):
                step_pruned()
            p1.record(stream)
        barrier()
        tp = torch.tensor([p0.elapsed_time(p1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        pr_ms = float(tp.item()) / args.steps
        pp.run(sync=True)
        pr_h, pr_c = (searcher.exchange_and_merge(pp) if world > 1 else pp.fetch())
        pruned = {"execution": "bmw", "value": args.queries / (pr_ms / 1e3), "unit": "queries/s", "ms_per_step": pr_ms,
                  "speedup_vs_exhaustive": ms_step / pr_ms,
                  "identical_to_exhaustive": bool(pr_h.tobytes() == got_h.tobytes() and pr_c.tobytes() == got_c.tobytes())}
        if world == 1 and not args.no_e2e:
            for _ in range(max(1, args.warmup)):
                gi.search_batch(qb, k, "bmw")
            torch.cuda.synchronize()
            w0 = time.perf_counter()
            for _ in range(args.steps):
                gi.search_batch(qb, k, "bmw")
            torch.cuda.synchronize()
            pe_ms = 1e3 * (time.perf_counter() - w0) / args.steps
            pruned["e2e"] = {"value": args.queries / (pe_ms / 1e3), "unit": "queries/s", "ms_per_step": pe_ms}
        pp.free()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    hbm_peak, peak_kind = read_peaks()
    alg_bytes = 5.0 * posting_count
    achieved = alg_bytes / (score_ms / 1e3) / 1e9 if score_ms > 0 else 0.0
    line = {
        "metric": "bm25_top10_queries_per_sec", "value": args.queries / (ms_step / 1e3), "unit": "queries/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, world), "execution": args.execution,
                   "l2": "inputs larger than L2 (resident postings >> 126 MB); no explicit flush",
                   "postings_resident_this_rank": int(n_postings), "kernel": args.kernel, "options": options},
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": None, "peak_kind": peak_kind,
                     "kernel": {"auto": "slg_score_warp_kernel<COLS>", "reg": "slg_score_warp_kernel<COLS> / slg_score_sweep_kernel", "cta": "slg_score_tiles_kernel"}.get(args.kernel, "slg_score_warp_kernel"),
                     "kernel_ms": score_ms, "algorithmic_bytes_per_launch": alg_bytes,
                     "note": "5 B x sum of df over the batch's query terms (this rank's segment)"},
        "pruned": pruned,
        "setup": {"corpus_gen_s": round(gen_s, 1), "load_segment_s": round(load_s, 1), "resident_bytes": int(c1["resident_bytes"])},
    }
    if (world == 1 and args.execution == "bm25" and args.kernel in ("auto", "warp") and args.docs == 10_000_000
            and args.queries == 4096 and not options and not args.sub_docs):
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this workload, one `ncu --set full` capture
        if args.kernel == "auto":
            line["roofline"]["traffic"] = 7.356216e9 + 13.366272e6
            line["roofline"]["traffic_source"] = "profiles/r1_v8_warp_cols_kernel_summary.txt"
        else:
            line["roofline"]["traffic"] = 5.608199e9 + 13.091840e6
            line["roofline"]["traffic_source"] = "profiles/r1_v7_warp_kernel_summary.txt"

    # ---- CPU baseline + parity on a bounded sample (rank 0, N = 1) ----
    if host_seg is not None:
        from oracle import slo
        from tests.parity import parity_report
        slo.build()
        ora = slo.OracleIndex(host_seg)
        threads = slo.max_threads()
        n = min(args.cpu_sample, args.queries)
        sub = qb.subset(0, n)
        tcpu = time.perf_counter()
        ref_h, ref_c = ora.search_batch(sub, k, "bm25_dense", threads=threads)
        cpu_s = time.perf_counter() - tcpu
        line["cpu_baseline"] = {"value": n / cpu_s, "unit": "queries/s", "cores": threads, "kind": "port",
                                "sample": f"first {n} of the {args.queries} queries, oracle bm25_dense (pre-decoded postings, "
                                          f"{threads} threads over queries; the 'fair' port of BASELINE.md §3)"}
        rep = parity_report(ref_h, ref_c, got_h[:n], got_c[:n])
        # the automatic kernel sums column terms first (include/searchlite_gpu.h): bit-exactness is checked against
        # the oracle run on that permutation of each query, the 1e-5 rule against the reference (query) order
        from tests.helpers import canonical_batch
        can_h, can_c = ora.search_batch(canonical_batch(gi, sub), k, "bm25_dense", threads=threads)
        rep["bit_exact_declared_order"] = parity_report(can_h, can_c, got_h[:n], got_c[:n])["bit_exact"]
        rep["note"] = ("bit_exact / within_rule: against the oracle in the reference's (query) summation order; "
                       "bit_exact_declared_order: against the oracle on the kernel's declared term order")
        line["parity"] = rep
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
