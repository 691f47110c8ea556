"""bench.py --config c4 / c5: the two remaining configurations of BASELINE.json, same JSON line shape as the headline.

c4  Boolean AND queries behind fast-field filters (BASELINE.json configs[3]): the C2 corpus plus a `lang`
    keyword column (8 values, Zipf) and a `year` i64 column (uniform 2000..2025); queries `Bool{must:[t1,t2(,t3)]}` — the
    first two or three terms of every C2 query — under `And[KeywordEq(lang), I64Range(year)]` at three selectivities (the
    survey asks for 50 / 10 / 1 %; the most frequent of 8 Zipf values covers 36.8 % of the docs, so the legs are 36.8 %,
    9.9 % and 0.94 %), and the 2-term adjacent phrase shape (slop 0, positions resident; `--no-phrases` skips it).
    Every leg is checked on the FULL batch against the oracle (`parity`).  `value` is the 10 % leg.
c5  Hybrid retrieval (configs[4]): BM25 top-1000 (k = 1001, bmw) -> exact rerank by 768-d bf16 vectors, alpha 0.5, one
    cosine clause; 12.5 M docs and 19.2 GB of vectors per GPU (100 M docs over 8 GPUs: weak scaling), every rank rescoring
    its own segment's top-k on the device (slg_rerank_batch), ONE all-gather of the reranked blocks, hybrid-order merge.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _node(op, column=-1, i=(0, 0), f=(0.0, 0.0), nc=0, v=(0, 0)):
    from searchlite_b200.engine import FILTER_DTYPE
    n = np.zeros(1, dtype=FILTER_DTYPE)
    n[0] = (op, column, i[0], i[1], f[0], f[1], nc, v[0], v[1])
    return n


class _Timer:
    """device-timed loops on the handle's stream, barrier + synchronize on both sides, max over ranks"""

    def __init__(self, torch, dist, stream, device, world, steps, warmup):
        self.torch, self.dist, self.stream, self.device, self.world, self.steps, self.warmup = torch, dist, stream, device, world, steps, warmup

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def run(self, step, wall: bool = False) -> float:
        for _ in range(max(self.warmup, 1)):
            step()
        self.barrier()
        e0, e1 = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        with self.torch.cuda.stream(self.stream):
            e0.record(self.stream)
            for _ in range(self.steps):
                step()
            e1.record(self.stream)
        self.barrier()
        ms = e0.elapsed_time(e1)
        if wall:
            ms = max(ms, 1e3 * (time.perf_counter() - w0))
        return self.max_over_ranks(ms) / self.steps


def main(args, rank: int, local_rank: int, world: int) -> None:
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    line = run_c4(args, rank, local_rank, world, device) if args.config == "c4" else run_c5(args, rank, local_rank, world, device)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
def run_c4(args, rank, local_rank, world, device):
    import torch
    import torch.distributed as dist
    from bench import ClockSampler, read_peaks
    from searchlite_b200 import GpuIndex, synth
    from searchlite_b200.engine import F_AND, F_I64_RANGE, F_KEYWORD_EQ, QueryBatch
    from searchlite_b200.shard import ShardedSearcher, shard_ranges

    k = args.limit + 1
    lo, hi = shard_ranges(args.docs, world)[rank]
    spec = synth.CorpusSpec(n_docs=hi - lo, vocab=args.vocab, seed=20260101, segment_ord=rank, doc_base=lo)
    t0 = time.time()
    seg = synth.generate_segment(spec, device)
    names, lang, year = synth.fast_fields(spec)
    seg.fast_str["lang"] = (names, lang)
    seg.fast_i64["year"] = (year, None)
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    t0 = time.time()
    gi = GpuIndex(local_rank, kernel=args.kernel)
    cols = gi.load_segment(seg)
    load_s = time.time() - t0
    n_postings = gi.segment_stats(rank)["n_postings"]

    base = synth.generate_queries(args.queries, args.vocab, seed=20260102)
    bools = []
    for q in range(base.n_queries):
        t = base.terms["term_id"][int(base.term_off[q]): int(base.term_off[q + 1])].tolist()
        bools.append({"must": t[: (2 if len(t) == 2 else 3)]})
    legs_spec = [  # (label, lang value, year range): selectivity = p(lang) x years / 26
        ("sel_37pct", "en", (2000, 2025)),
        ("sel_10pct", "en", (2000, 2006)),
        ("sel_1pct", "de", (2010, 2011)),
    ]
    stream = torch.cuda.ExternalStream(gi.stream_ptr(), device=device)
    tm = _Timer(torch, dist, stream, device, world, args.steps, args.warmup)
    searcher = ShardedSearcher(gi, args.queries, k, threshold_board=args.threshold_board) if world > 1 else None

    def make_leg(label, qb, fid_desc, filter_nodes, strings, selectivity):
        out = {"leg": label, "filter": fid_desc, "selectivity": selectivity}
        res = {}
        for execution in (args.execution, "bm25"):
            p = gi.prepare(qb, k, execution)

            def step():
                if world > 1:
                    return searcher.run(p)
                p.run(sync=True)
            c0 = gi.counters()
            ms = tm.run(step)
            c1 = gi.counters()
            p.run(sync=True)
            res[execution] = searcher.exchange_and_merge(p) if world > 1 else p.fetch()
            n_score = max(c1["score_launches"] - c0["score_launches"], 1)
            out[execution] = {"value": args.queries / (ms / 1e3), "unit": "queries/s", "ms_per_step": ms,
                              "kernel_ms": (c1["score_ms_total"] - c0["score_ms_total"]) / n_score,
                              "gpu_launches": (c1["kernel_launches"] - c0["kernel_launches"]) / (args.steps + max(args.warmup, 1))}
            p.free()
            if execution == "bm25":
                out["posting_count"] = int(gi.counters()["last_posting_count"])
            if args.execution == "bm25":
                break
        if not args.no_e2e:
            def e2e_step():
                if world == 1:
                    return gi.search_batch(qb, k, args.execution)
                p2 = gi.prepare(qb, k, args.execution)
                r = searcher.run(p2)
                p2.free()
                return r
            ms = tm.run(e2e_step, wall=True)
            cc = gi.counters()
            out["e2e"] = {"value": args.queries / (ms / 1e3), "unit": "queries/s", "ms_per_step": ms,
                          "h2d_bytes_per_step": int(cc["last_h2d_bytes"]), "d2h_bytes_per_step": int(cc["last_d2h_bytes"])}
        if args.execution != "bm25":
            a, b = res[args.execution], res["bm25"]
            out["pruned_identical_to_exhaustive"] = bool(a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes())
        return out, res[args.execution]

    sampler = ClockSampler(local_rank) if rank == 0 else None
    legs, results, filters = [], {}, {}
    for label, lang_v, (y0, y1) in legs_spec:
        prog = np.concatenate([_node(F_AND, nc=2), _node(F_KEYWORD_EQ, column=cols["lang"], v=(0, 1)),
                               _node(F_I64_RANGE, column=cols["year"], i=(y0, y1))])
        t0 = time.perf_counter()
        fid = gi.compile_filter(prog, [lang_v])
        torch.cuda.synchronize()
        comp_ms = 1e3 * (time.perf_counter() - t0)
        bits = gi.filter_bitmap(fid, rank, spec.n_docs)
        sel = float(np.unpackbits(bits.view(np.uint8)).sum()) / max(spec.n_docs, 1)
        qb = QueryBatch.from_bool([dict(b, filter_id=fid) for b in bools])
        leg, got = make_leg(label, qb, f"And[KeywordEq(lang, {lang_v!r}), I64Range(year, {y0}, {y1})]", prog, [lang_v], sel)
        leg["filter_compile_ms"] = comp_ms
        legs.append(leg)
        results[label] = got
        filters[label] = (prog, [lang_v])
    clocks = sampler.stop() if sampler else None

    # ---- the phrase shape: 2 adjacent terms, slop 0, positions resident ----
    phrase_leg = None
    if not args.no_phrases and world == 1:
        phrase_leg = _c4_phrases(args, gi, spec, seg, device, tm, k)
    del seg
    torch.cuda.empty_cache()

    # ---- CPU oracle: parity on EVERY query of every leg + the baseline timing ----
    n_cpu = args.queries if args.cpu_sample < 0 else min(args.cpu_sample, args.queries)
    cpu = None
    if rank == 0 and n_cpu > 0:
        from oracle import slo
        from tests.parity import parity_report
        slo.build()
        threads = slo.max_threads()
        oras = []
        for r in range(world):
            lo_r, hi_r = shard_ranges(args.docs, world)[r]
            sp = synth.CorpusSpec(n_docs=hi_r - lo_r, vocab=args.vocab, seed=20260101, segment_ord=r, doc_base=lo_r)
            host = synth.generate_segment(sp, device).to_host()
            torch.cuda.empty_cache()
            nm, lg, yr = synth.fast_fields(sp)
            host.fast_str["lang"] = (nm, lg)
            host.fast_i64["year"] = (yr, None)
            oras.append(slo.OracleIndex(host))
        cpu_s = 0.0
        for leg in legs:
            prog, strings = filters[leg["leg"]]
            # the oracle's column handles follow its own registration order: rewrite the program's column fields
            o_prog = prog.copy()
            o_prog["column"][1] = oras[0].columns["lang"]
            o_prog["column"][2] = oras[0].columns["year"]
            qb = QueryBatch.from_bool(bools[:n_cpu])
            t0 = time.perf_counter()
            per_seg = [o.search_batch(qb, k, "bm25_dense", filter_nodes=o_prog, strings=strings, threads=threads) for o in oras]
            dt = time.perf_counter() - t0
            if leg["leg"] == "sel_10pct":
                cpu_s = dt
            if world == 1:
                ref_h, ref_c = per_seg[0]
            else:
                ref_h, ref_c = np.zeros_like(per_seg[0][0]), np.zeros_like(per_seg[0][1])
                for qi in range(n_cpu):
                    m = slo.merge_hits([h[qi, : c[qi]] for h, c in per_seg], k)
                    ref_h[qi, : len(m)] = m
                    ref_c[qi] = len(m)
            got_h, got_c = results[leg["leg"]]
            leg["parity"] = parity_report(ref_h, ref_c, got_h[:n_cpu], got_c[:n_cpu])
            leg["parity"]["hits_per_query_mean"] = float(ref_c.mean())
            if world == 1:  # the posting scan's declared summation order (terms without a dense column first): bit for bit
                from tests.helpers import canonical_batch
                can_h, can_c = oras[0].search_batch(canonical_batch(gi, qb, rank), k, "bm25_dense", filter_nodes=o_prog, strings=strings, threads=threads)
                leg["parity"]["bit_exact_declared_order"] = parity_report(can_h, can_c, got_h[:n_cpu], got_c[:n_cpu])["bit_exact"]
        cpu = {"value": n_cpu / cpu_s if cpu_s else 0.0, "unit": "queries/s", "cores": threads, "kind": "port",
               "sample": f"{n_cpu} of the {args.queries} queries of the 10 % leg, oracle bm25_dense with the matcher and the filter evaluated per "
                         f"candidate as the reference's accept does ({threads} threads over queries), {world} segment(s)"}

    hbm_peak, peak_kind = read_peaks()
    head = next(l for l in legs if l["leg"] == "sel_10pct")
    x = head["bm25"]
    alg_bytes = 5.0 * head["posting_count"]
    achieved = alg_bytes / (x["kernel_ms"] / 1e3) / 1e9 if x["kernel_ms"] > 0 else 0.0
    line = {
        "metric": "bm25_top10_queries_per_sec", "value": head[args.execution]["value"], "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": head[args.execution]["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C4: C2 corpus ({args.docs / 1e6:g}M docs) + lang (8 values, Zipf) / year (2000..2025) columns, {args.queries} "
                               f"Bool{{must:[t1,t2(,t3)]}} queries behind And[KeywordEq(lang), I64Range(year)], top-{args.limit}; headline leg = 10 % "
                               f"selectivity; {world} doc-range segment(s)",
                   "execution": args.execution, "l2": "inputs larger than L2 (resident postings >> 126 MB); no explicit flush",
                   "postings_resident_this_rank": int(n_postings), "kernel": args.kernel},
        "e2e": head.get("e2e"), "gpu_launches": int(round(head[args.execution]["gpu_launches"] * args.steps)), "clocks": clocks,
        "setup": {"corpus_gen_s": round(gen_s, 1), "load_segment_s": round(load_s, 1), "resident_bytes": int(gi.counters()["resident_bytes"])},
        "legs": legs, "phrases": phrase_leg,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                     "peak_kind": peak_kind, "kernel": "slg_scan_kernel<false> in AND mode (execution bm25, 10 % leg): the rarest list of every query is scanned, the others are asked",
                     "kernel_ms": x["kernel_ms"], "algorithmic_bytes_per_launch": alg_bytes,
                     "note": "5 B x sum of df over the batch's must terms: what the reference reads (it scores every posting of every term and "
                             "rejects in accept: AND = OR-scan + reject, SURVEY.md §8 a13).  The engine intersects from the rarest list, so this "
                             "fraction measures the algorithmic saving, not HBM pressure"},
        "cpu_baseline": cpu,
        "parity": head.get("parity"),
    }
    gi.close()
    return line


def _c4_phrases(args, gi, spec, seg, device, tm, k, n_phr: int = 256):
    """2-term adjacent phrases sampled from the corpus; full-batch parity against the oracle searching behind an I64List
    column that lists, per doc, the phrases it holds — built from the token generator alone (no postings, no positions)"""
    import torch
    from searchlite_b200 import synth
    from searchlite_b200.engine import F_I64_RANGE, QueryBatch
    out = {"n_phrases": n_phr}
    t0 = time.perf_counter()
    positions = synth.generate_positions(spec, device)
    pos_off = torch.zeros(seg.post_tfs.shape[0] + 1, dtype=torch.int64, device=device)
    pos_off[1:] = torch.cumsum(seg.post_tfs.to(torch.int64), 0)
    torch.cuda.synchronize()
    out["positions_gen_s"] = round(time.perf_counter() - t0, 1)
    t0 = time.perf_counter()
    gi.load_positions(spec.segment_ord, seg.term_offsets, pos_off, positions)
    torch.cuda.synchronize()
    out["positions_resident_s"] = round(time.perf_counter() - t0, 2)
    out["n_positions"] = int(positions.shape[0])
    del positions, pos_off
    torch.cuda.empty_cache()
    cdf = synth.zipf_cdf(spec.vocab, spec.zipf_s).to(device)
    rng = np.random.default_rng(20260105)
    phrases, seen = [], set()
    while len(phrases) < n_phr:
        d = int(rng.integers(0, spec.n_docs))
        term, valid = synth.token_terms(spec, d, d + 1, cdf, device)
        t = term[0][valid[0]].cpu().numpy()
        p = int(rng.integers(0, len(t) - 1))
        a, b = int(t[p]), int(t[p + 1])
        if a >= 9 and b >= 9 and a != b and (a, b) not in seen:
            seen.add((a, b))
            phrases.append((a, b))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ids = gi.compile_phrases([[a, b] for a, b in phrases])
    torch.cuda.synchronize()
    out["compile_ms_per_batch"] = 1e3 * (time.perf_counter() - t0)
    qb = QueryBatch.from_term_lists([[a, b] for a, b in phrases])
    qb.filter_id = np.array(ids, dtype=np.int32)
    got = None
    for execution in (args.execution, "bm25"):
        p = gi.prepare(qb, k, execution)
        ms = tm.run(lambda: p.run(sync=True))
        p.run(sync=True)
        r = p.fetch()
        p.free()
        out[execution] = {"value": n_phr / (ms / 1e3), "unit": "queries/s", "ms_per_step": ms}
        if got is None:
            got = r
        else:
            out["pruned_identical_to_exhaustive"] = bool(got[0].tobytes() == r[0].tobytes() and got[1].tobytes() == r[1].tobytes())
        if args.execution == "bm25":
            break
    # generator-derived (doc, phrase) pairs: a at position p and b at p + 1
    V = spec.vocab
    keys = torch.tensor(sorted(a * V + b for a, b in phrases), dtype=torch.int64, device=device)
    order = {a * V + b: i for i, (a, b) in enumerate(phrases)}
    pair_doc, pair_phr = [], []
    chunk = 1 << 18
    for d0 in range(0, spec.n_docs, chunk):
        d1 = min(spec.n_docs, d0 + chunk)
        term, valid = synth.token_terms(spec, d0, d1, cdf, device)
        pk = term[:, :-1] * V + term[:, 1:]
        hit = torch.isin(pk, keys) & valid[:, 1:]
        rows, colsx = torch.nonzero(hit, as_tuple=True)
        pair_doc.append((rows + d0).cpu().numpy())
        pair_phr.append(pk[rows, colsx].cpu().numpy())
        del term, valid, pk, hit
    docs = np.concatenate(pair_doc)
    phr = np.array([order[int(v)] for v in np.concatenate(pair_phr)], dtype=np.int64)
    pairs = np.unique(np.stack([docs, phr], axis=1), axis=0)  # (doc, phrase) sorted by doc
    offs = np.zeros(spec.n_docs + 1, dtype=np.uint32)
    np.add.at(offs, pairs[:, 0] + 1, 1)
    offs = np.cumsum(offs, dtype=np.uint64).astype(np.uint32)
    out["docs_holding_a_phrase"] = int(len(np.unique(pairs[:, 0])))
    if args.cpu_sample != 0:
        from oracle import slo
        from tests.parity import parity_report
        slo.build()
        host = seg.to_host()
        host.fast_i64_list["phrases"] = (offs, pairs[:, 1].astype(np.int64))
        ora = slo.OracleIndex(host)
        ref_h = np.zeros_like(got[0])
        ref_c = np.zeros_like(got[1])
        t0 = time.perf_counter()
        for i in range(n_phr):
            prog = _node(F_I64_RANGE, column=ora.columns["phrases"], i=(i, i))
            h, c = ora.search_batch(qb.subset(i, i + 1), k, "bm25_dense", filter_nodes=prog)
            ref_h[i], ref_c[i] = h[0], c[0]
        out["oracle_s"] = round(time.perf_counter() - t0, 1)
        out["parity"] = parity_report(ref_h, ref_c, got[0], got[1])
        out["parity"]["note"] = ("every phrase query against the oracle's search behind an I64List column (any-value semantics) that lists each "
                                 "doc's phrases, derived from the token generator alone")
        del ora, host
    return out


# ---------------------------------------------------------------------------------------------------------------------
def c5_vectors(torch, n_docs: int, dim: int, seed: int, device, chunk: int = 1 << 19):
    """(offsets int32 [n_docs], rows bf16 [n_rows, dim]) born on the device: unit-norm N(0,1) rows; one doc in 64 has none"""
    from searchlite_b200 import synth
    d = torch.arange(n_docs, dtype=torch.int64, device=device)
    has = (synth._lsr(synth.hash2(seed ^ 0xC5, d), 3) % 64) != 0
    row = torch.cumsum(has.to(torch.int64), 0) - 1
    offsets = torch.where(has, row, torch.full_like(row, -1)).to(torch.int32)  # (-1 = u32::MAX, "no vector")
    n_rows = int(has.sum().item())
    rows = torch.empty((n_rows, dim), dtype=torch.bfloat16, device=device)
    g = torch.Generator(device=device)
    for r0 in range(0, n_rows, chunk):
        r1 = min(n_rows, r0 + chunk)
        g.manual_seed(seed * 1_000_003 + r0)
        x = torch.randn((r1 - r0, dim), generator=g, device=device, dtype=torch.float32)
        x /= x.norm(dim=1, keepdim=True)
        rows[r0:r1] = x.to(torch.bfloat16)
        del x
    return offsets, rows


def run_c5(args, rank, local_rank, world, device):
    import torch
    import torch.distributed as dist
    from bench import ClockSampler, read_peaks
    from searchlite_b200 import GpuIndex, synth
    from searchlite_b200.shard import gather_blocks

    dim = 768
    per = args.docs if args.docs else 12_500_000  # docs PER GPU (weak scaling: 100 M docs over 8 GPUs)
    cand = args.candidates
    k = cand + 1  # top_k = candidate_size + 1, api/reader.rs:2611-2616
    spec = synth.CorpusSpec(n_docs=per, vocab=args.vocab, seed=20260101, segment_ord=rank, doc_base=rank * per)
    t0 = time.time()
    seg = synth.generate_segment(spec, device)
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    t0 = time.time()
    gi = GpuIndex(local_rank, kernel=args.kernel)
    gi.load_segment(seg)
    del seg
    torch.cuda.empty_cache()
    load_s = time.time() - t0
    t0 = time.time()
    offsets, rows = c5_vectors(torch, per, dim, 20260106 + rank, device)
    gi.load_vectors(rank, offsets, rows)
    torch.cuda.synchronize()
    vec_s = time.time() - t0
    n_postings = gi.segment_stats(rank)["n_postings"]

    qb = synth.generate_queries(args.queries, args.vocab, seed=20260102)
    rng = np.random.default_rng(20260107)
    qv = rng.standard_normal((args.queries, dim)).astype(np.float32)
    qv /= np.linalg.norm(qv, axis=1, keepdims=True)
    qv_pinned = torch.from_numpy(qv).pin_memory()
    qv_dev = qv_pinned.to(device)
    alpha = 0.5
    stream = torch.cuda.ExternalStream(gi.stream_ptr(), device=device)
    tm = _Timer(torch, dist, stream, device, world, args.steps, args.warmup)
    Q = args.queries

    def as_tensor(ptr, nbytes):
        class _A:
            pass
        a = _A()
        a.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}
        return torch.as_tensor(a, device=device)

    def finish(p):
        """exchange + merge (N > 1) or fetch: host hits, counts, vector scores"""
        if world == 1:
            h, c = p.fetch()
            return h, c, p.fetch_vector_scores()
        ptr, nbytes = p.packed_results()
        blocks = gather_blocks(as_tensor(ptr, nbytes))
        return gi.merge_gathered_hybrid(blocks.data_ptr(), world, Q, k)

    p = gi.prepare(qb, k, args.execution)

    def step_resident():
        with torch.cuda.stream(stream):
            p.run(sync=False)
            p.rerank([(qv_dev, alpha, 1.0, "cosine")], sync=(world == 1))
            if world > 1:
                return finish(p)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    c0 = gi.counters()
    ms_step = tm.run(step_resident)
    c1 = gi.counters()
    clocks = sampler.stop() if sampler else None
    launches = (c1["kernel_launches"] - c0["kernel_launches"]) / (args.steps + max(args.warmup, 1))
    # stage times on this rank (CUDA events on the handle's stream)
    p.run(sync=True)
    bm25_ms = gi.counters()["last_batch_ms"]
    bm25_h, bm25_c = p.fetch()
    p.rerank([(qv_dev, alpha, 1.0, "cosine")], sync=True)
    rerank_ms = gi.counters()["last_rerank_ms"]
    with torch.cuda.stream(stream):
        got_h, got_c, got_vs = finish(p)
    p.free()

    e2e = None
    if not args.no_e2e:
        def step_e2e():
            with torch.cuda.stream(stream):
                p2 = gi.prepare(qb, k, args.execution)  # host query structs -> packed H2D
                p2.run(sync=False)
                p2.rerank([(qv_pinned, alpha, 1.0, "cosine")], sync=False)  # query vectors from pinned host memory
                r = finish(p2)
                p2.free()
                return r
        ms = tm.run(step_e2e, wall=True)
        e2e = {"value": Q * 1.0 / (ms / 1e3), "unit": "queries/s", "ms_per_step": ms,
               "h2d_bytes_per_step": int(Q * dim * 4 + 300_000), "d2h_bytes_per_step": int(Q * k * 16 + Q * 4)}

    # ---- parity + CPU baseline: every rank checks a sample on its own segment, rank 0 checks the merged ranking ----
    n_cpu = min(args.queries, 32 if args.cpu_sample < 0 else args.cpu_sample)
    parity = cpu = None
    if n_cpu > 0:
        from oracle import slo
        from tests.helpers import canonical_batch
        from tests.parity import parity_report
        slo.build()
        threads = max(1, slo.max_threads() // world)
        host = synth.generate_segment(spec, device).to_host()
        torch.cuda.empty_cache()
        ora = slo.OracleIndex(host)
        sub = qb.subset(0, n_cpu)
        t0 = time.perf_counter()
        o_h, o_c = ora.search_batch(canonical_batch(gi, sub, rank), k, "bm25_dense", threads=threads)
        t_bm25 = time.perf_counter() - t0
        local_bm25 = parity_report(o_h, o_c, bm25_h[:n_cpu], bm25_c[:n_cpu]) if world == 1 else None
        # the rows of the sample's candidates, as stored (bf16 -> f32 is exact), in a compact store
        docs = np.unique(o_h["doc_id"][o_h["segment_ord"] == rank])
        off_host = offsets.cpu().numpy().view(np.uint32)
        have = docs[off_host[docs] != 0xFFFFFFFF]
        take = torch.from_numpy(off_host[have].astype(np.int64)).to(device)
        compact_rows = rows[take].to(torch.float32).cpu().numpy()
        compact_off = np.full(per, 0xFFFFFFFF, dtype=np.uint32)
        compact_off[have] = np.arange(len(have), dtype=np.uint32)
        t0 = time.perf_counter()
        r_h, r_c, r_vs = slo.rerank_batch(o_h, o_c, [(rank, compact_off, compact_rows)], [(qv[:n_cpu], alpha, 1.0, "cosine")], threads=threads)
        t_rr = time.perf_counter() - t0
        if world > 1:
            gathered = [None] * world if rank == 0 else None
            dist.gather_object((r_h, r_c), gathered, dst=0)
            if rank == 0:
                ref_h, ref_c = np.zeros_like(r_h), np.zeros_like(r_c)
                for qi in range(n_cpu):
                    m = slo.merge_hits([h[qi, : c[qi]] for h, c in gathered], k)
                    ref_h[qi, : len(m)] = m
                    ref_c[qi] = len(m)
        else:
            ref_h, ref_c = r_h, r_c
        if rank == 0:
            parity = parity_report(ref_h, ref_c, got_h[:n_cpu], got_c[:n_cpu])
            parity["note"] = (f"{n_cpu} queries: the oracle's BM25 top-{cand + 1} of every segment (declared summation order), rescored with "
                              "compute_hybrid_score on the candidates' stored rows (exact similarities, sequential f32 fold), merged in SortKey order")
            if local_bm25:
                parity["bm25_stage"] = local_bm25
            cpu = {"value": n_cpu / (t_bm25 + t_rr), "unit": "queries/s", "cores": threads, "kind": "port",
                   "sample": f"{n_cpu} of the {Q} queries on this rank's {per / 1e6:g} M-doc segment: oracle bm25_dense top-{k} ({t_bm25:.2f} s) + exact "
                             f"rerank of the candidates ({t_rr:.2f} s), {threads} threads over queries"}
        del ora, host

    hbm_peak, peak_kind = read_peaks()
    n_cands = int(bm25_c.sum())
    gather_bytes = float(n_cands) * dim * 2
    achieved = gather_bytes / (rerank_ms / 1e3) / 1e9 if rerank_ms > 0 else 0.0
    line = {
        "metric": "bm25_top10_queries_per_sec", "value": Q * 1.0 / (ms_step / 1e3), "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 (bf16 rows)", "data": "synthetic",
        "config": {"workload": f"C5: hybrid — BM25 top-{cand} (k={k}, {args.execution}) over {per / 1e6:g}M docs per GPU x {world} GPU(s) "
                               f"({per * world / 1e6:g}M docs), {Q} OR queries of 2-5 terms, exact rerank by {dim}-d bf16 rows (cosine, alpha {alpha}, "
                               f"one clause), one all-gather of the reranked blocks + hybrid-order merge",
                   "execution": args.execution, "l2": "inputs larger than L2 (postings and vector rows >> 126 MB); no explicit flush",
                   "postings_resident_this_rank": int(n_postings), "vector_bytes_this_rank": int(rows.numel() * 2), "kernel": args.kernel},
        "e2e": e2e, "gpu_launches": int(round(launches * args.steps)), "clocks": clocks,
        "setup": {"corpus_gen_s": round(gen_s, 1), "load_segment_s": round(load_s, 1), "vectors_s": round(vec_s, 1),
                  "resident_bytes": int(gi.counters()["resident_bytes"])},
        "stages": {"bm25_top_k_ms": bm25_ms, "rerank_ms": rerank_ms, "candidates": n_cands},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
                     "peak_kind": peak_kind, "kernel": "slg_rerank_scores_kernel<1, bf16> + slg_rerank_sort_kernel (this rank)",
                     "kernel_ms": rerank_ms, "algorithmic_bytes_per_launch": gather_bytes,
                     "note": "candidates x dim x 2 B of gathered vector rows (SURVEY.md §8d: bytes = Q*C*dim*2); flops = 2*Q*C*dim = "
                             f"{2.0 * n_cands * dim / 1e9:.2f} GFLOP per batch — a per-query matrix-vector product with no operand reuse, so the row "
                             "gather, not the FMA or tensor pipe, bounds it"},
        "cpu_baseline": cpu, "parity": parity,
    }
    gi.close()
    return line
