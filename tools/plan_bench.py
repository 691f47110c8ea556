"""Bounded timing of ScorePlan queries at C2 size (a parity-test shape, not a bench line):
the C2 corpus and its 4096 two-to-five-term OR queries, each query's terms dealt onto two leaves under
DisMax(tie_breaker 0.3) — the shape of a `multi_match best_fields` — on the warp kernel with accumulator
planes (the automatic choice) and on the CTA-per-item kernel with planes, next to the same queries without a plan.
A sample of the plan results is compared with the oracle.  Usage: python tools/plan_bench.py [n_docs]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from searchlite_b200 import GpuIndex, synth  # noqa: E402

n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000


def timed(fn, warm=1, steps=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / steps


spec = synth.CorpusSpec(n_docs=n_docs, vocab=1_000_000, seed=20260101)
seg = synth.generate_segment(spec, "cuda:0")
qb = synth.generate_queries(4096, spec.vocab, seed=20260102)
plan_qb = qb.subset(0, qb.n_queries)
plan_qb.terms["leaf"] = plan_qb.terms["leaf"] & 1  # terms alternate between two leaves
plan_qb.set_plans([("dismax", [("leaf", 0), ("leaf", 1)], 0.3)] * qb.n_queries)
plan_qb.leaf_count[:] = 2
plan_qb._structs = None

for label, kernel, batch in (("no plan, automatic kernel", "auto", qb), ("no plan, warp kernel (query order)", "warp", qb),
                             ("DisMax over 2 leaves, warp kernel with planes (automatic)", "auto", plan_qb),
                             ("DisMax over 2 leaves, CTA-per-item kernel with planes", "cta", plan_qb)):
    gi = GpuIndex(0, kernel=kernel)
    gi.load_segment(seg)
    for exe in ("bm25", "bmw"):
        p = gi.prepare(batch, 11, exe)
        ms = timed(lambda: p.run(sync=True))
        print(f"{label}, {exe}: {ms:.1f} ms/batch, {4096 / ms * 1e3:.0f} q/s", flush=True)
        p.free()
    if batch is plan_qb and kernel == "auto":
        from oracle import slo  # checker only
        from tests.parity import parity_report
        host = seg.to_host() if hasattr(seg, "to_host") else seg
        sample = plan_qb.subset(0, 32)
        t0 = time.perf_counter()
        ref = slo.OracleIndex(host).search_batch(sample, 11, "bm25", threads=slo.max_threads())
        got = gi.search_batch(sample, 11, "bm25")
        print(f"parity of 32 plan queries vs oracle ({time.perf_counter() - t0:.1f} s): {parity_report(*ref, *got)}", flush=True)
    gi.close()
    torch.cuda.empty_cache()
