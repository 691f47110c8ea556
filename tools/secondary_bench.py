"""Bounded timings of the secondary configurations of BASELINE.json (parity-test cases, not bench lines):
C3-like  8.84 M short docs, top-1000 (k = 1001), bmw — the CTA-per-item kernel
C4-like  C2 corpus, OR queries behind a root filter And[KeywordEq(lang), I64Range(year)] at ~10 % selectivity
C5-like  exact rerank of 1000 candidates per query by 768-d vectors (bf16 rows), 256 queries
Usage: python tools/secondary_bench.py [c3] [c4] [c5]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from searchlite_b200 import GpuIndex, synth  # noqa: E402
from searchlite_b200.engine import FILTER_DTYPE, F_AND, F_I64_RANGE, F_KEYWORD_EQ, HIT_DTYPE  # noqa: E402

which = set(sys.argv[1:]) or {"c3", "c4", "c5"}


def timed(fn, warm=1, steps=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / steps


def node(op, column=-1, i=(0, 0), f=(0.0, 0.0), nc=0, v=(0, 0)):
    n = np.zeros(1, dtype=FILTER_DTYPE)
    n[0] = (op, column, i[0], i[1], f[0], f[1], nc, v[0], v[1])
    return n


if "c3" in which:
    spec = synth.CorpusSpec(n_docs=8_841_823, vocab=1_000_000, seed=20260103, len_lo=8, len_hi=104)
    seg = synth.generate_segment(spec, "cuda:0")
    qb = synth.generate_queries(1024, spec.vocab, seed=20260104)
    gi = GpuIndex(0)
    gi.load_segment(seg)
    del seg
    p = gi.prepare(qb, 1001, "bmw")
    ms = timed(lambda: p.run(sync=True))
    print(f"C3-like: 8.84M docs (8..104 tokens), 1024 OR queries, top-1000 bmw: {ms:.1f} ms/batch, {1024 / ms * 1e3:.0f} q/s", flush=True)
    p.free()
    gi.close()
    torch.cuda.empty_cache()

if "c4" in which or "c5" in which:
    spec = synth.CorpusSpec(n_docs=10_000_000, vocab=1_000_000, seed=20260101)
    seg = synth.generate_segment(spec, "cuda:0")
    qb = synth.generate_queries(4096, spec.vocab, seed=20260102)
    gi = GpuIndex(0)
    gi.load_segment(seg)
    del seg
    torch.cuda.empty_cache()

if "c4" in which:
    names, lang, year = synth.fast_fields(spec)
    h_lang = gi._check(gi.lib.slg_add_str_column(gi.handle, 0, (__import__("ctypes").c_char_p * len(names))(*[n.encode() for n in names]), len(names),
                                                 lang.ctypes.data_as(__import__("ctypes").c_void_p)))
    h_year = gi._check(gi.lib.slg_add_i64_column(gi.handle, 0, year.ctypes.data_as(__import__("ctypes").c_void_p), None))
    prog = np.concatenate([node(F_AND, nc=2), node(F_KEYWORD_EQ, column=h_lang, v=(0, 1)), node(F_I64_RANGE, column=h_year, i=(2000, 2006))])
    t0 = time.perf_counter()
    fid = gi.compile_filter(prog, ["es"])
    torch.cuda.synchronize()
    comp_ms = 1e3 * (time.perf_counter() - t0)
    bits = gi.filter_bitmap(fid, 0, spec.n_docs)
    sel = float(np.unpackbits(bits.view(np.uint8)).sum()) / spec.n_docs
    qf = qb.subset(0, qb.n_queries)
    qf.filter_id = np.full(qb.n_queries, fid, dtype=np.int32)
    p = gi.prepare(qf, 11, "bm25")
    ms = timed(lambda: p.run(sync=True))
    print(f"C4-like: C2 corpus, 4096 OR queries behind And[lang == es, 2000 <= year <= 2006] (selectivity {sel:.3f}): filter compile "
          f"{comp_ms:.1f} ms, {ms:.1f} ms/batch, {4096 / ms * 1e3:.0f} q/s", flush=True)
    p.free()

if "c5" in which:
    nq, nc, dim = 256, 1000, 768
    n_rows = 2_000_000  # vectors for the first 2 M docs (3 GB as bf16); candidates are drawn from them
    rng = np.random.default_rng(5)
    offsets = np.full(spec.n_docs, 0xFFFFFFFF, dtype=np.uint32)
    offsets[:n_rows] = np.arange(n_rows, dtype=np.uint32)
    vals = rng.standard_normal((n_rows, dim), dtype=np.float32)
    vals /= np.linalg.norm(vals, axis=1, keepdims=True)
    gi.load_vectors(0, offsets, vals, store_bf16=True)
    del vals
    qv = rng.standard_normal((nq, dim)).astype(np.float32)
    qv /= np.linalg.norm(qv, axis=1, keepdims=True)
    cands = np.zeros((nq, nc), dtype=HIT_DTYPE)
    cands["doc_id"] = rng.integers(0, n_rows, size=(nq, nc), dtype=np.uint32)
    cands["score"] = rng.random((nq, nc), dtype=np.float32) * 10
    cc = np.full(nq, nc, dtype=np.uint32)
    ms = timed(lambda: gi.rerank(qv, cands, cc, 0.5, "cosine"))
    gb = nq * nc * dim * 2 / 1e9
    print(f"C5-like: rerank {nq} queries x {nc} candidates x {dim}-d bf16 rows (host in/out): {ms:.1f} ms/batch, {nq / ms * 1e3:.0f} q/s, "
          f"gather {gb / (ms / 1e3):.0f} GB/s", flush=True)
