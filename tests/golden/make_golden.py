#!/usr/bin/env python
"""Generates the golden fixtures in this directory (run from the repo root:
`python tests/golden/make_golden.py`).  TEST INFRASTRUCTURE.

The reference (Rust) cannot be built in this image and its tests hold no absolute BM25 numbers
(SURVEY.md §4), so the fixtures are produced by tests/pyref.py — an independent numpy-float32
restatement written from the reference text — on inputs that include every literal the reference's
own tests use for this path:

  bm25_scalar.json     query/bm25.rs:13-19 literals + query/wand.rs:1014-1021 literals + seeded draws
  wand_literal.json    query/wand.rs:969-1011 (two literal posting lists, k = 2) and :952-966 (tie-break)
  codec.json           util/varint.rs:55-63 values, index/postings.rs:280-310 literal list
  small_topk.json      120-doc seeded corpus, 12 queries, exhaustive top-k (ids + score bits)
  vectors.json         vectors/mod.rs:98-129 + api/reader.rs:218-254 literals from tests/vector_search.rs
"""
from __future__ import annotations

import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests import pyref  # noqa: E402

f32 = np.float32


def bits(x) -> int:
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f, indent=1, sort_keys=True)
        f.write("\n")


def gen_bm25_scalar():
    rng = np.random.default_rng(20260103)
    rows = [
        # query/bm25.rs:13-19
        dict(tf=3.0, df=5.0, doc_len=100.0, avgdl=120.0, docs=1000.0, k1=1.2, b=0.75, weight=1.0),
        dict(tf=1.0, df=1.0, doc_len=0.0, avgdl=0.0, docs=10.0, k1=1.2, b=0.75, weight=1.0),
        # query/wand.rs:1014-1021
        dict(tf=2.0, df=1.0, doc_len=5.0, avgdl=10.0, docs=100.0, k1=1.2, b=0.75, weight=1.0),
        dict(tf=2.0, df=1.0, doc_len=100.0, avgdl=10.0, docs=100.0, k1=1.2, b=0.75, weight=1.0),
        # idf floor: df > N/2 makes ln(.) negative -> max(.,0)+1 == 1
        dict(tf=1.0, df=900.0, doc_len=10.0, avgdl=10.0, docs=1000.0, k1=0.9, b=0.4, weight=1.0),
        # doc_len <= 0 falls back to max(avgdl, tf) (wand.rs:279-283)
        dict(tf=7.0, df=3.0, doc_len=0.0, avgdl=4.0, docs=50.0, k1=0.9, b=0.4, weight=2.0),
        dict(tf=2.0, df=3.0, doc_len=-1.0, avgdl=4.0, docs=50.0, k1=0.9, b=0.4, weight=0.5),
    ]
    for _ in range(249):
        docs = float(rng.integers(10, 10_000_000))
        df = float(rng.integers(1, int(docs) + 1))
        avgdl = float(f32(rng.uniform(5.0, 400.0)))
        rows.append(dict(tf=float(rng.integers(1, 40)), df=df, doc_len=float(rng.integers(1, 600)), avgdl=avgdl,
                         docs=docs, k1=0.9 if rng.random() < 0.5 else 1.2, b=0.4 if rng.random() < 0.5 else 0.75,
                         weight=float(f32(rng.choice([1.0, 1.0, 2.0, 0.5, 1.5])))))
    for r in rows:
        r["bm25_bits"] = bits(pyref.bm25(r["tf"], r["df"], r["doc_len"] if r["doc_len"] > 0 else max(r["avgdl"], r["tf"]),
                                         r["avgdl"], r["docs"], r["k1"], r["b"]))
        r["score_tf_bits"] = bits(pyref.score_tf(r["tf"], r["df"], r["doc_len"], r["avgdl"], r["docs"], r["k1"], r["b"], r["weight"]))
    dump("bm25_scalar.json", rows)


def gen_wand_literal():
    # query/wand.rs:930-950 term_from_entries: doc lengths 10.0, avgdl 10, docs 10, k1 1.2, b 0.75, weight 1
    t1 = ([1, 3], [2, 1])
    t2 = ([3], [3])
    lens = [10.0] * 4
    top = pyref.exhaustive_top_k([t1, t2], [1.0, 1.0], lens, 10.0, 10.0, 1.2, 0.75, 2)
    dump("wand_literal.json", {
        "terms": [{"docs": t1[0], "tfs": t1[1]}, {"docs": t2[0], "tfs": t2[1]}],
        "doc_lengths": lens, "avgdl": 10.0, "docs": 10.0, "k1": 1.2, "b": 0.75, "k": 2,
        "expected": [{"doc_id": d, "score_bits": bits(s)} for d, s in top],
        "tie_break": {"a": {"doc_id": 1, "score": 1.0}, "b": {"doc_id": 2, "score": 1.0}, "worst_doc_id": 2},
    })


def gen_codec():
    vals = [0, 1, 127, 128, 16384, 0xFFFFFFFF, 300, 2_097_151, 2_097_152, 9_999_999]
    literal = {"docs": [1, 2], "tfs": [2, 1], "positions": [[1, 3], [4]]}
    rng = np.random.default_rng(5)
    docs = np.sort(rng.choice(50_000, size=300, replace=False)).tolist()
    tfs = rng.integers(1, 9, size=300).tolist()
    dump("codec.json", {
        "varint": [{"value": v, "hex": pyref.write_u32_var(v).hex()} for v in vals],
        "postings_literal": {**literal, "hex_with_positions": pyref.encode_postings(literal["docs"], literal["tfs"], literal["positions"]).hex(),
                             "hex_without_positions": pyref.encode_postings(literal["docs"], literal["tfs"], None).hex()},
        "postings_300": {"docs": docs, "tfs": tfs, "hex": pyref.encode_postings(docs, tfs, None).hex()},
    })


def gen_small_topk():
    rng = np.random.default_rng(77)
    n_docs, vocab = 120, 30
    k1, b = 0.9, 0.4
    lens = rng.integers(3, 40, size=n_docs)
    lens[5] = 0  # missing `_len:` value -> doc_len falls back to max(avgdl, 1) (wand.rs:77-84)
    post = {t: ([], []) for t in range(vocab)}
    for d in range(n_docs):
        n_tok = int(lens[d]) if lens[d] > 0 else 7
        toks = np.minimum((rng.zipf(1.3, size=n_tok) - 1), vocab - 1)
        u, c = np.unique(toks, return_counts=True)
        for t, cnt in zip(u.tolist(), c.tolist()):
            post[t][0].append(d)
            post[t][1].append(cnt)
    total = int(lens.sum())
    avgdl = f32(f32(total) / f32(n_docs))  # index/segment.rs:946-957
    queries = []
    for qi in range(12):
        nt = int(rng.integers(1, 5))
        terms = rng.choice(vocab, size=nt, replace=False).tolist()
        weights = [1.0] * nt if qi % 3 else [float(f32(w)) for w in rng.choice([1.0, 2.0, 0.5], size=nt)]
        k = [1, 3, 11][qi % 3]
        top = pyref.exhaustive_top_k([post[t] for t in terms], weights, lens.astype(np.float32), avgdl, float(n_docs), k1, b, k)
        queries.append({"terms": terms, "weights": weights, "k": k,
                        "expected": [{"doc_id": d, "score_bits": bits(s)} for d, s in top]})
    dump("small_topk.json", {
        "n_docs": n_docs, "vocab": vocab, "k1": k1, "b": b, "field_lengths": lens.tolist(), "total_tokens": total,
        "postings": [{"docs": post[t][0], "tfs": post[t][1]} for t in range(vocab)], "queries": queries,
    })


def gen_vectors():
    rng = np.random.default_rng(9)
    rows = []
    for dim in (2, 8, 768):
        for metric in ("cosine", "l2"):
            a = rng.standard_normal(dim).astype(np.float32)
            v = rng.standard_normal(dim).astype(np.float32)
            rows.append({"metric": metric, "a": a.tolist(), "b": v.tolist(), "similarity_bits": bits(pyref.metric_similarity(metric, a, v))})
    # blend_scores vectors/mod.rs:122-129: alpha*bm25 + (1-alpha)*vec
    blends = []
    for bm, vs, alpha in [(2.5, 0.75, 0.5), (0.0, -1.0, 0.3), (7.25, 0.1, 0.9), (1.0, float(np.finfo(np.float32).min), 0.5)]:
        blends.append({"bm25": bm, "vec": vs, "alpha": alpha,
                       "bits": bits(f32(f32(alpha) * f32(bm) + f32(f32(1.0) - f32(alpha)) * f32(vs)))})
    dump("vectors.json", {"similarity": rows, "blend": blends, "missing": {"cosine": -1.0, "l2_bits": bits(np.finfo(np.float32).min)}})


if __name__ == "__main__":
    gen_bm25_scalar()
    gen_wand_literal()
    gen_codec()
    gen_small_topk()
    gen_vectors()
    print("fixtures written to", HERE)
