"""Shared builders for the CPU and GPU tests."""
from __future__ import annotations

import json
import os
import struct

import numpy as np

from searchlite_b200.engine import QueryBatch, SegmentData

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name: str):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def f32_bits(x) -> int:
    return struct.unpack("<I", struct.pack("<f", float(x)))[0]


def segment_from_postings(postings, field_lengths, total_tokens=None, segment_ord: int = 0, deleted=None) -> SegmentData:
    """postings: list over term ids of (docs, tfs) — docs ascending; host CSR SegmentData"""
    off = np.zeros(len(postings) + 1, dtype=np.uint64)
    docs, tfs = [], []
    for t, (d, f) in enumerate(postings):
        off[t + 1] = off[t] + len(d)
        docs.extend(d)
        tfs.extend(f)
    lens = np.asarray(field_lengths, dtype=np.int64)
    if total_tokens is None:
        total_tokens = int(lens.sum())
    seg = SegmentData(segment_ord, len(lens), off, np.asarray(docs, dtype=np.uint32), np.asarray(tfs, dtype=np.uint32),
                      lens, int(total_tokens))
    if deleted is not None:
        seg.deleted_docs = np.asarray(deleted, dtype=np.uint32)
    return seg


def token_corpus(doc_tokens, vocab: int, segment_ord: int = 0) -> SegmentData:
    """doc_tokens: list over docs of token-id lists -> SegmentData (tf = multiplicity, `_len:` = token count)"""
    post = [([], []) for _ in range(vocab)]
    for d, toks in enumerate(doc_tokens):
        u, c = np.unique(np.asarray(toks, dtype=np.int64), return_counts=True)
        for t, n in zip(u.tolist(), c.tolist()):
            post[t][0].append(d)
            post[t][1].append(n)
    return segment_from_postings(post, [len(t) for t in doc_tokens], segment_ord=segment_ord)


def gpu_index(seg: SegmentData, kernel: str = "auto", k1: float = 0.9, b: float = 0.4, **kw):
    """a GpuIndex on cuda:0 holding `seg`"""
    from searchlite_b200 import GpuIndex
    gi = GpuIndex(0, kernel=kernel, **kw)
    gi.load_segment(seg, k1=k1, b=b)
    return gi


def or_queries(term_lists, weights=None) -> QueryBatch:
    return QueryBatch.from_term_lists(term_lists, weights)


def hits_to_list(hits, counts, q: int = 0):
    return [(int(h["doc_id"]), f32_bits(h["score"])) for h in hits[q][: int(counts[q])]]


def canonical_batch(gi, qb: QueryBatch, segment_ord: int = 0) -> QueryBatch:
    """The column path's declared summation order (include/searchlite_gpu.h, slg_set_option): per query
    the terms WITHOUT a dense column first, then the terms WITH one, both in query order.  The oracle's
    `bm25` mode on this permuted batch is what the items kernel must reproduce bit for bit."""
    terms = qb.terms.copy()
    cache = {}
    for q in range(qb.n_queries):
        a, b = int(qb.term_off[q]), int(qb.term_off[q + 1])
        rows = terms[a:b].copy()
        has = []
        for t in rows["term_id"].tolist():
            if t not in cache:
                cache[t] = False if t == 0xFFFFFFFF else gi.term_has_column(segment_ord, t)
            has.append(cache[t])
        order = [i for i, h in enumerate(has) if not h] + [i for i, h in enumerate(has) if h]
        terms[a:b] = rows[order]
    out = QueryBatch(qb.term_off.copy(), terms)
    out.filter_id = None if qb.filter_id is None else qb.filter_id.copy()
    if qb.group_off is not None:  # Bool queries: the group of a term travels in its row, the roles stay where they are
        out.group_off, out.group_role, out.min_should = qb.group_off.copy(), qb.group_role.copy(), qb.min_should.copy()
    return out


def assert_engine_parity(gi, ora, qb: QueryBatch, k: int, got, segment_ord: int = 0, exact_order: bool = False, **oracle_kw):
    """got = (hits, counts) of the engine for plain OR batch `qb`.  Bit-exact against the oracle on the
    engine's declared summation order, and within the 1e-5 rule against the reference (query) order."""
    from tests.parity import assert_parity
    ref = ora.search_batch(qb, k, "bm25", **oracle_kw)
    if exact_order:
        assert_parity(*ref, *got, strict=True)
        return
    assert_parity(*ref, *got, strict=False)
    canon = ora.search_batch(canonical_batch(gi, qb, segment_ord), k, "bm25", **oracle_kw)
    assert_parity(*canon, *got, strict=True)
