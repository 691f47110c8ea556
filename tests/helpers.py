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


def or_queries(term_lists, weights=None) -> QueryBatch:
    return QueryBatch.from_term_lists(term_lists, weights)


def hits_to_list(hits, counts, q: int = 0):
    return [(int(h["doc_id"]), f32_bits(h["score"])) for h in hits[q][: int(counts[q])]]
