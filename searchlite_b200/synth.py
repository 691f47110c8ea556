"""Seeded synthetic corpora and query batches.

Everything is integer / float64 torch code that runs unchanged on CPU (small test corpora,
checked against the oracle) and on CUDA (the 10 M-doc bench corpus, born in HBM), and yields
bit-identical output on both: tokens come from a counter-based splitmix64 hash of
(seed, doc, position), turned into a Zipf rank by bisection of a float64 CDF computed once on
the CPU.

Corpus model (config C2): N docs, one text field, V integer terms, term rank r ~ Zipf(s)
(p(r) ∝ r^-s), doc length ~ uniform{len_lo..len_hi}, tf = multiplicity of the term in the doc.
Doc ordinal == doc id (the reference orders docs by `_id`, searchlite-core/src/api/writer.rs:126).
Term id = rank - 1.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from .engine import TERM_DTYPE, TERM_SCORED, QueryBatch, SegmentData

_M64 = (1 << 64) - 1


def _s64(v: int) -> int:
    v &= _M64
    return v - (1 << 64) if v >= (1 << 63) else v


_GOLD = _s64(0x9E3779B97F4A7C15)
_C1 = _s64(0xBF58476D1CE4E5B9)
_C2 = _s64(0x94D049BB133111EB)


def _lsr(x: torch.Tensor, s: int) -> torch.Tensor:
    """logical right shift of int64"""
    return (x >> s) & ((1 << (64 - s)) - 1)


def splitmix64(x: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser on int64 tensors (wrapping arithmetic)"""
    z = x + _GOLD
    z = (z ^ _lsr(z, 30)) * _C1
    z = (z ^ _lsr(z, 27)) * _C2
    return z ^ _lsr(z, 31)


def hash2(seed: int, a: torch.Tensor) -> torch.Tensor:
    s = torch.tensor(_s64(seed * 0x9E3779B97F4A7C15 + 0x632BE59BD9B4E019), dtype=torch.int64, device=a.device)
    return splitmix64(splitmix64(a ^ s) + s)


def uniform01(h: torch.Tensor) -> torch.Tensor:
    """53-bit uniform in [0,1) as float64 (exact)"""
    return _lsr(h, 11).to(torch.float64) * (2.0 ** -53)


def zipf_cdf(vocab: int, s: float = 1.0) -> torch.Tensor:
    """float64 CDF over ranks 1..vocab, computed on the CPU so every device sees the same bits"""
    w = 1.0 / torch.arange(1, vocab + 1, dtype=torch.float64) ** s
    c = torch.cumsum(w, 0)
    return c / c[-1]


@dataclass
class CorpusSpec:
    n_docs: int
    vocab: int
    seed: int = 20260101
    zipf_s: float = 1.0
    len_lo: int = 100
    len_hi: int = 300
    segment_ord: int = 0
    doc_base: int = 0  # global ordinal of this segment's first doc (shards of one logical corpus)


def doc_lengths(spec: CorpusSpec, device) -> torch.Tensor:
    d = torch.arange(spec.doc_base, spec.doc_base + spec.n_docs, dtype=torch.int64, device=device)
    h = hash2(spec.seed ^ 0x5151, d)
    return _lsr(h, 1) % (spec.len_hi - spec.len_lo + 1) + spec.len_lo


def generate_segment(spec: CorpusSpec, device="cpu", chunk_docs: int = 1 << 18) -> SegmentData:
    """Token-level simulation -> inverted index in CSR form (term-major, docs ascending)."""
    device = torch.device(device)
    assert spec.len_hi < 512 and spec.vocab < (1 << 23) and spec.n_docs < (1 << 32)
    cdf = zipf_cdf(spec.vocab, spec.zipf_s).to(device)
    lens = doc_lengths(spec, device)
    keys = []
    pos = torch.arange(spec.len_hi, dtype=torch.int64, device=device)
    for d0 in range(0, spec.n_docs, chunk_docs):
        d1 = min(spec.n_docs, d0 + chunk_docs)
        local = torch.arange(d0, d1, dtype=torch.int64, device=device)
        gdoc = local + spec.doc_base
        ctr = gdoc[:, None] * 512 + pos[None, :]
        u = uniform01(hash2(spec.seed, ctr))
        term = torch.searchsorted(cdf, u, right=True).clamp_(max=spec.vocab - 1)
        valid = pos[None, :] < lens[d0:d1, None]
        k = (term << 40) | (local[:, None] << 8)
        k = k[valid]
        del ctr, u, term, valid
        k, _ = torch.sort(k)
        uk, cnt = torch.unique_consecutive(k, return_counts=True)
        assert int(cnt.max()) < 256, "tf does not fit the packed key"
        keys.append(uk | cnt)
        del k, uk, cnt
    allk = torch.cat(keys) if len(keys) > 1 else keys[0]
    del keys
    allk, _ = torch.sort(allk)
    term = allk >> 40
    docs = ((allk >> 8) & 0xFFFFFFFF).to(torch.int32)
    tfs = (allk & 0xFF).to(torch.int32)
    counts = torch.bincount(term, minlength=spec.vocab)
    del allk, term
    offsets = torch.zeros(spec.vocab + 1, dtype=torch.int64, device=device)
    offsets[1:] = torch.cumsum(counts, 0)
    total_tokens = int(lens.sum().item())
    seg = SegmentData(spec.segment_ord, spec.n_docs, offsets, docs, tfs, lens, total_tokens)
    if device.type == "cpu":
        seg = SegmentData(spec.segment_ord, spec.n_docs, offsets.numpy().astype(np.uint64),
                          docs.numpy().view(np.uint32), tfs.numpy().view(np.uint32), lens.numpy(), total_tokens)
    return seg


def token_terms(spec: CorpusSpec, d0: int, d1: int, cdf: torch.Tensor, device) -> tuple:
    """term id of every (doc, position) of local docs [d0, d1): (term[n, len_hi], valid[n, len_hi]) — the pure function
    of (seed, doc, position) both generate_segment and generate_positions are built on"""
    pos = torch.arange(spec.len_hi, dtype=torch.int64, device=device)
    gdoc = torch.arange(d0, d1, dtype=torch.int64, device=device) + spec.doc_base
    u = uniform01(hash2(spec.seed, gdoc[:, None] * 512 + pos[None, :]))
    term = torch.searchsorted(cdf, u, right=True).clamp_(max=spec.vocab - 1)
    lens = doc_lengths(spec, device)[d0:d1]
    return term, pos[None, :] < lens[:, None]


def generate_positions(spec: CorpusSpec, device="cpu", chunk_docs: int = 1 << 18) -> torch.Tensor:
    """Token positions of the corpus in the posting order of generate_segment (term-major, docs ascending,
    positions ascending inside a posting): int32 [total_tokens].  The positions of posting p are
    positions[off[p] : off[p] + tf[p]] with off = exclusive cumsum of the segment's post_tfs
    (the writer records position = token index, searchlite-core/src/index/segment.rs:675-684)."""
    device = torch.device(device)
    assert spec.len_hi < 512 and spec.vocab <= (1 << 22) and spec.n_docs < (1 << 32)
    cdf = zipf_cdf(spec.vocab, spec.zipf_s).to(device)
    pos = torch.arange(spec.len_hi, dtype=torch.int64, device=device)
    keys = []
    for d0 in range(0, spec.n_docs, chunk_docs):
        d1 = min(spec.n_docs, d0 + chunk_docs)
        term, valid = token_terms(spec, d0, d1, cdf, device)
        local = torch.arange(d0, d1, dtype=torch.int64, device=device)
        k = ((term << 41) | (local[:, None] << 9) | pos[None, :])[valid]
        del term, valid
        keys.append(torch.sort(k)[0])
        del k
    allk = torch.cat(keys) if len(keys) > 1 else keys[0]
    del keys
    allk, _ = torch.sort(allk)
    return (allk & 511).to(torch.int32)


def generate_queries(n_queries: int, vocab: int, seed: int = 20260102, zipf_s: float = 1.0, min_terms: int = 2,
                     max_terms: int = 5, min_rank: int = 10) -> QueryBatch:
    """OR queries of min_terms..max_terms DISTINCT terms drawn from the corpus Zipf conditional on
    rank >= min_rank (drops the stop-word head)."""
    cdf = zipf_cdf(vocab, zipf_s)
    base = float(cdf[min_rank - 2]) if min_rank >= 2 else 0.0
    q = torch.arange(n_queries, dtype=torch.int64)
    nt = (_lsr(hash2(seed ^ 0x77, q), 1) % (max_terms - min_terms + 1) + min_terms).numpy()
    draws = 4 * max_terms
    ctr = q[:, None] * 64 + torch.arange(draws, dtype=torch.int64)[None, :]
    u = base + uniform01(hash2(seed, ctr)) * (1.0 - base)
    cand = torch.searchsorted(cdf, u, right=True).clamp_(max=vocab - 1).numpy()
    term_lists = []
    for i in range(n_queries):
        seen, out = set(), []
        for t in cand[i]:
            t = int(t)
            if t not in seen:
                seen.add(t)
                out.append(t)
                if len(out) == nt[i]:
                    break
        term_lists.append(out)
    return QueryBatch.from_term_lists(term_lists)


def fast_fields(spec: CorpusSpec, n_lang: int = 8) -> tuple:
    """C4 columns: `lang` keyword (n_lang values, Zipf) and `year` i64 uniform{2000..2025}."""
    d = torch.arange(spec.doc_base, spec.doc_base + spec.n_docs, dtype=torch.int64)
    year = (_lsr(hash2(spec.seed ^ 0xA1, d), 1) % 26 + 2000).numpy()
    lcdf = zipf_cdf(n_lang, 1.0)
    lang = torch.searchsorted(lcdf, uniform01(hash2(spec.seed ^ 0xB2, d)), right=True).clamp_(max=n_lang - 1)
    names = ["en", "es", "de", "fr", "pt", "it", "nl", "sv", "pl", "tr", "ja", "zh"][:n_lang]
    return names, lang.numpy().astype(np.uint32), year.astype(np.int64)
