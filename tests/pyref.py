"""Independent pure-Python / numpy-float32 restatement of the reference's scalar arithmetic and
codecs — TEST INFRASTRUCTURE.  It exists so that the C++ oracle (oracle/oracle.cc) is checked
against a second implementation written from the reference text, not against itself; the golden
fixtures under tests/golden/ are produced by tests/golden/make_golden.py from this file.

Every float operation is done on numpy.float32 scalars one at a time (no fused ops, no float64
intermediates) to mirror Rust's f32 semantics.  All citations: /root/reference/searchlite-core/src/.
"""
from __future__ import annotations

import math
import struct

import numpy as np

f32 = np.float32


def ln_f32(x: np.float32) -> np.float32:
    """f32::ln — the correctly rounded natural log (float64 log rounded once to f32).  glibc's
    logf is correctly rounded on every input the fixtures use (checked by make_golden.py)."""
    x = float(x)
    if x < 0.0 or x != x:
        return f32(np.nan)
    if x == 0.0:
        return f32(-np.inf)
    return f32(math.log(x))


def fmax(a, b):
    """f32::max: the non-NaN operand wins"""
    if a != a:
        return b
    if b != b:
        return a
    return a if a > b else b


def bm25(tf, df, doc_len, avgdl, docs, k1, b) -> np.float32:
    """query/bm25.rs:1-6"""
    tf, df, doc_len, avgdl, docs, k1, b = map(f32, (tf, df, doc_len, avgdl, docs, k1, b))
    idf = fmax(ln_f32((docs - df + f32(0.5)) / (df + f32(0.5))), f32(0.0)) + f32(1.0)
    norm_dl = doc_len / avgdl if avgdl > f32(0.0) else f32(1.0)
    denom = tf + k1 * (f32(1.0) - b + b * norm_dl)
    return f32(idf * (tf * (k1 + f32(1.0))) / fmax(denom, f32(1e-6)))


def score_tf(tf, df, doc_len, avgdl, docs, k1, b, weight) -> np.float32:
    """query/wand.rs:269-286"""
    tf, doc_len, avgdl = f32(tf), f32(doc_len), f32(avgdl)
    norm_len = doc_len if doc_len > f32(0.0) else max(avgdl, tf)
    return f32(bm25(tf, df, norm_len, avgdl, docs, k1, b) * f32(weight))


def doc_len_of(lens, doc, avgdl) -> np.float32:
    """ScoredTerm::doc_len query/wand.rs:77-84"""
    v = f32(lens[doc]) if doc < len(lens) else f32(0.0)
    return v if v > f32(0.0) else max(f32(avgdl), f32(1.0))


def total_cmp_key(x: np.float32) -> int:
    """f32::total_cmp as an integer key"""
    (i,) = struct.unpack("<i", struct.pack("<f", float(x)))
    return i ^ ((i >> 31) & 0x7FFFFFFF)


def exhaustive_top_k(term_postings, weights, lens, avgdl, docs, k1, b, k):
    """brute_force with a Sum-of-leaves plan, query/wand.rs:469-521 + planner Sum: every term is its
    own leaf, leaves are added in order starting from 0.0; then push_top_k/finalize_heap ordering
    (score desc by total_cmp, doc id asc), query/wand.rs:905-926.
    term_postings: list of (docs[], tfs[]) per term in leaf order."""
    leaves = {}
    n = len(term_postings)
    for li, (pd, pt) in enumerate(term_postings):
        df = f32(len(pd))
        for d, tf in zip(pd, pt):
            s = score_tf(tf, df, doc_len_of(lens, int(d), avgdl), avgdl, docs, k1, b, weights[li])
            buf = leaves.setdefault(int(d), [f32(0.0)] * n)
            buf[li] = f32(buf[li] + s)
    scored = []
    for d, buf in leaves.items():
        tot = f32(0.0)
        for v in buf:
            tot = f32(tot + v)
        scored.append((d, tot))
    scored.sort(key=lambda x: (-total_cmp_key(x[1]), x[0]))
    return scored[:k]


def write_u32_var(v: int) -> bytes:
    """util/varint.rs:5-15"""
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def read_u32_var(buf: bytes, pos: int = 0):
    """util/varint.rs:31-49 -> (value, next position); raises past 5 bytes"""
    shift, value = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        value |= (b & 0x7F) << shift
        if b & 0x80 == 0:
            return value & 0xFFFFFFFF, pos
        shift += 7
        if shift > 28:
            raise ValueError("varint too long")


def encode_postings(docs, tfs, positions=None, block_size: int = 128) -> bytes:
    """PostingsWriter::write_term index/postings.rs:78-129"""
    n = len(docs)
    keep = positions is not None
    out = bytearray()
    out += struct.pack("<I", n)
    out.append(1 if keep else 0)
    bc = (n + block_size - 1) // block_size
    out += struct.pack("<I", (bc | 0x80000000) if bc > 0 else 0)
    out += struct.pack("<I", int(docs[-1]) if n else 0)
    out += struct.pack("<f", float(max([float(t) for t in tfs], default=0.0)))
    if bc > 0:
        out += struct.pack("<I", block_size)
        for c in range(bc):
            out += struct.pack("<I", int(docs[min(n, (c + 1) * block_size) - 1]))
        for c in range(bc):
            out += struct.pack("<f", float(max(float(t) for t in tfs[c * block_size:(c + 1) * block_size])))
    for i in range(n):
        out += write_u32_var(int(docs[i]))
        out += write_u32_var(int(tfs[i]))
        if keep:
            out += write_u32_var(len(positions[i]))
            prev = 0
            for p in positions[i]:
                out += write_u32_var(p - prev)
                prev = p
    return bytes(out)


def metric_similarity(metric: str, a, b) -> np.float32:
    """vectors/mod.rs:98-120: cosine = dot of (pre-normalised) vectors, NaN -> 0; l2 = -sqrt(sum sq)"""
    acc = f32(0.0)
    if metric == "cosine":
        for x, y in zip(a, b):
            acc = f32(acc + f32(x) * f32(y))
        return f32(0.0) if np.isnan(acc) else acc
    for x, y in zip(a, b):
        d = f32(f32(x) - f32(y))
        acc = f32(acc + d * d)
    return f32(-np.sqrt(acc))
