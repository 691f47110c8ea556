"""Test-side writer of the reference's on-disk index formats (no Rust toolchain here, so the files the
reference's IndexWriter would produce are restated from its writer code):

  seg_<id>.terms   searchlite-core/src/index/terms.rs:10-25
  seg_<id>.post    src/index/postings.rs:78-129 (positions on: every reference test/bench enables them)
  seg_<id>.fast    src/index/fastfields.rs:409-424, 910-1134
  seg_<id>.meta    src/index/segment.rs:43-53, 936-944 (serde_json pretty)
  seg_<id>.docs    docstore (not on the search path; written empty so that checksums cover a real file)
  seg_<id>_vectors/<field>.bin   src/index/segment.rs:1030-1053
  MANIFEST.json    src/index/manifest.rs:14-47

Documents are lists of token strings per text field; the writer mirrors SegmentWriter's bookkeeping
(src/index/segment.rs:660-700): term key "field:token", tf = occurrences, position = token index,
`_len:<field>` i64 fast field = token count, avg_field_lengths = total tokens / doc count as f32.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import json
import os
import struct
import zlib
from typing import Dict, List, Optional, Sequence

import numpy as np


def varint(v: int) -> bytes:  # util/varint.rs:5-11
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def encode_postings(entries, keep_positions: bool) -> bytes:
    """PostingsWriter::write_term, index/postings.rs:78-129; entries = [(doc, tf, [positions])] ascending"""
    n = len(entries)
    bs = 128
    bc = (n + bs - 1) // bs
    out = bytearray()
    out += struct.pack("<I", n)
    out += bytes([1 if keep_positions else 0])
    out += struct.pack("<I", (bc | 0x80000000) if bc else 0)
    out += struct.pack("<I", entries[-1][0] if n else 0)
    out += struct.pack("<f", float(max([e[1] for e in entries], default=0)))
    if bc:
        out += struct.pack("<I", bs)
        for i in range(0, n, bs):
            out += struct.pack("<I", entries[min(i + bs, n) - 1][0])
        for i in range(0, n, bs):
            out += struct.pack("<f", float(max(e[1] for e in entries[i:i + bs])))
    for doc, tf, pos in entries:
        out += varint(doc)
        out += varint(tf)
        if keep_positions:
            out += varint(len(pos))
            prev = 0
            for p in pos:
                out += varint(p - prev)
                prev = p
    return bytes(out)


def f32_json(x) -> float:
    """serde_json prints an f32 with its shortest round-trip digits (ryu)"""
    return float(np.format_float_positional(np.float32(x), unique=True, trim="0"))


class Segment:
    """One segment's worth of docs.  docs: list of {"body": [tokens...], ...}; keyword / numeric fast fields and
    vectors are given column-wise."""

    def __init__(self, seg_id: str, docs: Sequence[Dict[str, List[str]]], text_fields: Sequence[str],
                 keywords: Optional[Dict[str, List[Optional[str]]]] = None,
                 i64s: Optional[Dict[str, List[Optional[int]]]] = None,
                 f64s: Optional[Dict[str, List[Optional[float]]]] = None,
                 i64_lists: Optional[Dict[str, List[List[int]]]] = None,
                 f64_lists: Optional[Dict[str, List[List[float]]]] = None,
                 keyword_lists: Optional[Dict[str, List[List[str]]]] = None,
                 i64_nested: Optional[Dict[str, List[List[List[int]]]]] = None,
                 keyword_nested: Optional[Dict[str, List[List[List[str]]]]] = None,
                 vectors: Optional[Dict[str, tuple]] = None,  # field -> (metric "cosine"|"l2", [vec or None per doc])
                 deleted: Sequence[int] = (), keep_positions: bool = True):
        self.id = seg_id
        self.docs = docs
        self.text_fields = list(text_fields)
        self.keywords = keywords or {}
        self.i64s = i64s or {}
        self.f64s = f64s or {}
        self.i64_lists = i64_lists or {}
        self.f64_lists = f64_lists or {}
        self.keyword_lists = keyword_lists or {}
        self.i64_nested = i64_nested or {}
        self.keyword_nested = keyword_nested or {}
        self.vectors = vectors or {}
        self.deleted = list(deleted)
        self.keep_positions = keep_positions

    # ---- derived tables ----
    def postings(self):
        """{"field:token": [(doc, tf, [positions])]} — PostingsBuilder::add_term, index/segment.rs:679-684"""
        table: Dict[str, list] = {}
        for d, doc in enumerate(self.docs):
            for field in self.text_fields:
                per = {}
                for pos, tok in enumerate(doc.get(field, [])):
                    per.setdefault(tok, []).append(pos)
                for tok, ps in per.items():
                    table.setdefault(f"{field}:{tok}", []).append((d, len(ps), ps))
        return table

    def avg_field_lengths(self):  # compute_avg_lengths, index/segment.rs:946-957
        n = len(self.docs)
        out = {}
        for field in self.text_fields:
            total = sum(len(doc.get(field, [])) for doc in self.docs)
            out[field] = float(np.float32(total) / np.float32(n)) if n else 0.0
        return out

    # ---- files ----
    def post_and_terms(self):
        table = self.postings()
        post = bytearray()
        entries = []
        for key in sorted(table):  # BTreeMap order in the writer; the reader does not depend on it
            entries.append((key, len(post)))
            post += encode_postings(table[key], self.keep_positions)
        body = bytearray()
        for key, off in entries:
            kb = key.encode()
            body += varint(len(kb)) + kb + struct.pack("<Q", off)
        terms = struct.pack("<Q", len(entries)) + bytes(body) + struct.pack("<I", zlib.crc32(bytes(body)))
        return bytes(post), terms

    def fast(self) -> bytes:
        n = len(self.docs)
        fields = []

        def name_hdr(name, ty):
            nb = name.encode()
            return struct.pack("<I", len(nb)) + nb + bytes([ty]) + struct.pack("<I", n)

        for field in self.text_fields:  # doc_length_key, fastfields.rs:1162-1164
            vals = [len(doc.get(field, [])) for doc in self.docs]
            fields.append(name_hdr(f"_len:{field}", 0) + bytes([1] * n) + b"".join(struct.pack("<q", v) for v in vals))
        for name, vals in self.i64s.items():
            fields.append(name_hdr(name, 0) + bytes([0 if v is None else 1 for v in vals]) +
                          b"".join(struct.pack("<q", 0 if v is None else v) for v in vals))
        for name, vals in self.f64s.items():
            fields.append(name_hdr(name, 1) + bytes([0 if v is None else 1 for v in vals]) +
                          b"".join(struct.pack("<d", 0.0 if v is None else v) for v in vals))
        for name, vals in self.keywords.items():
            dic: List[str] = []
            ords = []
            for v in vals:
                if v is None:
                    ords.append(0xFFFFFFFF)
                else:
                    if v not in dic:
                        dic.append(v)
                    ords.append(dic.index(v))
            b = name_hdr(name, 2) + struct.pack("<I", len(dic))
            for s in dic:
                sb = s.encode()
                b += struct.pack("<I", len(sb)) + sb
            fields.append(b + b"".join(struct.pack("<I", o) for o in ords))
        def running(lists):
            offs = [0]
            for l in lists:
                offs.append(offs[-1] + len(l))
            return b"".join(struct.pack("<I", o) for o in offs)

        def dictionary(values):
            dic: List[str] = []
            for v in values:
                if v not in dic:
                    dic.append(v)
            b = struct.pack("<I", len(dic))
            for s in dic:
                sb = s.encode()
                b += struct.pack("<I", len(sb)) + sb
            return dic, b

        # write_field, index/fastfields.rs:926-1110: list columns are running offsets + values, nested columns doc -> object
        # offsets, object -> value offsets, values
        for name, lists in self.i64_lists.items():
            fields.append(name_hdr(name, 3) + running(lists) + b"".join(struct.pack("<q", v) for l in lists for v in l))
        for name, lists in self.f64_lists.items():
            fields.append(name_hdr(name, 4) + running(lists) + b"".join(struct.pack("<d", v) for l in lists for v in l))
        for name, lists in self.keyword_lists.items():
            dic, db = dictionary([v for l in lists for v in l])
            fields.append(name_hdr(name, 5) + db + running(lists) + b"".join(struct.pack("<I", dic.index(v)) for l in lists for v in l))
        for name, docs in self.i64_nested.items():
            objs = [o for d in docs for o in d]
            fields.append(name_hdr(name, 6) + running(docs) + running(objs) + b"".join(struct.pack("<q", v) for o in objs for v in o))
        for name, docs in self.keyword_nested.items():
            objs = [o for d in docs for o in d]
            dic, db = dictionary([v for o in objs for v in o])
            fields.append(name_hdr(name, 8) + db + running(docs) + running(objs) +
                          b"".join(struct.pack("<I", dic.index(v)) for o in objs for v in o))
        # HashMap order in the reference: any order; rotate so that `_len:` is not first
        fields = fields[1:] + fields[:1]
        return b"FFV1" + struct.pack("<I", len(fields)) + b"".join(fields)

    def meta(self) -> bytes:
        n = len(self.docs)
        m = {
            "doc_offsets": [0] * n,
            "doc_ids": [f"{i:08d} \"q\" \\ {{" for i in range(n)],  # strings the JSON scanner must skip correctly
            "avg_field_lengths": {k: f32_json(v) for k, v in self.avg_field_lengths().items()},
            "vector_fields": {f: {"dim": len(next(v for v in vecs if v is not None)), "metric": "Cosine" if m_ == "cosine" else "L2"}
                              for f, (m_, vecs) in self.vectors.items()},
            "use_zstd": False,
        }
        return json.dumps(m, indent=2).encode()

    def vector_file(self, field: str) -> bytes:
        metric, vecs = self.vectors[field]
        dim = len(next(v for v in vecs if v is not None))
        offs, rows = [], []
        for v in vecs:
            if v is None:
                offs.append(0xFFFFFFFF)
            else:
                offs.append(len(rows))
                rows.append(np.asarray(v, dtype="<f4"))
        hdr = struct.pack("<IIIBBHII", 0x56435452, 1, dim, 0 if metric == "cosine" else 1, 0, 0, len(vecs), len(rows))
        return hdr + np.asarray(offs, dtype="<u4").tobytes() + (np.stack(rows).tobytes() if rows else b"")


def write_index(root: str, segments: Sequence[Segment], stored_root: Optional[str] = None) -> dict:
    """Writes the directory; returns the manifest dict.  stored_root: the path prefix recorded in the manifest
    (root.join(..) of the machine that wrote the index — may differ from where the files are now)."""
    os.makedirs(root, exist_ok=True)
    stored_root = stored_root or root
    metas = []
    for seg in segments:
        post, terms = seg.post_and_terms()
        files = {"terms": terms, "postings": post, "docstore": b"", "fast": seg.fast(), "meta": seg.meta()}
        names = {"terms": f"seg_{seg.id}.terms", "postings": f"seg_{seg.id}.post", "docstore": f"seg_{seg.id}.docs",
                 "fast": f"seg_{seg.id}.fast", "meta": f"seg_{seg.id}.meta"}
        for k, data in files.items():
            with open(os.path.join(root, names[k]), "wb") as f:
                f.write(data)
        paths = {k: os.path.join(stored_root, v) for k, v in names.items()}
        if seg.vectors:
            vdir = f"seg_{seg.id}_vectors"
            os.makedirs(os.path.join(root, vdir), exist_ok=True)
            for field in seg.vectors:
                with open(os.path.join(root, vdir, f"{field}.bin"), "wb") as f:
                    f.write(seg.vector_file(field))
            paths["vector_dir"] = os.path.join(stored_root, vdir)
        metas.append({
            "id": seg.id, "generation": 1, "paths": paths, "doc_count": len(seg.docs),
            "max_doc_id": max(len(seg.docs) - 1, 0), "blockmax": True, "deleted_docs": seg.deleted,
            "avg_field_lengths": {k: f32_json(v) for k, v in seg.avg_field_lengths().items()},
            "checksums": {k: zlib.crc32(v) for k, v in files.items()},
        })
    manifest = {"version": 1, "uuid": "00000000-0000-4000-8000-000000000000", "segments": metas,
                "committed_at": "2026-01-01T00:00:00+00:00", "schema": {"doc_id_field": "_id", "text_fields": []}}
    with open(os.path.join(root, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=2)
    return manifest


def csr_of(seg: Segment, field: str):
    """The same segment as CSR arrays for the CSR load path and the oracle: (keys, term_offsets, docs, tfs,
    pos_offsets, positions, lens).  keys are the field's terms in sorted key order; term id = index."""
    table = seg.postings()
    keys = sorted(k for k in table if k.startswith(field + ":"))
    toff = [0]
    docs, tfs, poff, pos = [], [], [0], []
    for k in keys:
        for d, tf, ps in table[k]:
            docs.append(d)
            tfs.append(tf)
            pos.extend(ps)
            poff.append(len(pos))
        toff.append(len(docs))
    lens = [len(doc.get(field, [])) for doc in seg.docs]
    return (keys, np.asarray(toff, dtype=np.uint64), np.asarray(docs, dtype=np.uint32), np.asarray(tfs, dtype=np.uint32),
            np.asarray(poff, dtype=np.uint64), np.asarray(pos, dtype=np.uint32), np.asarray(lens, dtype=np.int64))
