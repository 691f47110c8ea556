"""ctypes wrapper of the CPU oracle (oracle/libslo_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by searchlite_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libslo_oracle.so")

EXEC = {"bm25": 0, "wand": 1, "bmw": 2, "bm25_dense": 3}

HIT_DTYPE = np.dtype({"names": ["segment_ord", "doc_id", "score"], "formats": ["<u4", "<u4", "<f4"], "itemsize": 12})
STATS_DTYPE = np.dtype({"names": ["scored_docs", "candidates_examined", "postings_advanced", "total_matches", "saw_cursor"],
                        "formats": ["<u8"] * 5, "itemsize": 40})

FILTER_DTYPE = np.dtype(
    {"names": ["op", "column", "i_min", "i_max", "f_min", "f_max", "n_children", "value_begin", "value_end"],
     "formats": ["<u4", "<i4", "<i8", "<i8", "<f8", "<f8", "<u4", "<u4", "<u4"],
     "offsets": [0, 4, 8, 16, 24, 32, 40, 44, 48], "itemsize": 56})  # slo_filter_node_t

_LIB = None


def build(force: bool = False) -> str:
    src = [os.path.join(HERE, f) for f in ("oracle.cc", "searchlite_oracle.h", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.run(["make", "-C", HERE, "-B", "libslo_oracle.so"], check=True, capture_output=True)
    return LIB_PATH


def lib() -> C.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    f32, vp, u32, i32, u64, sz = C.c_float, C.c_void_p, C.c_uint32, C.c_int32, C.c_uint64, C.c_size_t
    for name in ("slo_bm25",):
        getattr(L, name).argtypes = [f32] * 7
        getattr(L, name).restype = f32
    for name in ("slo_score_tf", "slo_upper_bound_tf"):
        getattr(L, name).argtypes = [f32] * 8
        getattr(L, name).restype = f32
    L.slo_idf.argtypes = [f32, f32]
    L.slo_idf.restype = f32
    L.slo_varint_write_u32.argtypes = [u32, vp]
    L.slo_varint_write_u32.restype = sz
    L.slo_varint_read_u32.argtypes = [vp, sz, C.POINTER(u32)]
    L.slo_varint_read_u32.restype = sz
    L.slo_postings_encode.argtypes = [vp, vp, sz, C.c_int, vp, vp, vp, sz]
    L.slo_postings_encode.restype = sz
    L.slo_postings_peek_df.argtypes = [vp, sz, C.POINTER(u32), C.POINTER(u32)]
    L.slo_postings_decode.argtypes = [vp, sz, C.c_int, vp, vp, C.POINTER(f32), C.POINTER(u32), vp, vp, C.POINTER(u32), C.POINTER(sz)]
    L.slo_plan_evaluate.argtypes = [vp, u32, vp, u32]
    L.slo_plan_evaluate.restype = f32
    L.slo_index_new.argtypes = [u32, u32, f32, f32]
    L.slo_index_new.restype = vp
    L.slo_index_free.argtypes = [vp]
    L.slo_index_set_postings.argtypes = [vp, u64, vp, vp, vp]
    L.slo_index_set_field_lengths.argtypes = [vp, vp, vp, u64]
    L.slo_index_set_deleted.argtypes = [vp, vp, u32]
    L.slo_index_build_post_image.argtypes = [vp]
    L.slo_index_set_positions.argtypes = [vp, vp, vp]
    L.slo_index_post_image_size.argtypes = [vp]
    L.slo_index_post_image_size.restype = u64
    L.slo_index_post_image.argtypes = [vp]
    L.slo_index_post_image.restype = vp
    L.slo_index_post_offsets.argtypes = [vp]
    L.slo_index_post_offsets.restype = vp
    for name in ("slo_index_avgdl", "slo_index_live_docs", "slo_index_min_doc_len"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = f32
    L.slo_index_add_i64_column.argtypes = [vp, vp, vp]
    L.slo_index_add_f64_column.argtypes = [vp, vp, vp]
    L.slo_index_add_str_column.argtypes = [vp, C.POINTER(C.c_char_p), u32, vp]
    L.slo_index_add_i64_list_column.argtypes = [vp, vp, vp]
    L.slo_index_add_f64_list_column.argtypes = [vp, vp, vp]
    L.slo_index_add_str_list_column.argtypes = [vp, C.POINTER(C.c_char_p), u32, vp, vp]
    L.slo_search.argtypes = [vp, vp, u32, C.c_int, u32, vp, u32, C.POINTER(C.c_char_p), C.c_int, vp, vp]
    L.slo_search.restype = i32
    L.slo_search_batch.argtypes = [vp, vp, u32, u32, C.c_int, u32, vp, u32, C.POINTER(C.c_char_p), C.c_int, C.c_int, vp, vp, vp]
    L.slo_search_batch.restype = i32
    L.slo_max_threads.restype = C.c_int
    L.slo_merge_hits.argtypes = [vp, u32, u32, vp]
    L.slo_merge_hits.restype = u32
    L.slo_filter_bitmap.argtypes = [vp, vp, u32, C.POINTER(C.c_char_p), vp]
    L.slo_matches_phrase_positions.argtypes = [u32, vp, vp, u32]
    L.slo_matches_phrase_positions.restype = C.c_int
    L.slo_phrase_bitmap.argtypes = [u32, u64, vp, vp, vp, vp, vp, u32, u32, vp]
    L.slo_normalize_in_place.argtypes = [vp, sz]
    L.slo_normalize_in_place.restype = None
    L.slo_metric_similarity.argtypes = [C.c_int, vp, vp, sz]
    L.slo_metric_similarity.restype = f32
    L.slo_blend_scores.argtypes = [f32, f32, f32, C.c_int]
    L.slo_blend_scores.restype = f32
    L.slo_hybrid_score.argtypes = [f32, C.c_int, f32, f32, C.c_int]
    L.slo_hybrid_score.restype = f32
    L.slo_hybrid_score_clauses.argtypes = [f32, u32, vp, vp, vp, vp, vp, vp]
    L.slo_hybrid_score_clauses.restype = f32
    L.slo_round_bf16.argtypes = [vp, sz]
    L.slo_round_bf16.restype = None
    L.slo_rerank_batch.argtypes = [u32, u32, vp, vp, u32, vp, vp, vp, vp, vp, u32, u32, vp, vp, vp, vp, vp, vp, vp, C.c_int]
    L.slo_rerank_batch.restype = C.c_int
    _LIB = L
    return L


def _p(a):
    return 0 if a is None else a.ctypes.data


class OracleIndex:
    """One segment held by the oracle.  Arrays are borrowed: this object keeps them alive."""

    def __init__(self, seg, k1: float = 0.9, b: float = 0.4):
        """seg: searchlite_b200.engine.SegmentData with HOST (numpy) arrays"""
        L = lib()
        self.L = L
        self.seg = seg
        self.doc_count = seg.doc_count
        self.h = L.slo_index_new(seg.segment_ord, seg.doc_count, k1, b)
        self.off = np.ascontiguousarray(seg.term_offsets, dtype=np.uint64)
        self.docs = np.ascontiguousarray(seg.post_docs, dtype=np.uint32)
        self.tfs = np.ascontiguousarray(seg.post_tfs, dtype=np.uint32)
        self.lens = np.ascontiguousarray(seg.field_lengths, dtype=np.int64)
        self.present = None if seg.field_length_present is None else np.ascontiguousarray(seg.field_length_present, dtype=np.uint8)
        L.slo_index_set_postings(self.h, len(self.off) - 1, _p(self.off), _p(self.docs), _p(self.tfs))
        L.slo_index_set_field_lengths(self.h, _p(self.lens), _p(self.present), int(seg.total_tokens))
        if seg.deleted_docs is not None and len(seg.deleted_docs):
            d = np.ascontiguousarray(seg.deleted_docs, dtype=np.uint32)
            L.slo_index_set_deleted(self.h, _p(d), len(d))
        self.columns = {}
        for name, (vals, present) in seg.fast_i64.items():
            v = np.ascontiguousarray(vals, dtype=np.int64)
            pr = None if present is None else np.ascontiguousarray(present, dtype=np.uint8)
            self.columns[name] = L.slo_index_add_i64_column(self.h, _p(v), _p(pr))
        for name, (vals, present) in seg.fast_f64.items():
            v = np.ascontiguousarray(vals, dtype=np.float64)
            pr = None if present is None else np.ascontiguousarray(present, dtype=np.uint8)
            self.columns[name] = L.slo_index_add_f64_column(self.h, _p(v), _p(pr))
        for name, (dic, ords) in seg.fast_str.items():
            o = np.ascontiguousarray(ords, dtype=np.uint32)
            arr = (C.c_char_p * len(dic))(*[s.encode() for s in dic])
            self.columns[name] = L.slo_index_add_str_column(self.h, arr, len(dic), _p(o))
        for name, (offs, vals) in getattr(seg, "fast_i64_list", {}).items():
            o, v = np.ascontiguousarray(offs, dtype=np.uint32), np.ascontiguousarray(vals, dtype=np.int64)
            self.columns[name] = L.slo_index_add_i64_list_column(self.h, _p(o), _p(v))
        for name, (offs, vals) in getattr(seg, "fast_f64_list", {}).items():
            o, v = np.ascontiguousarray(offs, dtype=np.uint32), np.ascontiguousarray(vals, dtype=np.float64)
            self.columns[name] = L.slo_index_add_f64_list_column(self.h, _p(o), _p(v))
        for name, (dic, offs, ords) in getattr(seg, "fast_str_list", {}).items():
            o, v = np.ascontiguousarray(offs, dtype=np.uint32), np.ascontiguousarray(ords, dtype=np.uint32)
            arr = (C.c_char_p * len(dic))(*[s.encode() for s in dic])
            self.columns[name] = L.slo_index_add_str_list_column(self.h, arr, len(dic), _p(o), _p(v))
        self._has_image = False

    def __del__(self):
        try:
            if self.h:
                self.L.slo_index_free(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def avgdl(self):
        return self.L.slo_index_avgdl(self.h)

    @property
    def live_docs(self):
        return self.L.slo_index_live_docs(self.h)

    @property
    def min_doc_len(self):
        return self.L.slo_index_min_doc_len(self.h)

    def set_positions(self, pos_offsets, positions):
        """positions on: build_post_image then writes `varint npos | npos x varint delta` per posting"""
        self.pos_offsets = np.ascontiguousarray(pos_offsets, dtype=np.uint64)
        self.positions = np.ascontiguousarray(positions, dtype=np.uint32)
        self.L.slo_index_set_positions(self.h, _p(self.pos_offsets), _p(self.positions))
        self._has_image = False

    def build_post_image(self):
        """the reference's `.post` byte image of every term + per-term offsets (n_terms+1)"""
        if not self._has_image:
            self.L.slo_index_build_post_image(self.h)
            self._has_image = True
        n = self.L.slo_index_post_image_size(self.h)
        img = np.ctypeslib.as_array(C.cast(self.L.slo_index_post_image(self.h), C.POINTER(C.c_uint8)), shape=(n,)).copy()
        off = np.ctypeslib.as_array(C.cast(self.L.slo_index_post_offsets(self.h), C.POINTER(C.c_uint64)),
                                    shape=(len(self.off),)).copy()
        return img, off

    def search_batch(self, batch, k: int, execution: str = "bm25", block_size: int = 0, filter_nodes=None,
                     strings=(), faithful: bool = False, threads: int = 1, want_stats: bool = False):
        """batch: searchlite_b200.engine.QueryBatch.  Returns (hits[Q,k], counts[Q][, stats])."""
        if faithful and not self._has_image:
            self.build_post_image()
        s = batch.structs()
        q = batch.n_queries
        hits = np.zeros((q, k), dtype=HIT_DTYPE)
        counts = np.zeros(q, dtype=np.uint32)
        stats = np.zeros(q, dtype=STATS_DTYPE) if want_stats else None
        arr = (C.c_char_p * max(len(strings), 1))(*[x.encode() for x in strings])
        nf = 0 if filter_nodes is None else len(filter_nodes)
        fn = None if filter_nodes is None else np.ascontiguousarray(filter_nodes, dtype=FILTER_DTYPE)
        rc = self.L.slo_search_batch(self.h, _p(s), q, k, EXEC[execution], block_size, _p(fn), nf, arr,
                                     1 if faithful else 0, threads, _p(hits), _p(counts), _p(stats))
        if rc < 0:
            raise RuntimeError(f"oracle search failed: {rc}")
        return (hits, counts, stats) if want_stats else (hits, counts)

    def filter_bitmap(self, filter_nodes, strings=()):
        out = np.zeros((self.doc_count + 31) // 32, dtype=np.uint32)
        arr = (C.c_char_p * max(len(strings), 1))(*[x.encode() for x in strings])
        fn = np.ascontiguousarray(filter_nodes, dtype=FILTER_DTYPE)
        self.L.slo_filter_bitmap(self.h, _p(fn), len(fn), arr, _p(out))
        return out


def merge_hits(hit_lists, limit: int):
    allh = np.concatenate([np.ascontiguousarray(h, dtype=HIT_DTYPE).reshape(-1) for h in hit_lists])
    out = np.zeros(limit, dtype=HIT_DTYPE)
    n = lib().slo_merge_hits(_p(allh), len(allh), limit, _p(out))
    return out[:n]


def round_bf16(values: np.ndarray) -> np.ndarray:
    """a copy of `values` (f32) rounded to bf16 precision — the engine's storage option"""
    out = np.array(values, dtype=np.float32, order="C", copy=True)
    lib().slo_round_bf16(_p(out), out.size)
    return out


def rerank_batch(cands: np.ndarray, counts: np.ndarray, stores, clauses, threads: int = 1):
    """hybrid step over the BM25 candidates (slo_rerank_batch).  stores: [(segment_ord, offsets u32[doc_count], rows f32[n, dim])];
    clauses: [(query_vecs f32[Q, dim], alpha, boost, metric name)] -> (hits, counts, vector scores)"""
    L = lib()
    cands = np.ascontiguousarray(cands, dtype=HIT_DTYPE)
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    nq, stride = cands.shape
    ords = np.array([s[0] for s in stores], dtype=np.uint32)
    offs = [np.ascontiguousarray(s[1], dtype=np.uint32) for s in stores]
    vals = [np.ascontiguousarray(s[2], dtype=np.float32) for s in stores]
    dcs = np.array([len(o) for o in offs], dtype=np.uint32)
    rows = np.array([v.shape[0] for v in vals], dtype=np.uint64)
    dim = int(clauses[0][0].shape[1])
    off_ptrs = (C.c_void_p * max(len(offs), 1))(*[o.ctypes.data for o in offs])
    val_ptrs = (C.c_void_p * max(len(vals), 1))(*[v.ctypes.data for v in vals])
    qvs = [np.ascontiguousarray(c[0], dtype=np.float32) for c in clauses]
    qv_ptrs = (C.c_void_p * len(qvs))(*[q.ctypes.data for q in qvs])
    alpha = np.array([c[1] for c in clauses], dtype=np.float32)
    boost = np.array([c[2] for c in clauses], dtype=np.float32)
    metric = np.array([0 if c[3] == "cosine" else 1 for c in clauses], dtype=np.int32)
    out = np.zeros((nq, stride), dtype=HIT_DTYPE)
    oc = np.zeros(nq, dtype=np.uint32)
    vs = np.zeros((nq, stride), dtype=np.float32)
    rc = L.slo_rerank_batch(nq, stride, _p(cands), _p(counts), len(stores), _p(ords), _p(dcs), off_ptrs, val_ptrs, _p(rows), dim,
                            len(clauses), qv_ptrs, _p(alpha), _p(boost), _p(metric), _p(out), _p(oc), _p(vs), threads)
    assert rc == 0
    return out, oc, vs


def max_threads() -> int:
    return lib().slo_max_threads()


def matches_phrase_positions(position_lists, slop: int) -> bool:
    """query/phrase.rs:4-48 on one doc's position lists (phrase order)."""
    arrs = [np.ascontiguousarray(p, dtype=np.uint32) for p in position_lists]
    ptrs = (C.c_void_p * max(len(arrs), 1))(*[a.ctypes.data for a in arrs])
    counts = np.array([len(a) for a in arrs], dtype=np.uint32)
    return bool(lib().slo_matches_phrase_positions(len(arrs), ptrs, _p(counts), slop))


def phrase_bitmap(doc_count: int, term_offsets, docs, pos_offsets, positions, phrase_terms, slop: int) -> np.ndarray:
    """Bitmap (uint32 words, LSB first) of the docs matches_phrase accepts."""
    term_offsets = np.ascontiguousarray(term_offsets, dtype=np.uint64)
    docs = np.ascontiguousarray(docs, dtype=np.uint32)
    pos_offsets = np.ascontiguousarray(pos_offsets, dtype=np.uint64)
    positions = np.ascontiguousarray(positions, dtype=np.uint32)
    pt = np.ascontiguousarray(phrase_terms, dtype=np.uint32)
    out = np.zeros((doc_count + 31) // 32, dtype=np.uint32)
    lib().slo_phrase_bitmap(doc_count, len(term_offsets) - 1, _p(term_offsets), _p(docs), _p(pos_offsets), _p(positions), _p(pt),
                            len(pt), slop, _p(out))
    return out
