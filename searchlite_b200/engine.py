"""ctypes binding of libsearchlite_gpu.so — the host-side mirror of the reference's search API
for the hot path (searchlite-core/src/api/reader.rs:2539 ``IndexReader::search`` →
``search_segment`` :2908 → ``execute_top_k`` src/query/wand.rs:338).

Names follow the reference: segments, postings, ``execution`` ∈ {"bm25", "wand", "bmw"}
(src/api/types.rs:6-13), ``limit``/internal ``k = limit + 1`` (api/reader.rs:2595-2619).
The binding fails loudly when the CUDA library is missing or no device is present: there is
no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import build as _build

EXECUTION = {"bm25": 0, "wand": 1, "bmw": 2}
ROLE_SHOULD, ROLE_MUST, ROLE_MUST_NOT = 0, 1, 2
TERM_SCORED = 1
ABSENT_TERM = 0xFFFFFFFF
MEM_HOST, MEM_DEVICE = 0, 1

F_KEYWORD_EQ, F_KEYWORD_IN, F_I64_RANGE, F_F64_RANGE, F_AND, F_OR, F_NOT = range(7)
METRIC = {"cosine": 0, "l2": 1}

# numpy mirrors of the C structs (include/searchlite_gpu.h)
TERM_DTYPE = np.dtype(
    {"names": ["term_id", "weight", "leaf", "group", "flags"],
     "formats": ["<u4", "<f4", "<u4", "<u4", "<u4"], "itemsize": 20})
QUERY_DTYPE = np.dtype(
    {"names": ["n_terms", "terms", "n_groups", "group_role", "min_should", "leaf_count", "filter_id", "n_plan_nodes", "plan",
               "has_cursor", "cursor_segment_ord", "cursor_doc_id", "cursor_score"],
     "formats": ["<u4", "<u8", "<u4", "<u8", "<u4", "<u4", "<i4", "<u4", "<u8", "<u4", "<u4", "<u4", "<f4"],
     "offsets": [0, 8, 16, 24, 32, 36, 40, 44, 48, 56, 60, 64, 68], "itemsize": 72})
PLAN_DTYPE = np.dtype({"names": ["op", "arg", "tie_breaker"], "formats": ["<u4", "<u4", "<f4"], "itemsize": 12})
PLAN_LEAF, PLAN_SUM, PLAN_DISMAX = 0, 1, 2


def plan_postfix(expr) -> list:
    """ScoreExpr (query/planner.rs:113-122) -> postfix slg_plan_node_t rows.  expr is
    ("leaf", i) | ("sum", [children]) | ("dismax", [children], tie_breaker)."""
    if expr[0] == "leaf":
        return [(PLAN_LEAF, int(expr[1]), 0.0)]
    rows = []
    for c in expr[1]:
        rows += plan_postfix(c)
    if expr[0] == "sum":
        return rows + [(PLAN_SUM, len(expr[1]), 0.0)]
    if expr[0] == "dismax":
        return rows + [(PLAN_DISMAX, len(expr[1]), float(expr[2]))]
    raise ValueError(f"unknown score expression {expr[0]!r}")
HIT_DTYPE = np.dtype({"names": ["segment_ord", "doc_id", "score"], "formats": ["<u4", "<u4", "<f4"], "itemsize": 12})
STATS_DTYPE = np.dtype(
    {"names": ["scored_docs", "postings_advanced", "blocks_skipped", "candidates_examined", "total_matches"],
     "formats": ["<u8"] * 5, "itemsize": 40})
FILTER_DTYPE = np.dtype(
    {"names": ["op", "column", "i_min", "i_max", "f_min", "f_max", "n_children", "value_begin", "value_end"],
     "formats": ["<u4", "<i4", "<i8", "<i8", "<f8", "<f8", "<u4", "<u4", "<u4"],
     "offsets": [0, 4, 8, 16, 24, 32, 40, 44, 48], "itemsize": 56})


class SegmentView(C.Structure):
    _fields_ = [
        ("segment_ord", C.c_uint32), ("doc_count", C.c_uint32), ("n_terms", C.c_uint64),
        ("term_offsets", C.c_void_p), ("post_docs", C.c_void_p), ("post_tfs", C.c_void_p),
        ("field_lengths", C.c_void_p), ("field_length_present", C.c_void_p),
        ("total_tokens", C.c_uint64), ("deleted_docs", C.c_void_p), ("n_deleted", C.c_uint32),
        ("memory_space", C.c_int32),
    ]


class SegmentFiles(C.Structure):
    """slg_segment_files_t: one segment as the reference's writer left it on disk"""
    _fields_ = [
        ("segment_ord", C.c_uint32), ("doc_count", C.c_uint32),
        ("terms", C.c_void_p), ("terms_bytes", C.c_uint64), ("post", C.c_void_p), ("post_bytes", C.c_uint64),
        ("fast", C.c_void_p), ("fast_bytes", C.c_uint64), ("meta", C.c_void_p), ("meta_bytes", C.c_uint64),
        ("deleted_docs", C.c_void_p), ("n_deleted", C.c_uint32), ("checksums", C.c_void_p),
    ]


class SegmentInfo(C.Structure):
    _fields_ = [
        ("n_terms_total", C.c_uint64), ("n_terms_field", C.c_uint64), ("n_postings", C.c_uint64), ("avgdl", C.c_float),
        ("has_positions", C.c_uint32), ("has_length_column", C.c_uint32), ("n_fast_columns", C.c_uint32),
        ("n_scalar_columns", C.c_uint32), ("crc_terms", C.c_uint32), ("crc_postings", C.c_uint32), ("crc_fast", C.c_uint32),
        ("crc_meta", C.c_uint32), ("n_list_columns", C.c_uint32), ("reserved", C.c_uint32),
    ]


COMBINE = {"and": 0, "or": 1, "and_not": 2}


class Counters(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_uint64), ("score_launches", C.c_uint64), ("score_ms_total", C.c_double),
        ("last_score_ms", C.c_double), ("last_batch_ms", C.c_double), ("last_posting_count", C.c_uint64),
        ("resident_bytes", C.c_uint64), ("last_h2d_bytes", C.c_uint64), ("last_d2h_bytes", C.c_uint64),
        ("last_postings_scattered", C.c_uint64), ("last_subtiles_skipped", C.c_uint64),
        ("last_column_blocks_streamed", C.c_uint64), ("last_items", C.c_uint64),
        ("last_postings_verified", C.c_uint64), ("last_items_dropped", C.c_uint64),
        ("rerank_launches", C.c_uint64), ("rerank_ms_total", C.c_double), ("last_rerank_ms", C.c_double),
    ]


class VectorClause(C.Structure):
    """slg_vector_clause_t"""
    _fields_ = [("query_vecs", C.c_void_p), ("alpha", C.c_float), ("boost", C.c_float), ("metric", C.c_int32), ("reserved", C.c_uint32)]


class SearchliteGpuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"searchlite_gpu error {code}: {message}")
        self.code = code


_LIB = None

# every symbol include/searchlite_gpu.h declares
EXPORTED_SYMBOLS = [
    "slg_open", "slg_close", "slg_last_error", "slg_configure", "slg_load_segment", "slg_load_segment_post_image",
    "slg_add_i64_column", "slg_add_f64_column", "slg_add_str_column", "slg_add_i64_list_column", "slg_add_f64_list_column", "slg_add_str_list_column", "slg_segment_stats", "slg_filter_compile",
    "slg_filter_bitmap", "slg_search_batch", "slg_batch_prepare", "slg_batch_run", "slg_batch_fetch",
    "slg_batch_device_results", "slg_batch_free", "slg_merge_gathered", "slg_load_vectors", "slg_rerank",
    "slg_get_counters", "slg_version", "slg_batch_copy_results_device", "slg_get_stream", "slg_selftest_div", "slg_batch_enable_stats",
    "slg_set_option", "slg_term_has_column",
    "slg_inspect_segment_files", "slg_load_segment_files", "slg_load_index_dir", "slg_load_index_dir_shard", "slg_load_vector_file", "slg_term_lookup",
    "slg_column_lookup", "slg_field_stats", "slg_load_positions", "slg_phrase_compile", "slg_phrase_compile_batch", "slg_filter_combine", "slg_filter_combine_batch", "slg_filter_free",
    "slg_batch_cursor_seen", "slg_cursor_encode", "slg_cursor_decode",
    "slg_batch_run_seeds", "slg_batch_threshold_keys", "slg_batch_import_thresholds", "slg_batch_run_sweep",
    "slg_batch_packed_results", "slg_merge_gathered_packed",
    "slg_batch_set_threshold_board", "slg_segment_residency", "slg_load_vectors_bf16", "slg_rerank_clauses", "slg_rerank_batch", "slg_batch_fetch_vector_scores", "slg_merge_gathered_hybrid",
]


def load_library(build_if_missing: bool = True) -> C.CDLL:
    """dlopen the in-tree CUDA library; raises if it is absent (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB_PATH
    if not os.path.exists(path):
        if not build_if_missing:
            raise RuntimeError(f"{path} is missing: build it with `python -m searchlite_b200.build`")
        _build.build()
    lib = C.CDLL(path)
    lib.slg_last_error.restype = C.c_char_p
    lib.slg_last_error.argtypes = [C.c_void_p]
    lib.slg_version.restype = C.c_char_p
    vp, u32, i32, u64, f32 = C.c_void_p, C.c_uint32, C.c_int32, C.c_uint64, C.c_float
    sigs = {
        "slg_open": [i32, C.POINTER(vp)],
        "slg_close": [vp],
        "slg_configure": [vp, u32, u32, u32, u32],
        "slg_load_segment": [vp, C.POINTER(SegmentView), f32, f32],
        "slg_load_segment_post_image": [vp, C.POINTER(SegmentView), vp, u64, vp, f32, f32],
        "slg_add_i64_column": [vp, u32, vp, vp],
        "slg_add_f64_column": [vp, u32, vp, vp],
        "slg_add_str_column": [vp, u32, C.POINTER(C.c_char_p), u32, vp],
        "slg_add_i64_list_column": [vp, u32, vp, vp],
        "slg_add_f64_list_column": [vp, u32, vp, vp],
        "slg_add_str_list_column": [vp, u32, C.POINTER(C.c_char_p), u32, vp, vp],
        "slg_segment_stats": [vp, u32, C.POINTER(f32), C.POINTER(f32), C.POINTER(f32), C.POINTER(u64)],
        "slg_filter_compile": [vp, vp, u32, C.POINTER(C.c_char_p)],
        "slg_filter_bitmap": [vp, i32, u32, vp],
        "slg_search_batch": [vp, vp, u32, u32, i32, u32, vp, vp, vp],
        "slg_batch_prepare": [vp, vp, u32, u32, i32, u32, C.POINTER(vp)],
        "slg_batch_run": [vp, i32],
        "slg_batch_fetch": [vp, vp, vp, vp],
        "slg_batch_device_results": [vp, C.POINTER(vp), C.POINTER(vp)],
        "slg_batch_free": [vp],
        "slg_batch_cursor_seen": [vp, vp],
        "slg_cursor_encode": [u32, u32, vp, C.c_char_p],
        "slg_cursor_decode": [C.c_char_p, u32, vp, C.POINTER(u32), C.c_char_p, u64],
        "slg_merge_gathered": [vp, vp, vp, u32, u32, u32, vp, vp],
        "slg_load_vectors": [vp, u32, u32, vp, vp, u64, i32],
        "slg_rerank": [vp, vp, u32, u32, vp, vp, u32, f32, i32, vp, vp],
        "slg_get_counters": [vp, C.POINTER(Counters)],
        "slg_batch_copy_results_device": [vp, vp, vp],
        "slg_get_stream": [vp, C.POINTER(vp)],
        "slg_selftest_div": [vp, u64, u64, C.POINTER(u64)],
        "slg_batch_enable_stats": [vp, i32],
        "slg_set_option": [vp, C.c_char_p, u64],
        "slg_term_has_column": [vp, u32, u32],
        "slg_inspect_segment_files": [C.POINTER(SegmentFiles), C.c_char_p, C.POINTER(SegmentInfo), C.c_char_p, u64],
        "slg_load_segment_files": [vp, C.POINTER(SegmentFiles), C.c_char_p, f32, f32],
        "slg_load_index_dir": [vp, C.c_char_p, C.c_char_p, f32, f32, C.c_char_p, i32, C.POINTER(u32)],
        "slg_load_index_dir_shard": [vp, C.c_char_p, C.c_char_p, f32, f32, C.c_char_p, i32, u32, u32, C.POINTER(u32)],
        "slg_load_vector_file": [vp, u32, vp, u64, i32, C.POINTER(i32)],
        "slg_term_lookup": [vp, C.c_char_p, C.POINTER(u32)],
        "slg_column_lookup": [vp, C.c_char_p],
        "slg_field_stats": [vp, u32, u32, C.POINTER(f32), C.POINTER(f32)],
        "slg_load_positions": [vp, u32, vp, vp, vp, i32],
        "slg_phrase_compile": [vp, vp, u32, u32],
        "slg_phrase_compile_batch": [vp, vp, vp, vp, u32, vp],
        "slg_filter_combine": [vp, u32, i32, i32],
        "slg_filter_combine_batch": [vp, u32, vp, vp, u32, vp],
        "slg_filter_free": [vp, i32],
        "slg_batch_run_seeds": [vp],
        "slg_batch_threshold_keys": [vp, C.POINTER(vp)],
        "slg_batch_import_thresholds": [vp, vp],
        "slg_batch_run_sweep": [vp, i32],
        "slg_batch_packed_results": [vp, C.POINTER(vp), C.POINTER(u64)],
        "slg_merge_gathered_packed": [vp, vp, u64, u32, u32, u32, vp, vp],
        "slg_batch_set_threshold_board": [vp, vp, vp, u32, u32],
        "slg_segment_residency": [vp, u32, C.c_char_p, u64],
        "slg_load_vectors_bf16": [vp, u32, u32, vp, vp, u64],
        "slg_rerank_clauses": [vp, vp, u32, u32, u32, vp, vp, u32, vp, vp, vp],
        "slg_rerank_batch": [vp, vp, u32, u32, i32],
        "slg_batch_fetch_vector_scores": [vp, vp],
        "slg_merge_gathered_hybrid": [vp, vp, u64, u32, u32, u32, vp, vp, vp],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int32
    _LIB = lib
    return lib


def segment_files_struct(segment_ord: int, doc_count: int, terms: bytes, post: bytes, fast: bytes, meta: bytes,
                         deleted_docs=None, checksums=None):
    """(slg_segment_files_t, keepalive) over byte strings / uint8 arrays"""
    keep = [np.frombuffer(x, dtype=np.uint8) if not isinstance(x, np.ndarray) else np.ascontiguousarray(x, dtype=np.uint8)
            for x in (terms, post, fast, meta)]
    f = SegmentFiles()
    f.segment_ord, f.doc_count = segment_ord, doc_count
    for name, a in zip(("terms", "post", "fast", "meta"), keep):
        setattr(f, name, a.ctypes.data if a.size else 0)
        setattr(f, name + "_bytes", a.size)
    if deleted_docs is not None and len(deleted_docs):
        d = np.ascontiguousarray(deleted_docs, dtype=np.uint32)
        keep.append(d)
        f.deleted_docs, f.n_deleted = d.ctypes.data, len(d)
    if checksums is not None:
        c = np.ascontiguousarray(checksums, dtype=np.uint32)
        assert len(c) == 4, "checksums = crc32 of (terms, postings, fast, meta)"
        keep.append(c)
        f.checksums = c.ctypes.data
    return f, keep


def inspect_segment_files(doc_count: int, terms: bytes, post: bytes, fast: bytes, meta: bytes, field: str, checksums=None) -> dict:
    """Host-only validation of a segment's files (no device needed): raises SearchliteGpuError on a bad file."""
    lib = load_library()
    f, keep = segment_files_struct(0, doc_count, terms, post, fast, meta, None, checksums)
    info = SegmentInfo()
    err = C.create_string_buffer(512)
    rc = lib.slg_inspect_segment_files(C.byref(f), field.encode(), C.byref(info), err, 512)
    del keep
    if rc:
        raise SearchliteGpuError(rc, err.value.decode(errors="replace"))
    return {n: getattr(info, n) for n, _ in SegmentInfo._fields_}


def _ptr(a) -> int:
    """address of a numpy array / torch tensor / None"""
    if a is None:
        return 0
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a.data_ptr())  # torch tensor


@dataclass
class SegmentData:
    """One segment as SegmentReader holds it (index/segment.rs:1239): CSR postings + `_len:` column.

    Arrays are numpy (host) or torch CUDA tensors (device); all five must live in one space."""
    segment_ord: int
    doc_count: int
    term_offsets: object   # u64/int64 [n_terms+1]
    post_docs: object      # u32/int32 [n_postings]
    post_tfs: object       # u32/int32 [n_postings]
    field_lengths: object  # int64 [doc_count]
    total_tokens: int
    field_length_present: object = None
    deleted_docs: Optional[np.ndarray] = None
    fast_i64: dict = field(default_factory=dict)   # name -> (values int64, present u8|None)
    fast_f64: dict = field(default_factory=dict)
    fast_str: dict = field(default_factory=dict)   # name -> (dict list[str], ords u32)
    # list columns (I64List / F64List / StrList, index/fastfields.rs:926-1068): offsets u32 [doc_count+1], values
    fast_i64_list: dict = field(default_factory=dict)   # name -> (offsets, values int64)
    fast_f64_list: dict = field(default_factory=dict)   # name -> (offsets, values float64)
    fast_str_list: dict = field(default_factory=dict)   # name -> (dict list[str], offsets, ords u32)

    @property
    def n_terms(self) -> int:
        return int(self.term_offsets.shape[0]) - 1

    def is_device(self) -> bool:
        return not isinstance(self.term_offsets, np.ndarray)

    def to_host(self) -> "SegmentData":
        if not self.is_device():
            return self
        cv = lambda t: None if t is None else t.cpu().numpy()
        return SegmentData(self.segment_ord, self.doc_count, cv(self.term_offsets), cv(self.post_docs), cv(self.post_tfs),
                           cv(self.field_lengths), self.total_tokens, cv(self.field_length_present), self.deleted_docs,
                           self.fast_i64, self.fast_f64, self.fast_str, self.fast_i64_list, self.fast_f64_list, self.fast_str_list)


@dataclass
class QueryBatch:
    """Flat batch of queries as search_segment derives them (api/reader.rs:2971-3002): per query the
    merged "field:term" keys with weights and leaves, the flat matcher (term-group roles +
    minimum_should_match, api/reader.rs:1485-1565) and an optional root filter id."""
    term_off: np.ndarray                  # int64 [Q+1]
    terms: np.ndarray                     # TERM_DTYPE [T]
    group_off: Optional[np.ndarray] = None  # int64 [Q+1]; None = plain OR queries
    group_role: Optional[np.ndarray] = None  # u8 [G]
    min_should: Optional[np.ndarray] = None  # u32 [Q]
    filter_id: Optional[np.ndarray] = None   # i32 [Q]
    plan_off: Optional[np.ndarray] = None    # [Q+1] into plan_nodes (ScorePlan per query; empty = running sum)
    plan_nodes: Optional[np.ndarray] = None  # PLAN_DTYPE
    leaf_count: Optional[np.ndarray] = None  # u32 [Q]
    cursor: Optional[np.ndarray] = None      # HIT_DTYPE [Q]: search-after keys (SearchRequest.cursor)
    has_cursor: Optional[np.ndarray] = None  # u8 [Q]
    _structs: Optional[np.ndarray] = None

    def set_cursors(self, cursors: Sequence) -> "QueryBatch":
        """one (segment_ord, doc_id, score) triple or HIT_DTYPE row per query; None = no cursor"""
        self.cursor = np.zeros(self.n_queries, dtype=HIT_DTYPE)
        self.has_cursor = np.zeros(self.n_queries, dtype=np.uint8)
        for qi, c in enumerate(cursors):
            if c is not None:
                self.cursor[qi] = (int(c[0]), int(c[1]), np.float32(c[2]))
                self.has_cursor[qi] = 1
        self._structs = None
        return self

    def set_plans(self, exprs: Sequence) -> "QueryBatch":
        """attach one ScoreExpr per query (None = no plan); leaf_count = 1 + the largest leaf of the query's terms"""
        off, rows, leaves = [0], [], []
        for qi, e in enumerate(exprs):
            if e is not None:
                rows += plan_postfix(e)
            off.append(len(rows))
            t = self.terms[int(self.term_off[qi]):int(self.term_off[qi + 1])]
            leaves.append(0 if e is None else (int(t["leaf"].max()) + 1 if len(t) else 1))
        self.plan_off = np.array(off, dtype=np.int64)
        self.plan_nodes = np.array(rows, dtype=PLAN_DTYPE) if rows else np.zeros(0, dtype=PLAN_DTYPE)
        self.leaf_count = np.array(leaves, dtype=np.uint32)
        self._structs = None
        return self

    @property
    def n_queries(self) -> int:
        return len(self.term_off) - 1

    @staticmethod
    def from_term_lists(term_lists: Sequence[Sequence[int]], weights: Optional[Sequence[Sequence[float]]] = None) -> "QueryBatch":
        """plain OR queries (QueryString, minimum_should_match 1): each term is its own leaf"""
        off = np.zeros(len(term_lists) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(t) for t in term_lists])
        terms = np.zeros(int(off[-1]), dtype=TERM_DTYPE)
        p = 0
        for qi, tl in enumerate(term_lists):
            for j, t in enumerate(tl):
                terms[p] = (t, 1.0 if weights is None else weights[qi][j], j, 0, TERM_SCORED)
                p += 1
        return QueryBatch(off, terms)

    @staticmethod
    def from_bool(queries: Sequence[dict]) -> "QueryBatch":
        """Bool queries: each dict has `must`, `should`, `must_not` (lists of term ids; one group per
        term), optional `min_should` (defaults as api/reader.rs:1553-1561), optional `filter_id`."""
        toff, goff = [0], [0]
        trows, roles, mins, fids = [], [], [], []
        for q in queries:
            g = 0
            leaf = 0
            for role, key in ((ROLE_MUST, "must"), (ROLE_SHOULD, "should"), (ROLE_MUST_NOT, "must_not")):
                for t in q.get(key, []):
                    scored = role != ROLE_MUST_NOT
                    trows.append((t, 1.0, leaf if scored else 0, g, TERM_SCORED if scored else 0))
                    if scored:
                        leaf += 1
                    roles.append(role)
                    g += 1
            ms = q.get("min_should")
            if ms is None:
                ms = 0 if (not q.get("should") or q.get("must") or q.get("filter_id", -1) >= 0) else 1
            mins.append(ms)
            fids.append(q.get("filter_id", -1))
            toff.append(len(trows))
            goff.append(len(roles))
        terms = np.array(trows, dtype=TERM_DTYPE) if trows else np.zeros(0, dtype=TERM_DTYPE)
        return QueryBatch(np.array(toff, dtype=np.int64), terms, np.array(goff, dtype=np.int64),
                          np.array(roles, dtype=np.uint8), np.array(mins, dtype=np.uint32), np.array(fids, dtype=np.int32))

    def structs(self) -> np.ndarray:
        """array of slg_query_t (same layout as the oracle's slo_query_t plus filter_id)"""
        if self._structs is not None:
            return self._structs
        q = self.n_queries
        self.terms = np.ascontiguousarray(self.terms)
        s = np.zeros(q, dtype=QUERY_DTYPE)
        s["n_terms"] = (self.term_off[1:] - self.term_off[:-1]).astype(np.uint32)
        s["terms"] = self.terms.ctypes.data + self.term_off[:-1].astype(np.uint64) * np.uint64(TERM_DTYPE.itemsize)
        if self.group_off is not None:
            self.group_role = np.ascontiguousarray(self.group_role, dtype=np.uint8)
            s["n_groups"] = (self.group_off[1:] - self.group_off[:-1]).astype(np.uint32)
            s["group_role"] = self.group_role.ctypes.data + self.group_off[:-1].astype(np.uint64)
            s["min_should"] = self.min_should
        else:
            s["n_groups"] = 0
            s["group_role"] = 0
            s["min_should"] = 1
        s["leaf_count"] = 0
        s["filter_id"] = -1 if self.filter_id is None else self.filter_id
        if self.plan_off is not None:
            self.plan_nodes = np.ascontiguousarray(self.plan_nodes)
            n = (self.plan_off[1:] - self.plan_off[:-1]).astype(np.uint32)
            s["n_plan_nodes"] = n
            s["plan"] = np.where(n > 0, self.plan_nodes.ctypes.data + self.plan_off[:-1].astype(np.uint64) * np.uint64(PLAN_DTYPE.itemsize), 0)
            s["leaf_count"] = self.leaf_count
        if self.cursor is not None:
            s["has_cursor"] = self.has_cursor
            s["cursor_segment_ord"] = self.cursor["segment_ord"]
            s["cursor_doc_id"] = self.cursor["doc_id"]
            s["cursor_score"] = self.cursor["score"]
        self._structs = s
        return s

    def subset(self, lo: int, hi: int) -> "QueryBatch":
        t0, t1 = int(self.term_off[lo]), int(self.term_off[hi])
        qb = QueryBatch(self.term_off[lo:hi + 1] - t0, self.terms[t0:t1].copy())
        if self.group_off is not None:
            g0, g1 = int(self.group_off[lo]), int(self.group_off[hi])
            qb.group_off = self.group_off[lo:hi + 1] - g0
            qb.group_role = self.group_role[g0:g1].copy()
            qb.min_should = self.min_should[lo:hi].copy()
        if self.filter_id is not None:
            qb.filter_id = self.filter_id[lo:hi].copy()
        if self.plan_off is not None:
            p0, p1 = int(self.plan_off[lo]), int(self.plan_off[hi])
            qb.plan_off = self.plan_off[lo:hi + 1] - p0
            qb.plan_nodes = self.plan_nodes[p0:p1].copy()
            qb.leaf_count = self.leaf_count[lo:hi].copy()
        if self.cursor is not None:
            qb.cursor = self.cursor[lo:hi].copy()
            qb.has_cursor = self.has_cursor[lo:hi].copy()
        return qb


def vector_clauses(clauses):
    """[(query_vecs [Q, dim] f32 numpy array or CUDA tensor, alpha, boost, metric name)] -> (slg_vector_clause_t array, keepalive, dim)"""
    arr = (VectorClause * len(clauses))()
    keep, dim = [], 0
    for i, (qv, alpha, boost, metric) in enumerate(clauses):
        if isinstance(qv, np.ndarray):
            qv = np.ascontiguousarray(qv, dtype=np.float32)
            ptr = qv.ctypes.data
        else:  # torch tensor (host or device)
            qv = qv.contiguous().float()
            ptr = qv.data_ptr()
        keep.append(qv)
        dim = int(qv.shape[1])
        arr[i].query_vecs = ptr
        arr[i].alpha, arr[i].boost, arr[i].metric = float(alpha), float(boost), METRIC[metric]
    return arr, keep, dim


class PreparedBatch:
    """A query batch resident on the device (slg_batch_prepare)."""

    def __init__(self, index: "GpuIndex", handle: int, n_queries: int, k: int, keepalive, execution: str = "bm25"):
        self.index, self.handle, self.n_queries, self.k = index, handle, n_queries, k
        self._keepalive = keepalive
        self.execution = execution
        # a batch the posting scan (or the pruned items kernel) takes can run in two steps — first part, exchange of the
        # per-query k-th keys between shards, rest; the first refusal (SLG_ERR_UNSUPPORTED) switches this off
        self.two_step_ok = True

    def enable_stats(self, on: bool = True) -> None:
        self.index._check(self.index.lib.slg_batch_enable_stats(self.handle, 1 if on else 0))

    def run(self, sync: bool = True) -> None:
        self.index._check(self.index.lib.slg_batch_run(self.handle, 1 if sync else 0))

    def fetch(self, want_stats: bool = False):
        hits = np.zeros((self.n_queries, self.k), dtype=HIT_DTYPE)
        counts = np.zeros(self.n_queries, dtype=np.uint32)
        stats = np.zeros(self.n_queries, dtype=STATS_DTYPE) if want_stats else None
        self.index._check(self.index.lib.slg_batch_fetch(self.handle, _ptr(hits), _ptr(counts), _ptr(stats)))
        return (hits, counts, stats) if want_stats else (hits, counts)

    # two-step pruned run for one-segment-per-GPU sharding: seeds -> threshold exchange -> sweep
    def run_seeds(self) -> bool:
        """False (and nothing was run) when this batch cannot run in two steps"""
        rc = self.index.lib.slg_batch_run_seeds(self.handle)
        if rc == -4:  # SLG_ERR_UNSUPPORTED
            self.two_step_ok = False
            return False
        self.index._check(rc)
        return True

    def threshold_keys_ptr(self) -> int:
        p = C.c_void_p()
        self.index._check(self.index.lib.slg_batch_threshold_keys(self.handle, C.byref(p)))
        return p.value

    def import_thresholds(self, dev_keys_ptr: int) -> None:
        self.index._check(self.index.lib.slg_batch_import_thresholds(self.handle, dev_keys_ptr))

    def run_sweep(self, sync: bool = True) -> None:
        self.index._check(self.index.lib.slg_batch_run_sweep(self.handle, 1 if sync else 0))

    def rerank(self, clauses, sync: bool = True) -> None:
        """slg_rerank_batch: hybrid rescoring of the device-resident top-k of the last run; clauses as vector_clauses()"""
        arr, keep, dim = vector_clauses(clauses)
        self.index._check(self.index.lib.slg_rerank_batch(self.handle, arr, len(clauses), dim, 1 if sync else 0))
        self._rerank_keepalive = keep

    def fetch_vector_scores(self) -> np.ndarray:
        vs = np.zeros((self.n_queries, self.k), dtype=np.float32)
        self.index._check(self.index.lib.slg_batch_fetch_vector_scores(self.handle, _ptr(vs)))
        return vs

    def set_threshold_board(self, local_ptr: int, peer_ptrs, epoch: int) -> None:
        """slg_batch_set_threshold_board: device pointers of this shard's board and of the peers' boards (peer mappings)"""
        arr = (C.c_void_p * max(len(peer_ptrs), 1))(*peer_ptrs)
        self.index._check(self.index.lib.slg_batch_set_threshold_board(self.handle, local_ptr, arr, len(peer_ptrs), epoch))

    def packed_results(self):
        """(device pointer, bytes) of the last run's result block: n_queries*k hits, then n_queries counts (then, after a
        rerank, n_queries*k vector scores)"""
        p, n = C.c_void_p(), C.c_uint64()
        self.index._check(self.index.lib.slg_batch_packed_results(self.handle, C.byref(p), C.byref(n)))
        return p.value, n.value

    def device_results(self):
        dh, dc = C.c_void_p(), C.c_void_p()
        self.index._check(self.index.lib.slg_batch_device_results(self.handle, C.byref(dh), C.byref(dc)))
        return dh.value, dc.value

    def copy_results_to(self, dst_hits_ptr: int, dst_counts_ptr: int) -> None:
        self.index._check(self.index.lib.slg_batch_copy_results_device(self.handle, dst_hits_ptr, dst_counts_ptr))

    def cursor_seen(self) -> np.ndarray:
        """saw_cursor per query of the last run (api/reader.rs:2663, :3022-3024); False means the reference would fail
        the request with 'stale or invalid cursor for this result set'"""
        out = np.zeros(self.n_queries, dtype=np.uint8)
        self.index._check(self.index.lib.slg_batch_cursor_seen(self.handle, _ptr(out)))
        return out.astype(bool)

    def free(self) -> None:
        if self.handle:
            self.index.lib.slg_batch_free(self.handle)
            self.handle = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def cursor_encode(generation: int, returned: int, hit) -> str:
    """PaginationCursor::encode (api/reader.rs:630-650): hit = (segment_ord, doc_id, score) of the last returned hit"""
    h = np.zeros(1, dtype=HIT_DTYPE)
    h[0] = (int(hit[0]), int(hit[1]), np.float32(hit[2]))
    buf = C.create_string_buffer(43)
    rc = load_library().slg_cursor_encode(generation, returned, _ptr(h), buf)
    if rc:
        raise SearchliteGpuError(rc, "cursor encode failed")
    return buf.value.decode()


def cursor_decode(raw: str, manifest_generation: int):
    """PaginationCursor::decode + the generation check of decode_cursor (api/reader.rs:652-691, :821-841) ->
    ((segment_ord, doc_id, score), returned); raises with the reference's message"""
    h = np.zeros(1, dtype=HIT_DTYPE)
    ret = C.c_uint32()
    err = C.create_string_buffer(256)
    rc = load_library().slg_cursor_decode(raw.encode(), manifest_generation, _ptr(h), C.byref(ret), err, 256)
    if rc:
        raise SearchliteGpuError(rc, err.value.decode())
    return (int(h[0]["segment_ord"]), int(h[0]["doc_id"]), np.float32(h[0]["score"])), int(ret.value)


class GpuIndex:
    """Device-resident index: the GPU stand-in for Index::reader() + IndexReader::search on the
    BM25 top-k path.  One instance owns one CUDA device/stream."""

    # "items": the posting-driven kernel is required (an unsupported batch is an error instead of a fallback);
    # "reg" is its round-1 name
    KERNEL = {"auto": 0, "cta": 1, "warp": 2, "items": 3, "reg": 3, "warp-inplace": 2 + 256, "auto-inplace": 256}

    def __init__(self, device: int = 0, tile_docs: int = 0, ctas_per_sm: int = 0, sub_docs: int = 0, kernel: str = "auto",
                 options: Optional[dict] = None):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.slg_open(device, C.byref(h))
        if rc != 0:
            raise SearchliteGpuError(rc, (self.lib.slg_last_error(None) or b"").decode())
        self.handle = h
        self.device = device
        self._keep = []
        if tile_docs or ctas_per_sm or sub_docs or kernel != "auto":
            self._check(self.lib.slg_configure(self.handle, tile_docs, ctas_per_sm, sub_docs, self.KERNEL[kernel]))
        for name, value in (options or {}).items():
            self.set_option(name, value)

    def set_option(self, name: str, value: int) -> None:
        """residency / tuning options (slg_set_option); applies to segments loaded afterwards"""
        self._check(self.lib.slg_set_option(self.handle, name.encode(), int(value)))

    def term_has_column(self, segment_ord: int, term_id: int) -> bool:
        return bool(self._check(self.lib.slg_term_has_column(self.handle, segment_ord, int(term_id))))

    def _check(self, rc: int) -> int:
        if rc < 0:
            raise SearchliteGpuError(rc, (self.lib.slg_last_error(self.handle) or b"").decode())
        return rc

    def close(self) -> None:
        if self.handle:
            self.lib.slg_close(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- residency -------------------------------------------------------------------------
    def _view(self, seg: SegmentData, with_postings: bool = True) -> SegmentView:
        dev = seg.is_device()
        v = SegmentView()
        v.segment_ord, v.doc_count, v.n_terms = seg.segment_ord, seg.doc_count, seg.n_terms
        if with_postings:
            for name in ("term_offsets", "post_docs", "post_tfs"):
                a = getattr(seg, name)
                if isinstance(a, np.ndarray):
                    want = np.uint64 if name == "term_offsets" else np.uint32
                    if a.dtype.itemsize != np.dtype(want).itemsize:
                        a = a.astype(want)
                    a = np.ascontiguousarray(a)
                    setattr(seg, name, a)
                setattr(v, name, _ptr(a))
        if isinstance(seg.field_lengths, np.ndarray):
            seg.field_lengths = np.ascontiguousarray(seg.field_lengths, dtype=np.int64)
        v.field_lengths = _ptr(seg.field_lengths)
        v.field_length_present = _ptr(seg.field_length_present)
        v.total_tokens = int(seg.total_tokens)
        if seg.deleted_docs is not None and len(seg.deleted_docs):
            seg.deleted_docs = np.ascontiguousarray(seg.deleted_docs, dtype=np.uint32)
            v.deleted_docs = _ptr(seg.deleted_docs)
            v.n_deleted = len(seg.deleted_docs)
        v.memory_space = MEM_DEVICE if dev else MEM_HOST
        return v

    def load_segment(self, seg: SegmentData, k1: float = 0.9, b: float = 0.4) -> dict:
        """SegmentReader::open for the hot path; also registers the segment's fast-field columns.
        Returns {column name: handle}."""
        if seg.is_device():
            import torch
            torch.cuda.synchronize()
        v = self._view(seg)
        self._check(self.lib.slg_load_segment(self.handle, C.byref(v), k1, b))
        return self._add_columns(seg)

    def load_segment_post_image(self, seg: SegmentData, post_image: np.ndarray, post_offsets: np.ndarray,
                                k1: float = 0.9, b: float = 0.4) -> dict:
        """Load from the reference's `.post` byte image (index/postings.rs:78-129)."""
        host = seg.to_host()
        v = self._view(host, with_postings=False)
        v.memory_space = MEM_HOST
        post_image = np.ascontiguousarray(post_image, dtype=np.uint8)
        post_offsets = np.ascontiguousarray(post_offsets, dtype=np.uint64)
        self._check(self.lib.slg_load_segment_post_image(self.handle, C.byref(v), _ptr(post_image), post_image.nbytes,
                                                         _ptr(post_offsets), k1, b))
        return self._add_columns(host)

    def load_segment_files(self, segment_ord: int, doc_count: int, terms: bytes, post: bytes, fast: bytes, meta: bytes,
                           field: str, deleted_docs=None, checksums=None, k1: float = 0.9, b: float = 0.4) -> None:
        """SegmentReader::open from the reference's own files (index/segment.rs:1239-1330), given as bytes."""
        f, keep = segment_files_struct(segment_ord, doc_count, terms, post, fast, meta, deleted_docs, checksums)
        self._check(self.lib.slg_load_segment_files(self.handle, C.byref(f), field.encode(), k1, b))
        del keep

    def load_index_dir(self, path: str, field: str, k1: float = 0.9, b: float = 0.4, vector_field: Optional[str] = None,
                       store_bf16: bool = False, shard_rank: int = 0, shard_world: int = 1) -> int:
        """Every segment of an index directory in MANIFEST.json order (or, one process per GPU, the segments with
        position % shard_world == shard_rank); returns the number of segments loaded."""
        n = C.c_uint32()
        vf = vector_field.encode() if vector_field else None
        self._check(self.lib.slg_load_index_dir_shard(self.handle, path.encode(), field.encode(), k1, b, vf, int(store_bf16),
                                                      shard_rank, shard_world, C.byref(n)))
        return n.value

    def term_lookup(self, key: str) -> int:
        """term id of a "field:token" key in the handle's term space (ABSENT_TERM if unknown)"""
        t = C.c_uint32()
        self._check(self.lib.slg_term_lookup(self.handle, key.encode(), C.byref(t)))
        return t.value

    def column_lookup(self, name: str) -> int:
        return int(self.lib.slg_column_lookup(self.handle, name.encode()))

    def load_positions(self, segment_ord: int, term_offsets, position_offsets, positions) -> None:
        """numpy arrays (host) or torch CUDA tensors (int64 offsets, int32 positions), all in one space"""
        if isinstance(term_offsets, np.ndarray):
            to = np.ascontiguousarray(term_offsets, dtype=np.uint64)
            po = np.ascontiguousarray(position_offsets, dtype=np.uint64)
            ps = np.ascontiguousarray(positions, dtype=np.uint32)
            space = MEM_HOST
        else:
            import torch
            to, po, ps = term_offsets.contiguous(), position_offsets.contiguous(), positions.contiguous()
            assert to.dtype == torch.int64 and po.dtype == torch.int64 and ps.dtype == torch.int32
            torch.cuda.synchronize()
            space = MEM_DEVICE
        self._check(self.lib.slg_load_positions(self.handle, segment_ord, _ptr(to), _ptr(po), _ptr(ps), space))

    def compile_phrase(self, term_ids: Sequence[int], slop: int = 0) -> int:
        """matches_phrase (query/phrase.rs:4-48) as a per-segment doc bitmap; the id is used like a filter id"""
        t = np.ascontiguousarray(term_ids, dtype=np.uint32)
        return self._check(self.lib.slg_phrase_compile(self.handle, _ptr(t), len(t), slop))

    def compile_phrases(self, phrases: Sequence[Sequence[int]], slops: Optional[Sequence[int]] = None) -> np.ndarray:
        """every phrase of a query batch in one launch per segment; returns their (consecutive) ids"""
        off = np.zeros(len(phrases) + 1, dtype=np.uint32)
        off[1:] = np.cumsum([len(p) for p in phrases])
        flat = np.ascontiguousarray([t for p in phrases for t in p], dtype=np.uint32)
        sl = None if slops is None else np.ascontiguousarray(slops, dtype=np.uint32)
        out = np.zeros(len(phrases), dtype=np.int32)
        self._check(self.lib.slg_phrase_compile_batch(self.handle, _ptr(flat), _ptr(off), _ptr(sl), len(phrases), _ptr(out)))
        return out

    def combine_filters(self, op: str, a: int, b: int) -> int:
        return self._check(self.lib.slg_filter_combine(self.handle, COMBINE[op], a, b))

    def combine_filters_batch(self, op: str, a: Sequence[int], b: Sequence[int]) -> np.ndarray:
        """out[i] = a[i] op b[i], one launch per segment"""
        aa, bb = np.ascontiguousarray(a, dtype=np.int32), np.ascontiguousarray(b, dtype=np.int32)
        assert len(aa) == len(bb)
        out = np.zeros(len(aa), dtype=np.int32)
        self._check(self.lib.slg_filter_combine_batch(self.handle, COMBINE[op], _ptr(aa), _ptr(bb), len(aa), _ptr(out)))
        return out

    def free_filter(self, filter_id: int) -> None:
        self._check(self.lib.slg_filter_free(self.handle, filter_id))

    def _add_columns(self, seg: SegmentData) -> dict:
        handles = {}
        for name, (vals, present) in seg.fast_i64.items():
            vals = np.ascontiguousarray(vals, dtype=np.int64)
            pres = None if present is None else np.ascontiguousarray(present, dtype=np.uint8)
            handles[name] = self._check(self.lib.slg_add_i64_column(self.handle, seg.segment_ord, _ptr(vals), _ptr(pres)))
        for name, (vals, present) in seg.fast_f64.items():
            vals = np.ascontiguousarray(vals, dtype=np.float64)
            pres = None if present is None else np.ascontiguousarray(present, dtype=np.uint8)
            handles[name] = self._check(self.lib.slg_add_f64_column(self.handle, seg.segment_ord, _ptr(vals), _ptr(pres)))
        for name, (dic, ords) in seg.fast_str.items():
            ords = np.ascontiguousarray(ords, dtype=np.uint32)
            arr = (C.c_char_p * len(dic))(*[s.encode() for s in dic])
            handles[name] = self._check(self.lib.slg_add_str_column(self.handle, seg.segment_ord, arr, len(dic), _ptr(ords)))
        for name, (offs, vals) in seg.fast_i64_list.items():
            offs = np.ascontiguousarray(offs, dtype=np.uint32)
            vals = np.ascontiguousarray(vals, dtype=np.int64)
            handles[name] = self._check(self.lib.slg_add_i64_list_column(self.handle, seg.segment_ord, _ptr(offs), _ptr(vals)))
        for name, (offs, vals) in seg.fast_f64_list.items():
            offs = np.ascontiguousarray(offs, dtype=np.uint32)
            vals = np.ascontiguousarray(vals, dtype=np.float64)
            handles[name] = self._check(self.lib.slg_add_f64_list_column(self.handle, seg.segment_ord, _ptr(offs), _ptr(vals)))
        for name, (dic, offs, ords) in seg.fast_str_list.items():
            offs = np.ascontiguousarray(offs, dtype=np.uint32)
            ords = np.ascontiguousarray(ords, dtype=np.uint32)
            arr = (C.c_char_p * len(dic))(*[s.encode() for s in dic])
            handles[name] = self._check(self.lib.slg_add_str_list_column(self.handle, seg.segment_ord, arr, len(dic), _ptr(offs), _ptr(ords)))
        return handles

    def segment_stats(self, segment_ord: int) -> dict:
        a, l, m, n = C.c_float(), C.c_float(), C.c_float(), C.c_uint64()
        self._check(self.lib.slg_segment_stats(self.handle, segment_ord, C.byref(a), C.byref(l), C.byref(m), C.byref(n)))
        return {"avgdl": a.value, "live_docs": l.value, "min_doc_len": m.value, "n_postings": n.value}

    def segment_residency(self, segment_ord: int) -> dict:
        """device bytes per resident array of the segment"""
        import json
        buf = C.create_string_buffer(2048)
        self._check(self.lib.slg_segment_residency(self.handle, segment_ord, buf, 2048))
        return json.loads(buf.value.decode())

    def field_stats(self, segment_ord: int, field_index: int) -> dict:
        a, m = C.c_float(), C.c_float()
        self._check(self.lib.slg_field_stats(self.handle, segment_ord, field_index, C.byref(a), C.byref(m)))
        return {"avgdl": a.value, "min_doc_len": m.value}

    # ---- filters ---------------------------------------------------------------------------
    def compile_filter(self, nodes: np.ndarray, strings: Sequence[str] = ()) -> int:
        nodes = np.ascontiguousarray(nodes, dtype=FILTER_DTYPE)
        arr = (C.c_char_p * max(len(strings), 1))(*[s.encode() for s in strings])
        return self._check(self.lib.slg_filter_compile(self.handle, _ptr(nodes), len(nodes), arr))

    def filter_bitmap(self, filter_id: int, segment_ord: int, doc_count: int) -> np.ndarray:
        out = np.zeros((doc_count + 31) // 32, dtype=np.uint32)
        self._check(self.lib.slg_filter_bitmap(self.handle, filter_id, segment_ord, _ptr(out)))
        return out

    # ---- search ----------------------------------------------------------------------------
    def search_batch(self, batch: QueryBatch, k: int, execution: str = "bm25", bmw_block_size: int = 0,
                     want_stats: bool = False):
        """One call, host buffers in and out (the call a searchlite-core shim would make per batch)."""
        s = batch.structs()
        hits = np.zeros((batch.n_queries, k), dtype=HIT_DTYPE)
        counts = np.zeros(batch.n_queries, dtype=np.uint32)
        stats = np.zeros(batch.n_queries, dtype=STATS_DTYPE) if want_stats else None
        self._check(self.lib.slg_search_batch(self.handle, _ptr(s), batch.n_queries, k, EXECUTION[execution],
                                              bmw_block_size, _ptr(hits), _ptr(counts), _ptr(stats)))
        return (hits, counts, stats) if want_stats else (hits, counts)

    def prepare(self, batch: QueryBatch, k: int, execution: str = "bm25", bmw_block_size: int = 0) -> PreparedBatch:
        s = batch.structs()
        h = C.c_void_p()
        self._check(self.lib.slg_batch_prepare(self.handle, _ptr(s), batch.n_queries, k, EXECUTION[execution],
                                               bmw_block_size, C.byref(h)))
        return PreparedBatch(self, h, batch.n_queries, k, batch, execution)

    def merge_gathered(self, dev_hits_ptr: int, dev_counts_ptr: int, n_shards: int, n_queries: int, k: int):
        hits = np.zeros((n_queries, k), dtype=HIT_DTYPE)
        counts = np.zeros(n_queries, dtype=np.uint32)
        self._check(self.lib.slg_merge_gathered(self.handle, dev_hits_ptr, dev_counts_ptr, n_shards, n_queries, k,
                                                _ptr(hits), _ptr(counts)))
        return hits, counts

    def merge_gathered_packed(self, dev_blocks_ptr: int, n_shards: int, n_queries: int, k: int, shard_stride: int = 0):
        hits = np.zeros((n_queries, k), dtype=HIT_DTYPE)
        counts = np.zeros(n_queries, dtype=np.uint32)
        self._check(self.lib.slg_merge_gathered_packed(self.handle, dev_blocks_ptr, shard_stride, n_shards, n_queries, k,
                                                       _ptr(hits), _ptr(counts)))
        return hits, counts

    def merge_gathered_hybrid(self, dev_blocks_ptr: int, n_shards: int, n_queries: int, k: int, shard_stride: int = 0):
        """merge of reranked blocks (hits, counts, vector scores) -> (hits, counts, vector scores)"""
        hits = np.zeros((n_queries, k), dtype=HIT_DTYPE)
        counts = np.zeros(n_queries, dtype=np.uint32)
        vs = np.zeros((n_queries, k), dtype=np.float32)
        self._check(self.lib.slg_merge_gathered_hybrid(self.handle, dev_blocks_ptr, shard_stride, n_shards, n_queries, k,
                                                       _ptr(hits), _ptr(counts), _ptr(vs)))
        return hits, counts, vs

    # ---- vectors ---------------------------------------------------------------------------
    def load_vectors(self, segment_ord: int, offsets, values, store_bf16: bool = False) -> None:
        """numpy arrays (host) or torch CUDA tensors: offsets u32/int32 [doc_count], values f32 or bf16 [n_rows, dim]"""
        if isinstance(values, np.ndarray):
            offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
            values = np.ascontiguousarray(values, dtype=np.float32)
            n_rows, dim = (values.shape if values.ndim == 2 else (0, 0))
            self._check(self.lib.slg_load_vectors(self.handle, segment_ord, dim, _ptr(offsets), _ptr(values), n_rows,
                                                  1 if store_bf16 else 0))
            return
        import torch
        assert offsets.dtype in (torch.int32, torch.uint32) and values.is_contiguous() and offsets.is_contiguous()
        n_rows, dim = values.shape
        torch.cuda.synchronize()
        if values.dtype == torch.bfloat16:
            self._check(self.lib.slg_load_vectors_bf16(self.handle, segment_ord, dim, offsets.data_ptr(), values.data_ptr(), n_rows))
        else:
            assert values.dtype == torch.float32
            self._check(self.lib.slg_load_vectors(self.handle, segment_ord, dim, offsets.data_ptr(), values.data_ptr(), n_rows,
                                                  1 if store_bf16 else 0))

    def rerank_clauses(self, clauses, cands: np.ndarray, cand_counts: np.ndarray):
        """slg_rerank_clauses on host candidates -> (hits, counts, vector scores); clauses as vector_clauses()"""
        arr, keep, dim = vector_clauses(clauses)
        cands = np.ascontiguousarray(cands, dtype=HIT_DTYPE)
        cand_counts = np.ascontiguousarray(cand_counts, dtype=np.uint32)
        nq, stride = cands.shape
        out = np.zeros((nq, stride), dtype=HIT_DTYPE)
        oc = np.zeros(nq, dtype=np.uint32)
        vs = np.zeros((nq, stride), dtype=np.float32)
        self._check(self.lib.slg_rerank_clauses(self.handle, arr, len(clauses), nq, dim, _ptr(cands), _ptr(cand_counts), stride,
                                                _ptr(out), _ptr(oc), _ptr(vs)))
        del keep
        return out, oc, vs

    def rerank(self, query_vecs: np.ndarray, cands: np.ndarray, cand_counts: np.ndarray, alpha: float, metric: str = "cosine"):
        query_vecs = np.ascontiguousarray(query_vecs, dtype=np.float32)
        cands = np.ascontiguousarray(cands, dtype=HIT_DTYPE)
        cand_counts = np.ascontiguousarray(cand_counts, dtype=np.uint32)
        nq, stride = cands.shape
        out = np.zeros((nq, stride), dtype=HIT_DTYPE)
        vs = np.zeros((nq, stride), dtype=np.float32)
        self._check(self.lib.slg_rerank(self.handle, _ptr(query_vecs), nq, query_vecs.shape[1], _ptr(cands), _ptr(cand_counts),
                                        stride, alpha, METRIC[metric], _ptr(out), _ptr(vs)))
        return out, vs

    def selftest_div(self, n: int, seed: int = 1) -> int:
        m = C.c_uint64()
        self._check(self.lib.slg_selftest_div(self.handle, n, seed, C.byref(m)))
        return m.value

    def stream_ptr(self) -> int:
        s = C.c_void_p()
        self._check(self.lib.slg_get_stream(self.handle, C.byref(s)))
        return s.value

    def counters(self) -> dict:
        c = Counters()
        self._check(self.lib.slg_get_counters(self.handle, C.byref(c)))
        return {n: getattr(c, n) for n, _ in Counters._fields_}
