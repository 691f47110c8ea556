"""CPU tests: the C++ oracle against the golden fixtures (tests/golden/, produced by an independent
numpy-float32 restatement) and against the literals / properties of the reference's own tests.

Reference citations are relative to /root/reference/searchlite-core/."""
import ctypes as C

import numpy as np
import pytest

from oracle import slo
from tests import pyref
from tests.helpers import f32_bits, golden, hits_to_list, or_queries, segment_from_postings, token_corpus


@pytest.fixture(scope="module")
def L():
    slo.build()
    return slo.lib()


# ---- scalar arithmetic: src/query/bm25.rs:1-6, src/query/wand.rs:269-303 -------------------------
def test_bm25_and_score_tf_match_golden_bits(L):
    rows = golden("bm25_scalar.json")
    assert len(rows) >= 250
    for r in rows:
        dl = r["doc_len"] if r["doc_len"] > 0 else max(r["avgdl"], r["tf"])
        got = L.slo_bm25(r["tf"], r["df"], dl, r["avgdl"], r["docs"], r["k1"], r["b"])
        assert f32_bits(got) == r["bm25_bits"], r
        got = L.slo_score_tf(r["tf"], r["df"], r["doc_len"], r["avgdl"], r["docs"], r["k1"], r["b"], r["weight"])
        assert f32_bits(got) == r["score_tf_bits"], r


def test_reference_bm25_literals(L):
    # src/query/bm25.rs:13-19
    assert np.isfinite(L.slo_bm25(3.0, 5.0, 100.0, 120.0, 1000.0, 1.2, 0.75))
    assert L.slo_bm25(1.0, 1.0, 0.0, 0.0, 10.0, 1.2, 0.75) > 0.0
    # src/query/wand.rs:1014-1021 bm25_penalizes_long_documents
    short = L.slo_score_tf(2.0, 1.0, 5.0, 10.0, 100.0, 1.2, 0.75, 1.0)
    long_ = L.slo_score_tf(2.0, 1.0, 100.0, 10.0, 100.0, 1.2, 0.75, 1.0)
    assert short > long_
    # upper_bound_tf: 0 when tf <= 0 (wand.rs:289-303)
    assert L.slo_upper_bound_tf(0.0, 5.0, 10.0, 10.0, 100.0, 1.2, 0.75, 1.0) == 0.0
    assert L.slo_upper_bound_tf(2.0, 1.0, 5.0, 10.0, 100.0, 1.2, 0.75, 1.0) == short


def test_idf_is_hoistable(L):
    """the device tables use idf alone; bm25 == idf * (tf*(k1+1)) / max(tf + k1*(1-b+b*dl/avgdl), 1e-6)"""
    f = np.float32
    for r in golden("bm25_scalar.json")[:60]:
        dl = r["doc_len"] if r["doc_len"] > 0 else max(r["avgdl"], r["tf"])
        idf = f(L.slo_idf(r["df"], r["docs"]))
        norm = f(dl) / f(r["avgdl"]) if r["avgdl"] > 0 else f(1.0)
        nk = f(r["k1"]) * (f(1.0) - f(r["b"]) + f(r["b"]) * norm)
        val = idf * (f(r["tf"]) * (f(r["k1"]) + f(1.0))) / max(f(r["tf"]) + nk, f(1e-6))
        assert f32_bits(val) == r["bm25_bits"]


# ---- src/query/wand.rs:952-1011 literals ---------------------------------------------------------
def _literal_index():
    g = golden("wand_literal.json")
    post = [(t["docs"], t["tfs"]) for t in g["terms"]]
    # term_from_entries: lengths 10.0 for docs 0..max, avgdl 10, docs 10 -> 10 docs of length 10
    seg = segment_from_postings(post, [10] * 10)
    return g, slo.OracleIndex(seg, k1=g["k1"], b=g["b"])


def test_brute_force_matches_wand_on_reference_literal():
    g, ora = _literal_index()
    assert ora.avgdl == 10.0 and ora.live_docs == 10.0
    qb = or_queries([[0, 1]])
    want = [(e["doc_id"], e["score_bits"]) for e in g["expected"]]
    for mode in ("bm25", "bm25_dense", "wand", "bmw"):
        h, c = ora.search_batch(qb, g["k"], mode)
        assert hits_to_list(h, c) == want, mode


def test_tie_break_prefers_smaller_doc_id():
    # wand.rs:952-966: equal scores -> the larger doc id is the worst; top-1 keeps the smaller id
    seg = segment_from_postings([([1, 2], [1, 1])], [10] * 4)
    ora = slo.OracleIndex(seg, k1=1.2, b=0.75)
    for mode in ("bm25", "wand", "bmw", "bm25_dense"):
        h, c = ora.search_batch(or_queries([[0]]), 1, mode)
        assert c[0] == 1 and h[0][0]["doc_id"] == 1
        h, c = ora.search_batch(or_queries([[0]]), 2, mode)
        assert [int(x) for x in h[0]["doc_id"][:2]] == [1, 2]
        assert h[0]["score"][0] == h[0]["score"][1]


def test_small_corpus_topk_matches_golden():
    g = golden("small_topk.json")
    seg = segment_from_postings([(p["docs"], p["tfs"]) for p in g["postings"]], g["field_lengths"], g["total_tokens"])
    ora = slo.OracleIndex(seg, k1=g["k1"], b=g["b"])
    for q in g["queries"]:
        qb = or_queries([q["terms"]], [q["weights"]])
        want = [(e["doc_id"], e["score_bits"]) for e in q["expected"]]
        for mode in ("bm25", "bm25_dense"):
            h, c = ora.search_batch(qb, q["k"], mode)
            assert hits_to_list(h, c) == want, (mode, q["terms"])
        # wand sums in heap-pop order: same ids, scores within 1e-5 relative (north-star rule)
        h, c = ora.search_batch(qb, q["k"], "wand")
        assert [d for d, _ in hits_to_list(h, c)] == [d for d, _ in want] or _near_ties(want)
        ref_scores = np.array([e["score_bits"] for e in q["expected"]], dtype=np.uint32).view(np.float32)
        np.testing.assert_allclose(h[0]["score"][: c[0]], ref_scores, rtol=1e-5)


def _near_ties(want) -> bool:
    s = np.array([b for _, b in want], dtype=np.uint32).view(np.float32)
    return bool(np.any(np.abs(np.diff(s)) <= 1e-5 * np.abs(s[1:])))


# ---- codecs: src/util/varint.rs:5-63, src/index/postings.rs:78-212, :280-310 ---------------------
def test_varint_matches_golden(L):
    for row in golden("codec.json")["varint"]:
        buf = (C.c_uint8 * 8)()
        n = L.slo_varint_write_u32(row["value"], buf)
        assert bytes(buf[:n]).hex() == row["hex"]
        out = C.c_uint32()
        raw = bytes.fromhex(row["hex"])
        arr = (C.c_uint8 * len(raw)).from_buffer_copy(raw)
        assert L.slo_varint_read_u32(arr, len(raw), C.byref(out)) == len(raw)
        assert out.value == row["value"]
    # varint.rs:44-46: a 6th continuation byte is an error
    bad = (C.c_uint8 * 6)(0x80, 0x80, 0x80, 0x80, 0x80, 0x01)
    out = C.c_uint32()
    assert L.slo_varint_read_u32(bad, 6, C.byref(out)) == 0
    trunc = (C.c_uint8 * 2)(0x80, 0x80)
    assert L.slo_varint_read_u32(trunc, 2, C.byref(out)) == 0


def _encode(L, docs, tfs, positions=None):
    d = np.asarray(docs, dtype=np.uint32)
    t = np.asarray(tfs, dtype=np.uint32)
    po = pp = None
    if positions is not None:
        po = np.zeros(len(docs) + 1, dtype=np.uint32)
        po[1:] = np.cumsum([len(p) for p in positions])
        pp = np.asarray([x for p in positions for x in p], dtype=np.uint32)
    n = L.slo_postings_encode(d.ctypes.data, t.ctypes.data, len(d), 1 if positions is not None else 0,
                              0 if po is None else po.ctypes.data, 0 if pp is None else pp.ctypes.data, 0, 0)
    out = np.zeros(n, dtype=np.uint8)
    assert L.slo_postings_encode(d.ctypes.data, t.ctypes.data, len(d), 1 if positions is not None else 0,
                                 0 if po is None else po.ctypes.data, 0 if pp is None else pp.ctypes.data, out.ctypes.data, n) == n
    return out


def _decode(L, img, keep_positions):
    df, bc = C.c_uint32(), C.c_uint32()
    assert L.slo_postings_peek_df(img.ctypes.data, len(img), C.byref(df), C.byref(bc)) == 0
    docs = np.zeros(df.value, dtype=np.uint32)
    tfs = np.zeros(df.value, dtype=np.uint32)
    bmd = np.zeros(max(bc.value, 1), dtype=np.uint32)
    bmt = np.zeros(max(bc.value, 1), dtype=np.float32)
    max_tf, bs, nb, used = C.c_float(), C.c_uint32(), C.c_uint32(), C.c_size_t()
    rc = L.slo_postings_decode(img.ctypes.data, len(img), keep_positions, docs.ctypes.data, tfs.ctypes.data, C.byref(max_tf),
                               C.byref(bs), bmd.ctypes.data, bmt.ctypes.data, C.byref(nb), C.byref(used))
    assert rc == 0
    return docs, tfs, max_tf.value, bs.value, bmd[: nb.value], bmt[: nb.value], used.value


def test_postings_codec_matches_golden_and_round_trips(L):
    g = golden("codec.json")
    lit = g["postings_literal"]
    # index/postings.rs:280-310 writes_and_reads_postings
    img = _encode(L, lit["docs"], lit["tfs"], lit["positions"])
    assert img.tobytes().hex() == lit["hex_with_positions"]
    docs, tfs, max_tf, bs, bmd, bmt, used = _decode(L, img, 1)
    assert docs.tolist() == [1, 2] and tfs.tolist() == [2, 1]
    assert max_tf >= 2.0 and len(bmd) == 1 and len(bmt) == 1 and used == len(img)
    img = _encode(L, lit["docs"], lit["tfs"], None)
    assert img.tobytes().hex() == lit["hex_without_positions"]
    big = g["postings_300"]
    img = _encode(L, big["docs"], big["tfs"], None)
    assert img.tobytes().hex() == big["hex"]
    docs, tfs, max_tf, bs, bmd, bmt, used = _decode(L, img, 0)
    assert docs.tolist() == big["docs"] and tfs.tolist() == big["tfs"]
    assert bs == 128 and len(bmd) == 3 and used == len(img)
    assert bmd.tolist() == [big["docs"][127], big["docs"][255], big["docs"][299]]
    assert bmt.tolist() == [float(max(big["tfs"][i:i + 128])) for i in (0, 128, 256)]
    assert max_tf == float(max(big["tfs"]))


def test_pyref_and_oracle_post_image_agree():
    rng = np.random.default_rng(3)
    doc_tokens = [rng.integers(0, 25, size=int(rng.integers(1, 30))).tolist() for _ in range(400)]
    seg = token_corpus(doc_tokens, 25)
    ora = slo.OracleIndex(seg)
    img, off = ora.build_post_image()
    for t in range(25):
        lo, hi = int(seg.term_offsets[t]), int(seg.term_offsets[t + 1])
        want = pyref.encode_postings(seg.post_docs[lo:hi], seg.post_tfs[lo:hi], None)
        assert img[int(off[t]): int(off[t + 1])].tobytes() == want


# ---- tests/pruning.rs:45-104: bm25 == wand == bmw(4) on a seeded random corpus --------------------
def test_pruning_property_three_strategies_agree():
    rng = np.random.default_rng(42)
    vocab = 7  # ["rust","search","engine","fast","tiny","wand","bmw"]
    doc_tokens = [rng.integers(0, vocab, size=6).tolist() for _ in range(40)]
    seg = token_corpus(doc_tokens, vocab)
    ora = slo.OracleIndex(seg, k1=1.2, b=0.75)
    for _ in range(5):
        terms = rng.permutation(vocab)[:3].tolist()
        qb = or_queries([terms])
        base_h, base_c = ora.search_batch(qb, 6, "bm25")  # limit 5 -> k = 6 (api/reader.rs:2595-2619)
        for mode, bs in (("wand", 0), ("bmw", 4)):
            h, c = ora.search_batch(qb, 6, mode, block_size=bs)
            assert c[0] == base_c[0]
            n = min(5, int(c[0]))
            assert h[0]["doc_id"][:n].tolist() == base_h[0]["doc_id"][:n].tolist()
            assert np.all(np.abs(h[0]["score"][:n] - base_h[0]["score"][:n]) < 1e-5)


def test_empty_and_absent_terms():
    seg = token_corpus([[0, 1], [1, 1, 2]], 3)
    ora = slo.OracleIndex(seg)
    # tests/pruning.rs:107-: an empty query returns no hits; so does a term absent from the segment
    h, c = ora.search_batch(or_queries([[], [0xFFFFFFFF], [2]]), 3, "bm25")
    assert c.tolist() == [0, 0, 1]
    assert int(h[2][0]["doc_id"]) == 1


def test_bmw_counter_example_documents_reference_deviation():
    """the reference's `bmw` stops at the first no-pivot; exact modes must not."""
    n = 4000
    docs = list(range(2000)) + list(range(3000, 3010))
    tfs = [2 if d < 40 else 1 for d in range(2000)] + [9] * 10
    seg = segment_from_postings([(docs, tfs)], [20] * n)
    ora = slo.OracleIndex(seg, k1=0.9, b=0.4)
    qb = or_queries([[0]])
    exact, _ = ora.search_batch(qb, 11, "bm25")
    wand, _ = ora.search_batch(qb, 11, "wand")
    bmw, _ = ora.search_batch(qb, 11, "bmw")
    assert sorted(exact[0]["doc_id"].tolist()) == [0] + list(range(3000, 3010))
    assert wand[0]["doc_id"].tolist() == exact[0]["doc_id"].tolist()
    assert sorted(bmw[0]["doc_id"].tolist()) == list(range(11))  # the faithful restatement of the bug


# ---- matcher + deleted docs: src/api/reader.rs:3009-3036, :1485-1565 -----------------------------
def test_deleted_docs_count_in_df_but_not_in_n_or_hits():
    seg = segment_from_postings([([0, 1, 2, 3], [1, 1, 1, 1])], [5, 5, 5, 5], deleted=[1])
    ora = slo.OracleIndex(seg)
    assert ora.live_docs == 3.0  # index/segment.rs:1365-1370
    h, c = ora.search_batch(or_queries([[0]]), 4, "bm25")
    assert h[0]["doc_id"][: c[0]].tolist() == [0, 2, 3]
    want = pyref.score_tf(1.0, 4.0, 5.0, 5.0, 3.0, 0.9, 0.4, 1.0)  # df counts the deleted doc, N does not
    assert f32_bits(h[0]["score"][0]) == f32_bits(want)


def test_bool_matcher_semantics():
    # tests/query_ast.rs:52-58 style corpus: must / should / must_not / minimum_should_match
    from searchlite_b200.engine import QueryBatch
    docs = [[0, 1], [0, 2], [1, 2], [0, 1, 2], [3]]
    seg = token_corpus(docs, 4)
    ora = slo.OracleIndex(seg)
    qb = QueryBatch.from_bool([
        {"must": [0, 1]},                       # docs with both 0 and 1
        {"must": [0], "must_not": [2]},         # 0 but not 2
        {"should": [0, 1, 2], "min_should": 2}, # at least two of three
        {"should": [3]},
    ])
    h, c = ora.search_batch(qb, 5, "bm25")
    got = [sorted(h[q]["doc_id"][: c[q]].tolist()) for q in range(4)]
    assert got == [[0, 3], [0], [0, 1, 2, 3], [4]]
    for mode in ("wand", "bm25_dense"):
        h2, c2 = ora.search_batch(qb, 5, mode)
        assert [sorted(h2[q]["doc_id"][: c2[q]].tolist()) for q in range(4)] == got


# ---- filters: src/query/filters.rs:84-149, :198-330; src/index/fastfields.rs:475-657 -------------
def test_filter_semantics():
    from searchlite_b200.engine import FILTER_DTYPE, F_AND, F_I64_RANGE, F_KEYWORD_EQ, F_KEYWORD_IN, F_F64_RANGE, F_NOT, F_OR
    seg = token_corpus([[0], [0], [0], [0]], 1)
    seg.fast_str["cat"] = (["News", "sports", "other"], np.array([0, 1, 0xFFFFFFFF, 2], dtype=np.uint32))
    seg.fast_i64["year"] = (np.array([2024, 2019, 2025, 0], dtype=np.int64), np.array([1, 1, 1, 0], dtype=np.uint8))
    seg.fast_f64["score"] = (np.array([0.75, 0.5, 1.5, 0.0]), np.array([1, 1, 1, 0], dtype=np.uint8))
    ora = slo.OracleIndex(seg)
    col = ora.columns

    def node(op, column=-1, i=(0, 0), f=(0.0, 0.0), nc=0, v=(0, 0)):
        n = np.zeros(1, dtype=FILTER_DTYPE)
        n[0] = (op, column, i[0], i[1], f[0], f[1], nc, v[0], v[1])
        return n

    def bits(nodes, strings=()):
        bm = ora.filter_bitmap(np.concatenate(nodes), strings)
        return [d for d in range(4) if (bm[0] >> d) & 1]

    # keyword_filters_are_case_insensitive / evaluates_all_filter_types (filters.rs:198-330)
    assert bits([node(F_KEYWORD_EQ, col["cat"], v=(0, 1))], ["news"]) == [0]
    assert bits([node(F_KEYWORD_IN, col["cat"], v=(0, 2))], ["sports", "NEWS"]) == [0, 1]
    assert bits([node(F_KEYWORD_EQ, col["cat"], v=(0, 1))], ["absent"]) == []
    assert bits([node(F_I64_RANGE, col["year"], i=(2020, 2025))]) == [0, 2]      # inclusive both ends
    assert bits([node(F_I64_RANGE, col["year"], i=(2025, 2030))]) == [2]
    assert bits([node(F_F64_RANGE, col["score"], f=(0.5, 1.0))]) == [0, 1]
    # missing value => predicate false; Not inverts that
    assert bits([node(F_NOT, nc=1), node(F_I64_RANGE, col["year"], i=(0, 3000))]) == [3]
    assert bits([node(F_AND, nc=2), node(F_KEYWORD_EQ, col["cat"], v=(0, 1)), node(F_I64_RANGE, col["year"], i=(2020, 2030))], ["NEWS"]) == [0]
    assert bits([node(F_OR, nc=2), node(F_KEYWORD_EQ, col["cat"], v=(0, 1)), node(F_I64_RANGE, col["year"], i=(2019, 2019))], ["other"]) == [1, 3]
    assert bits([node(F_KEYWORD_EQ, -1, v=(0, 1))], ["news"]) == []                # unknown field


# ---- segment merge: src/api/reader.rs:2777, src/query/sort.rs:80-136 ------------------------------
def test_merge_order_score_then_segment_then_doc():
    a = np.array([(1, 5, 2.0), (1, 9, 1.0)], dtype=slo.HIT_DTYPE)
    b = np.array([(0, 7, 2.0), (0, 8, 2.0), (0, 1, 0.5)], dtype=slo.HIT_DTYPE)
    m = slo.merge_hits([a, b], 4)
    assert [(int(x["segment_ord"]), int(x["doc_id"])) for x in m] == [(0, 7), (0, 8), (1, 5), (1, 9)]


# ---- vectors: src/vectors/mod.rs:74-129, src/api/reader.rs:218-254 --------------------------------
def test_vector_functions_match_golden(L):
    g = golden("vectors.json")
    for r in g["similarity"]:
        a = np.asarray(r["a"], dtype=np.float32)
        b = np.asarray(r["b"], dtype=np.float32)
        got = L.slo_metric_similarity(0 if r["metric"] == "cosine" else 1, a.ctypes.data, b.ctypes.data, len(a))
        assert f32_bits(got) == r["similarity_bits"]
    for r in g["blend"]:
        assert f32_bits(L.slo_blend_scores(r["bm25"], r["vec"], r["alpha"], 1)) == r["bits"]
        assert f32_bits(L.slo_hybrid_score(r["bm25"], 1, r["vec"], r["alpha"], 0)) == r["bits"]
    # missing vector: -1 for cosine, f32::MIN for L2 (api/reader.rs:218-223); alpha >= 1 / <= 0 shortcuts (:241-247)
    assert L.slo_hybrid_score(2.0, 0, 0.0, 0.5, 0) == np.float32(0.5) * np.float32(2.0) + np.float32(0.5) * np.float32(-1.0)
    assert L.slo_hybrid_score(2.0, 0, 0.0, 1.0, 1) == 2.0
    assert f32_bits(L.slo_hybrid_score(2.0, 0, 0.0, 0.0, 1)) == g["missing"]["l2_bits"]
    v = np.array([3.0, 4.0], dtype=np.float32)
    L.slo_normalize_in_place(v.ctypes.data, 2)
    assert v.tolist() == [np.float32(3.0) / np.float32(5.0), np.float32(4.0) / np.float32(5.0)]


# ---- list fast-field columns: "any value" semantics, index/fastfields.rs:490-657 ---------------------------------
def test_list_columns_match_any_value():
    """matches_keyword / matches_keyword_in / matches_i64_range / matches_f64_range over StrList / I64List / F64List
    columns restated directly (any value of the doc; a doc without values fails; Not inverts that) against the
    oracle's filter evaluation"""
    import numpy as np
    from oracle import slo
    from searchlite_b200.engine import FILTER_DTYPE
    from tests.helpers import token_corpus
    rng = np.random.default_rng(17)
    n = 257
    seg = token_corpus([[0]] * n, 1)
    dic = ["Rust", "go", "ZIG", "c"]
    tags = [[dic[j] for j in rng.choice(4, size=int(rng.integers(0, 4)), replace=False)] for _ in range(n)]
    nums = [[int(v) for v in rng.integers(0, 50, size=int(rng.integers(0, 5)))] for _ in range(n)]
    flts = [[float(v) for v in rng.random(int(rng.integers(0, 3)))] for _ in range(n)]

    def offs(ls):
        return np.cumsum([0] + [len(l) for l in ls]).astype(np.uint32)
    seg.fast_str_list["tags"] = (dic, offs(tags), np.array([dic.index(v) for l in tags for v in l], dtype=np.uint32))
    seg.fast_i64_list["nums"] = (offs(nums), np.array([v for l in nums for v in l], dtype=np.int64))
    seg.fast_f64_list["flts"] = (offs(flts), np.array([v for l in flts for v in l], dtype=np.float64))
    ora = slo.OracleIndex(seg)
    c = ora.columns

    def bits(nodes, strings=()):
        bm = ora.filter_bitmap(np.array(nodes, dtype=FILTER_DTYPE), list(strings))
        return [bool((bm[d >> 5] >> (d & 31)) & 1) for d in range(n)]
    F_EQ, F_IN, F_I64, F_F64, F_AND, F_OR, F_NOT = range(7)
    assert bits([(F_EQ, c["tags"], 0, 0, 0, 0, 0, 0, 1)], ["rust"]) == [any(v.lower() == "rust" for v in l) for l in tags]
    assert bits([(F_IN, c["tags"], 0, 0, 0, 0, 0, 0, 2)], ["GO", "zig"]) == [any(v.lower() in ("go", "zig") for v in l) for l in tags]
    assert bits([(F_I64, c["nums"], 10, 20, 0, 0, 0, 0, 0)]) == [any(10 <= v <= 20 for v in l) for l in nums]
    assert bits([(F_F64, c["flts"], 0, 0, 0.25, 0.5, 0, 0, 0)]) == [any(0.25 <= v <= 0.5 for v in l) for l in flts]
    assert bits([(F_NOT, -1, 0, 0, 0, 0, 1, 0, 0), (F_I64, c["nums"], 0, 100, 0, 0, 0, 0, 0)]) == [len(l) == 0 for l in nums]
    # a numeric predicate on a keyword list (wrong column type) is false, as `_ => false`
    assert not any(bits([(F_I64, c["tags"], 0, 100, 0, 0, 0, 0, 0)]))


UNICODE_WORDS = ["Straße", "ÉCOLE", "école", "Ελλάς", "ΕΛΛΆΣ", "ΟΔΟΣ", "οδος", "οδός", "Москва", "МОСКВА", "İstanbul", "i̇stanbul", "ŁÓDŹ", "łódź",
                 "Ÿ", "ÿ", "Ǆ", "ǅ", "ՀԱՅ", "հայ", "ＡＢＣ", "ａｂｃ", "Ḃḃ", "ḃḃ", "plain", "PLAIN", "naïve", "NAÏVE", "Σ", "σ", "ς", "日本語", "ǅungla"]


def test_keyword_compare_lowercases_unicode_like_the_reference():
    """index/fastfields.rs:475-481: non-ASCII values compare by to_lowercase(); checked against Python's str.lower(), which
    implements the same Unicode mapping (final sigma, İ included)"""
    from oracle import slo
    from searchlite_b200.engine import FILTER_DTYPE, F_KEYWORD_EQ
    from tests.helpers import token_corpus
    words = UNICODE_WORDS
    seg = token_corpus([[0]] * len(words), 1)
    seg.fast_str["tag"] = (words, np.arange(len(words), dtype=np.uint32))
    ora = slo.OracleIndex(seg)
    for probe in words:
        n = np.zeros(1, dtype=FILTER_DTYPE)
        n[0] = (F_KEYWORD_EQ, ora.columns["tag"], 0, 0, 0, 0, 0, 0, 1)
        bits = ora.filter_bitmap(n, [probe])
        got = [d for d in range(len(words)) if (bits[d >> 5] >> (d & 31)) & 1]
        if all(ord(ch) < 128 for ch in probe):
            want = [d for d, w in enumerate(words) if (all(ord(ch) < 128 for ch in w) and w.lower() == probe.lower()) or
                    (not all(ord(ch) < 128 for ch in w) and w.lower() == probe.lower())]
        else:
            want = [d for d, w in enumerate(words) if w.lower() == probe.lower()]
        assert got == want, (probe, [words[d] for d in got], [words[d] for d in want])
