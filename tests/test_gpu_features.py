"""GPU parity for everything around the plain OR scan: golden fixtures through the C ABI, edge
cases, deleted docs, the Bool matcher, fast-field filters, the `.post` image load path,
multi-segment merge, weights, wide term frequencies, shard merge and the vector rerank.
Every case is checked against the CPU oracle (or the committed golden bits) on the same inputs."""
import numpy as np
import pytest

from searchlite_b200 import GpuIndex, QueryBatch, SearchliteGpuError, synth
from searchlite_b200.engine import (FILTER_DTYPE, F_AND, F_F64_RANGE, F_I64_RANGE, F_KEYWORD_EQ, F_KEYWORD_IN, F_NOT, F_OR,
                                    HIT_DTYPE)
from tests.helpers import (assert_engine_parity, f32_bits, golden, hits_to_list, or_queries, segment_from_postings,
                           token_corpus)
from tests.parity import assert_parity

pytestmark = pytest.mark.gpu

MATCHER_KERNELS = ["cta", "warp", "warp-inplace"]      # Bool queries: query-order kernels
KERNELS = MATCHER_KERNELS + ["reg", "reg-dense", "reg-nocol", "reg-small"]  # plain OR queries: also the items kernel
# "reg" = the automatic choice for plain OR queries: the posting-driven items kernel (column terms streamed from their columns).
# reg-dense: every term with >= 2 postings and df >= N/64 gets a column (exercises the column path on tiny corpora)
# reg-nocol: the items kernel without columns; reg-small: 128-doc sub-tiles, so tiny corpora span several items
REG_OPTIONS = {"reg-dense": {"dense_min_df": 2, "dense_den": 64},
               "reg-nocol": {"dense_den": 0},
               "reg-small": {"dense_min_df": 2, "dense_den": 16}}
REG_KW = {"reg-small": {"sub_docs": 128}}


def _oracle(seg, **kw):
    from oracle import slo
    return slo.OracleIndex(seg, **kw)


def _gpu(seg, kernel="auto", k1=0.9, b=0.4, **kw):
    if kernel in REG_OPTIONS:
        gi = GpuIndex(0, kernel="reg", options=REG_OPTIONS[kernel], **{**REG_KW.get(kernel, {}), **kw})
    else:
        gi = GpuIndex(0, kernel=kernel, **kw)
    cols = gi.load_segment(seg, k1=k1, b=b)
    return gi, cols


def _check(gi, ora, qb, k, got, kernel, **oracle_kw):
    """query-order kernels: bit-exact vs the oracle; register-tile kernel: bit-exact vs the oracle on its
    declared term order + 1e-5 rule vs the query order"""
    assert_engine_parity(gi, ora, qb, k, got, exact_order=kernel in MATCHER_KERNELS, **oracle_kw)


def node(op, column=-1, i=(0, 0), f=(0.0, 0.0), nc=0, v=(0, 0)):
    n = np.zeros(1, dtype=FILTER_DTYPE)
    n[0] = (op, column, i[0], i[1], f[0], f[1], nc, v[0], v[1])
    return n


# ---- golden fixtures (tests/golden, made by tests/golden/make_golden.py) through the C ABI ----------
@pytest.mark.parametrize("kernel", MATCHER_KERNELS + ["reg"])
def test_reference_literal_known_answer(kernel):
    """query/wand.rs:969-1011 literal postings; expected bits frozen in wand_literal.json"""
    g = golden("wand_literal.json")
    seg = segment_from_postings([(t["docs"], t["tfs"]) for t in g["terms"]], [10] * 10)
    gi, _ = _gpu(seg, kernel, k1=g["k1"], b=g["b"])
    want = [(e["doc_id"], e["score_bits"]) for e in g["expected"]]
    for mode in ("bm25", "wand", "bmw"):
        h, c = gi.search_batch(or_queries([[0, 1]]), g["k"], mode)
        assert hits_to_list(h, c) == want, mode
    gi.close()


@pytest.mark.parametrize("kernel", MATCHER_KERNELS + ["reg"])  # no term of this corpus reaches the default column cutoff
def test_small_corpus_topk_matches_golden_bits(kernel):
    g = golden("small_topk.json")
    seg = segment_from_postings([(p["docs"], p["tfs"]) for p in g["postings"]], g["field_lengths"], g["total_tokens"])
    gi, _ = _gpu(seg, kernel, k1=g["k1"], b=g["b"])
    for q in g["queries"]:
        qb = or_queries([q["terms"]], [q["weights"]])
        want = [(e["doc_id"], e["score_bits"]) for e in q["expected"]]
        for mode in ("bm25", "wand", "bmw"):
            h, c = gi.search_batch(qb, q["k"], mode)
            assert hits_to_list(h, c) == want, (mode, q["terms"])
    gi.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_tie_break_prefers_smaller_doc_id(kernel):
    """query/wand.rs:952-966"""
    seg = segment_from_postings([([1, 2], [1, 1])], [10] * 4)
    gi, _ = _gpu(seg, kernel, k1=1.2, b=0.75)
    h, c = gi.search_batch(or_queries([[0]]), 1, "bm25")
    assert c[0] == 1 and h[0][0]["doc_id"] == 1
    h, c = gi.search_batch(or_queries([[0]]), 2, "wand")
    assert [int(x) for x in h[0]["doc_id"][:2]] == [1, 2] and h[0]["score"][0] == h[0]["score"][1]
    gi.close()


# ---- edge cases ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", KERNELS)
def test_empty_absent_and_ragged_queries(kernel):
    seg = token_corpus([[0, 1], [1, 1, 2], [4], [0, 2, 2, 2, 4]], 6)  # term 3 and 5 have no postings
    ora = _oracle(seg)
    qb = or_queries([[], [0xFFFFFFFF], [2], [3], [5, 3], [0, 1, 2, 3, 4], [1], [999]])
    gi, _ = _gpu(seg, kernel)
    for k in (1, 3, 11):
        for mode in ("bm25", "wand", "bmw"):
            _check(gi, ora, qb, k, gi.search_batch(qb, k, mode), kernel)
    h, c = gi.search_batch(qb, 3, "bm25")
    assert c.tolist()[:2] == [0, 0] and c[3] == 0 and c[4] == 0 and c[7] == 0
    # unused output slots are marked invalid
    assert h[0][0]["doc_id"] == 0xFFFFFFFF and h[0][0]["segment_ord"] == 0xFFFFFFFF
    gi.close()


def test_argument_errors_are_reported_not_fatal():
    seg = token_corpus([[0, 1], [1]], 2)
    gi = GpuIndex(0)
    with pytest.raises(SearchliteGpuError):  # no segment yet
        gi.search_batch(or_queries([[0]]), 3)
    gi.load_segment(seg)
    with pytest.raises(SearchliteGpuError):  # api/reader.rs:2540 bails on limit == 0
        gi.search_batch(or_queries([[0]]), 0)
    with pytest.raises(SearchliteGpuError) as e:
        gi.search_batch(or_queries([[0]]), 1 << 20)
    assert e.value.code == -4
    with pytest.raises(SearchliteGpuError):
        gi.search_batch(or_queries([[0]], [[-1.0]]), 3)
    # the handle is still usable after errors
    h, c = gi.search_batch(or_queries([[1]]), 3)
    assert c[0] == 2
    gi.close()


@pytest.mark.parametrize("k", [1, 2, 33, 257, 2048])
def test_large_k(k):
    """k > 32 on plain OR queries: the flat posting scan with per-query candidate pools (declared summation order); an
    explicit kernel="cta" keeps the reference's order"""
    spec = synth.CorpusSpec(n_docs=9_000, vocab=500, seed=5, len_lo=5, len_hi=40)
    seg = synth.generate_segment(spec, "cpu")
    qb = synth.generate_queries(40, spec.vocab, seed=6, min_rank=2)
    ora = _oracle(seg)
    gi, _ = _gpu(seg)
    for mode in ("bm25", "bmw"):
        got = gi.search_batch(qb, k, mode)
        assert_engine_parity(gi, ora, qb, k, got)
    gi.close()
    gi, _ = _gpu(seg, "cta")
    assert_engine_parity(gi, ora, qb, k, gi.search_batch(qb, k, "bm25"), exact_order=True)
    gi.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_many_terms_per_query(kernel):
    spec = synth.CorpusSpec(n_docs=20_000, vocab=3_000, seed=15, len_lo=20, len_hi=60)
    seg = synth.generate_segment(spec, "cpu")
    n_terms = 40 if kernel == "cta" else 8
    qb = synth.generate_queries(30, spec.vocab, seed=16, min_terms=n_terms, max_terms=n_terms, min_rank=2)
    gi, _ = _gpu(seg, kernel)
    _check(gi, _oracle(seg), qb, 11, gi.search_batch(qb, 11, "bm25"), kernel)
    if kernel != "cta":
        with pytest.raises(SearchliteGpuError):  # an explicit kernel choice that cannot hold the batch is an error
            gi.search_batch(synth.generate_queries(3, spec.vocab, seed=18, min_terms=9, max_terms=9, min_rank=2), 11, "bm25")
    gi.close()
    if kernel == "warp":
        gi, _ = _gpu(seg, "auto")  # more terms than the warp kernel holds: automatic choice must fall to the CTA kernel
        qb2 = synth.generate_queries(10, spec.vocab, seed=17, min_terms=20, max_terms=20, min_rank=2)
        assert_parity(*_oracle(seg).search_batch(qb2, 11, "bm25"), *gi.search_batch(qb2, 11, "bm25"), strict=True)
        gi.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_weights_and_duplicate_keys(kernel):
    """api/reader.rs:2971-2983: duplicate keys arrive merged (weight 2.0); boosts are plain weights"""
    spec = synth.CorpusSpec(n_docs=12_000, vocab=1_000, seed=25, len_lo=10, len_hi=50)
    seg = synth.generate_segment(spec, "cpu")
    rng = np.random.default_rng(3)
    tl = [rng.choice(np.arange(1, 400), size=rng.integers(1, 6), replace=False).tolist() for _ in range(60)]
    w = [[float(np.float32(rng.choice([0.5, 1.0, 2.0, 3.25, 0.1]))) for _ in t] for t in tl]
    qb = or_queries(tl, w)
    ora = _oracle(seg)
    gi, _ = _gpu(seg, kernel)
    for mode in ("bm25", "wand", "bmw"):
        _check(gi, ora, qb, 11, gi.search_batch(qb, 11, mode), kernel)
    gi.close()


@pytest.mark.parametrize("kernel", KERNELS)
def test_wide_term_frequencies(kernel):
    """tf >= 255 does not fit the byte column: the exact side table must be used"""
    n = 3000
    rng = np.random.default_rng(9)
    d0 = np.sort(rng.choice(n, size=1500, replace=False))
    tf0 = rng.integers(1, 6, size=1500)
    tf0[::37] = rng.integers(255, 70000, size=len(tf0[::37]))
    tf0[5] = 255
    tf0[6] = 254
    d1 = np.sort(rng.choice(n, size=700, replace=False))
    tf1 = rng.integers(1, 4, size=700)
    lens = rng.integers(50, 400, size=n)
    seg = segment_from_postings([(d0.tolist(), tf0.tolist()), (d1.tolist(), tf1.tolist())], lens.tolist())
    ora = _oracle(seg)
    qb = or_queries([[0], [0, 1], [1, 0]])
    gi, _ = _gpu(seg, kernel)
    for mode in ("bm25", "bmw"):
        _check(gi, ora, qb, 11, gi.search_batch(qb, 11, mode), kernel)
    gi.close()


def test_zero_and_missing_field_lengths():
    """query/wand.rs:77-84: a non-positive length scores as max(avgdl, 1)"""
    lens = [0, 12, 7, 0, 30, 9]
    present = np.array([1, 1, 1, 0, 1, 1], dtype=np.uint8)
    seg = segment_from_postings([([0, 1, 3, 4], [1, 2, 1, 3]), ([2, 3, 5], [1, 1, 4])], lens, total_tokens=58)
    seg.field_length_present = present
    ora = _oracle(seg)
    gi, _ = _gpu(seg)
    st = gi.segment_stats(0)
    assert st["min_doc_len"] == ora.min_doc_len == 7.0 and st["avgdl"] == ora.avgdl
    assert_parity(*ora.search_batch(or_queries([[0, 1], [1], [0]]), 6, "bm25"), *gi.search_batch(or_queries([[0, 1], [1], [0]]), 6, "bm25"),
                  strict=True)
    gi.close()


# ---- deleted docs + matcher: api/reader.rs:3009-3036, :1485-1565 --------------------------------------
@pytest.mark.parametrize("kernel", KERNELS)
def test_deleted_docs(kernel):
    spec = synth.CorpusSpec(n_docs=10_000, vocab=900, seed=31, len_lo=10, len_hi=40)
    seg = synth.generate_segment(spec, "cpu")
    seg.deleted_docs = np.unique(np.random.default_rng(1).integers(0, spec.n_docs, size=2500)).astype(np.uint32)
    qb = synth.generate_queries(80, spec.vocab, seed=32, min_rank=2)
    ora = _oracle(seg)
    ref = ora.search_batch(qb, 11, "bm25")
    assert not np.isin(ref[0]["doc_id"][ref[0]["segment_ord"] == 0], seg.deleted_docs).any()
    gi, _ = _gpu(seg, kernel)
    assert gi.segment_stats(0)["live_docs"] == ora.live_docs == float(spec.n_docs - len(seg.deleted_docs))
    for mode in ("bm25", "wand", "bmw"):
        _check(gi, ora, qb, 11, gi.search_batch(qb, 11, mode), kernel)
    gi.close()


@pytest.mark.parametrize("kernel", MATCHER_KERNELS + ["auto"])
def test_bool_matcher_literal(kernel):
    """tests/query_ast.rs:52-58 style corpus"""
    seg = token_corpus([[0, 1], [0, 2], [1, 2], [0, 1, 2], [3]], 4)
    qb = QueryBatch.from_bool([
        {"must": [0, 1]}, {"must": [0], "must_not": [2]}, {"should": [0, 1, 2], "min_should": 2}, {"should": [3]},
        {"must": [0], "should": [1]}, {"must_not": [0], "should": [1, 2, 3]},
    ])
    ref = _oracle(seg).search_batch(qb, 5, "bm25")
    gi, _ = _gpu(seg, kernel)
    got = gi.search_batch(qb, 5, "bm25")
    assert_parity(*ref, *got, strict=True)
    assert [sorted(got[0][q]["doc_id"][: got[1][q]].tolist()) for q in range(4)] == [[0, 3], [0], [0, 1, 2, 3], [4]]
    assert_parity(*ref, *gi.search_batch(qb, 5, "bmw"), strict=True)
    gi.close()


@pytest.mark.parametrize("kernel", MATCHER_KERNELS + ["auto"])
def test_bool_matcher_random(kernel):
    """C4 shape: Bool{must:[t1,t2(,t3)]} executed as OR-scan + reject (api/reader.rs:1527-1563)"""
    spec = synth.CorpusSpec(n_docs=30_000, vocab=2_000, seed=41, len_lo=20, len_hi=80)
    seg = synth.generate_segment(spec, "cpu")
    rng = np.random.default_rng(4)
    qs = []
    for i in range(90):
        t = rng.choice(np.arange(1, 120), size=5, replace=False).tolist()
        kind = i % 5
        if kind == 0:
            qs.append({"must": t[:2]})
        elif kind == 1:
            qs.append({"must": t[:3]})
        elif kind == 2:
            qs.append({"must": t[:1], "must_not": t[1:2], "should": t[2:4]})
        elif kind == 3:
            qs.append({"should": t[:4], "min_should": 2})
        else:
            qs.append({"should": t[:3], "must_not": t[3:5]})
    qb = QueryBatch.from_bool(qs)
    ref = _oracle(seg).search_batch(qb, 11, "bm25")
    assert ref[1].sum() > 0
    gi, _ = _gpu(seg, kernel)
    for mode in ("bm25", "wand", "bmw"):
        assert_parity(*ref, *gi.search_batch(qb, 11, mode), strict=True)
    gi.close()


# ---- filters: query/filters.rs:84-149, index/fastfields.rs:475-657 --------------------------------------
def test_filter_bitmaps_match_oracle_literal():
    seg = token_corpus([[0], [0], [0], [0]], 1)
    seg.fast_str["cat"] = (["News", "sports", "other"], np.array([0, 1, 0xFFFFFFFF, 2], dtype=np.uint32))
    seg.fast_i64["year"] = (np.array([2024, 2019, 2025, 0], dtype=np.int64), np.array([1, 1, 1, 0], dtype=np.uint8))
    seg.fast_f64["score"] = (np.array([0.75, 0.5, 1.5, 0.0]), np.array([1, 1, 1, 0], dtype=np.uint8))
    ora = _oracle(seg)
    gi, col = _gpu(seg)
    ocol = ora.columns
    cases = [
        ([(F_KEYWORD_EQ, "cat", dict(v=(0, 1)))], ["news"], [0]),
        ([(F_KEYWORD_IN, "cat", dict(v=(0, 2)))], ["sports", "NEWS"], [0, 1]),
        ([(F_KEYWORD_EQ, "cat", dict(v=(0, 1)))], ["absent"], []),
        ([(F_I64_RANGE, "year", dict(i=(2020, 2025)))], [], [0, 2]),
        ([(F_I64_RANGE, "year", dict(i=(2025, 2030)))], [], [2]),
        ([(F_F64_RANGE, "score", dict(f=(0.5, 1.0)))], [], [0, 1]),
        ([(F_NOT, None, dict(nc=1)), (F_I64_RANGE, "year", dict(i=(0, 3000)))], [], [3]),
        ([(F_AND, None, dict(nc=2)), (F_KEYWORD_EQ, "cat", dict(v=(0, 1))), (F_I64_RANGE, "year", dict(i=(2020, 2030)))], ["NEWS"], [0]),
        ([(F_OR, None, dict(nc=2)), (F_KEYWORD_EQ, "cat", dict(v=(0, 1))), (F_I64_RANGE, "year", dict(i=(2019, 2019)))], ["other"], [1, 3]),
        ([(F_KEYWORD_EQ, "nope", dict(v=(0, 1)))], ["news"], []),
        ([(F_I64_RANGE, "cat", dict(i=(0, 10)))], [], []),  # wrong column type: predicate false
    ]
    for spec_nodes, strings, want in cases:
        gnodes = np.concatenate([node(op, col.get(c, -1) if c else -1, **kw) for op, c, kw in spec_nodes])
        onodes = np.concatenate([node(op, ocol.get(c, -1) if c else -1, **kw) for op, c, kw in spec_nodes])
        fid = gi.compile_filter(gnodes, strings)
        bm = gi.filter_bitmap(fid, 0, 4)
        assert [d for d in range(4) if (bm[0] >> d) & 1] == want, spec_nodes
        assert np.array_equal(bm, ora.filter_bitmap(onodes, strings))
    gi.close()


def test_list_column_filters_match_oracle():
    """StrList / I64List / F64List columns: a predicate holds when ANY value of the doc satisfies it
    (index/fastfields.rs:497-509, 548-562, 602-609, 632-639); bitmaps equal to the oracle's, also under Not / And / Or
    and mixed with scalar columns, and a filtered search on top"""
    rng = np.random.default_rng(23)
    spec = synth.CorpusSpec(n_docs=5_003, vocab=400, seed=24, len_lo=10, len_hi=40)
    seg = synth.generate_segment(spec, "cpu")
    n = spec.n_docs
    dic = ["Rust", "go", "ZIG", "c", "ada"]
    tags = [[dic[j] for j in rng.choice(5, size=int(rng.integers(0, 4)), replace=False)] for _ in range(n)]
    nums = [[int(v) for v in rng.integers(0, 60, size=int(rng.integers(0, 6)))] for _ in range(n)]
    flts = [[float(v) for v in rng.random(int(rng.integers(0, 3)))] for _ in range(n)]

    def offs(ls):
        return np.cumsum([0] + [len(l) for l in ls]).astype(np.uint32)
    seg.fast_str_list["tags"] = (dic, offs(tags), np.array([dic.index(v) for l in tags for v in l], dtype=np.uint32))
    seg.fast_i64_list["nums"] = (offs(nums), np.array([v for l in nums for v in l], dtype=np.int64))
    seg.fast_f64_list["flts"] = (offs(flts), np.array([v for l in flts for v in l], dtype=np.float64))
    seg.fast_i64["year"] = (rng.integers(2000, 2026, size=n).astype(np.int64), (rng.random(n) > 0.1).astype(np.uint8))
    ora = _oracle(seg)
    gi, col = _gpu(seg)
    ocol = ora.columns
    cases = [
        ([(F_KEYWORD_EQ, "tags", dict(v=(0, 1)))], ["rust"]),
        ([(F_KEYWORD_IN, "tags", dict(v=(0, 2)))], ["GO", "zig"]),
        ([(F_KEYWORD_EQ, "tags", dict(v=(0, 1)))], ["absent"]),
        ([(F_I64_RANGE, "nums", dict(i=(10, 12)))], []),
        ([(F_F64_RANGE, "flts", dict(f=(0.25, 0.3)))], []),
        ([(F_NOT, None, dict(nc=1)), (F_I64_RANGE, "nums", dict(i=(0, 100)))], []),
        ([(F_AND, None, dict(nc=3)), (F_KEYWORD_EQ, "tags", dict(v=(0, 1))), (F_I64_RANGE, "nums", dict(i=(0, 30))),
          (F_I64_RANGE, "year", dict(i=(2005, 2020)))], ["C"]),
        ([(F_OR, None, dict(nc=2)), (F_F64_RANGE, "flts", dict(f=(0.9, 1.0))), (F_KEYWORD_EQ, "tags", dict(v=(0, 1)))], ["ada"]),
        ([(F_I64_RANGE, "tags", dict(i=(0, 10)))], []),   # wrong column type: predicate false
        ([(F_KEYWORD_EQ, "nums", dict(v=(0, 1)))], ["rust"]),
    ]
    fids = []
    for spec_nodes, strings in cases:
        gnodes = np.concatenate([node(op, col.get(c, -1) if c else -1, **kw) for op, c, kw in spec_nodes])
        onodes = np.concatenate([node(op, ocol.get(c, -1) if c else -1, **kw) for op, c, kw in spec_nodes])
        fid = gi.compile_filter(gnodes, strings)
        fids.append((fid, onodes, strings))
        assert np.array_equal(gi.filter_bitmap(fid, 0, n), ora.filter_bitmap(onodes, strings)), spec_nodes
    want0 = [any(v.lower() == "rust" for v in l) for l in tags]
    bm = gi.filter_bitmap(fids[0][0], 0, n)
    assert [bool((bm[d >> 5] >> (d & 31)) & 1) for d in range(n)] == want0 and 0 < sum(want0) < n
    # filtered search on the list predicates
    qb = synth.generate_queries(40, spec.vocab, seed=25, min_rank=2)
    for fid, onodes, strings in (fids[0], fids[6]):
        qb.filter_id = np.full(qb.n_queries, fid, dtype=np.int32)
        qb._structs = None  # (the struct array is cached per batch)
        ref = ora.search_batch(qb, 11, "bm25", filter_nodes=onodes, strings=strings)
        assert_parity(*ref, *gi.search_batch(qb, 11, "bm25"), strict=False)
    gi.close()


@pytest.mark.parametrize("kernel", MATCHER_KERNELS + ["auto"])
def test_filtered_search_c4_shape(kernel):
    """C4: Bool{must} + root filter And[KeywordEq(lang), I64Range(year)] at three selectivities"""
    spec = synth.CorpusSpec(n_docs=40_000, vocab=3_000, seed=51, len_lo=20, len_hi=80)
    seg = synth.generate_segment(spec, "cpu")
    names, lang, year = synth.fast_fields(spec)
    seg.fast_str["lang"] = (names, lang)
    seg.fast_i64["year"] = (year, None)
    ora = _oracle(seg)
    gi, col = _gpu(seg, kernel)
    rng = np.random.default_rng(6)
    filters = [("en", 2000, 2025), ("es", 2010, 2020), ("sv", 2024, 2024)]
    for lang_name, lo, hi in filters:
        def prog(c):
            return np.concatenate([node(F_AND, nc=2), node(F_KEYWORD_EQ, c["lang"], v=(0, 1)), node(F_I64_RANGE, c["year"], i=(lo, hi))])
        fid = gi.compile_filter(prog(col), [lang_name.upper()])
        want_bits = ora.filter_bitmap(prog(ora.columns), [lang_name.upper()])
        assert np.array_equal(gi.filter_bitmap(fid, 0, spec.n_docs), want_bits)
        qs = []
        for i in range(40):
            t = rng.choice(np.arange(1, 150), size=3, replace=False).tolist()
            qs.append({"must": t[:2], "filter_id": fid} if i % 2 else {"should": t, "min_should": 1, "filter_id": fid})
        qb = QueryBatch.from_bool(qs)
        ref = ora.search_batch(qb, 11, "bm25", filter_nodes=prog(ora.columns), strings=[lang_name.upper()])
        for mode in ("bm25", "bmw"):
            assert_parity(*ref, *gi.search_batch(qb, 11, mode), strict=True)
    # plain OR queries with a filter and without, mixed in one batch
    qb = synth.generate_queries(30, spec.vocab, seed=52, min_rank=2)
    qb.filter_id = np.array([0 if i % 2 else -1 for i in range(30)], dtype=np.int32)
    got = gi.search_batch(qb, 11, "bm25")
    plain = synth.generate_queries(30, spec.vocab, seed=52, min_rank=2)
    ref_nofilter = ora.search_batch(plain, 11, "bm25")
    def prog0(c):
        return np.concatenate([node(F_AND, nc=2), node(F_KEYWORD_EQ, c["lang"], v=(0, 1)), node(F_I64_RANGE, c["year"], i=(2000, 2025))])
    ref_filter = ora.search_batch(plain, 11, "bm25", filter_nodes=prog0(ora.columns), strings=["EN"])
    for q in range(30):
        src = ref_filter if q % 2 else ref_nofilter
        # "auto" runs plain OR queries on the register-tile kernel (its own term order): 1e-5 rule
        assert_parity(src[0][q:q + 1], src[1][q:q + 1], got[0][q:q + 1], got[1][q:q + 1], strict=kernel != "auto")
    gi.close()


# ---- residency from the reference's `.post` byte image: index/postings.rs:78-212 --------------------------
def test_post_image_load_equals_csr_load():
    spec = synth.CorpusSpec(n_docs=25_000, vocab=2_500, seed=61, len_lo=20, len_hi=90)
    seg = synth.generate_segment(spec, "cpu")
    ora = _oracle(seg)
    img, off = ora.build_post_image()
    qb = synth.generate_queries(120, spec.vocab, seed=62, min_rank=2)
    gi = GpuIndex(0)
    gi.load_segment_post_image(seg, img, off)
    assert gi.segment_stats(0)["n_postings"] == len(seg.post_docs)
    for mode in ("bm25", "bmw"):
        assert_engine_parity(gi, ora, qb, 11, gi.search_batch(qb, 11, mode))
    gi.close()
    gi = GpuIndex(0, kernel="items", sub_docs=512, options={"dense_min_df": 64})
    gi.load_segment_post_image(seg, img, off)
    for mode in ("bm25", "bmw"):
        assert_engine_parity(gi, ora, qb, 11, gi.search_batch(qb, 11, mode))
    gi.close()
    gi = GpuIndex(0, kernel="warp")
    gi.load_segment_post_image(seg, img, off)
    assert_parity(*ora.search_batch(qb, 11, "bm25"), *gi.search_batch(qb, 11, "bm25"), strict=True)
    gi.close()
    # a truncated image is an error, not a crash
    gi = GpuIndex(0)
    with pytest.raises(SearchliteGpuError):
        gi.load_segment_post_image(seg, img[: len(img) // 2], off)
    gi.close()


@pytest.mark.parametrize("with_positions", [False, True])
@pytest.mark.parametrize("kernel", ["auto", "warp", "cta"])
def test_wide_term_frequencies_through_the_post_image(kernel, with_positions):
    """tf >= 255 in a reference-format posting image (long documents): the decode keeps the exact values in a side list and
    the load builds the same wide-tf table as the CSR path (such indexes used to be refused)"""
    n = 4000
    rng = np.random.default_rng(19)
    d0 = np.sort(rng.choice(n, size=2500, replace=False))
    tf0 = rng.integers(1, 6, size=2500)
    tf0[::41] = rng.integers(255, 3000, size=len(tf0[::41]))
    tf0[5], tf0[6], tf0[-1] = 255, 254, 900
    d1 = np.sort(rng.choice(n, size=900, replace=False))
    tf1 = rng.integers(1, 4, size=900)
    d2 = np.sort(rng.choice(n, size=300, replace=False))
    tf2 = np.full(300, 300)
    lens = rng.integers(50, 4000, size=n)
    seg = segment_from_postings([(d0.tolist(), tf0.tolist()), (d1.tolist(), tf1.tolist()), (d2.tolist(), tf2.tolist())], lens.tolist())
    ora = _oracle(seg)
    if with_positions:  # position lists of tf entries (capped: the image only needs SOME positions per posting to change its layout)
        npos = np.minimum(seg.post_tfs.astype(np.int64), 7)
        pos_off = np.zeros(len(npos) + 1, dtype=np.uint64)
        pos_off[1:] = np.cumsum(npos)
        positions = np.concatenate([np.arange(k, dtype=np.uint32) * 3 for k in npos]) if len(npos) else np.zeros(0, np.uint32)
        ora.set_positions(pos_off, positions)
    img, off = ora.build_post_image()
    qb = or_queries([[0], [0, 1], [1, 0], [2, 0], [2]])
    gi = GpuIndex(0, kernel=kernel)
    gi.load_segment_post_image(seg, img, off)
    csr, _ = _gpu(seg, kernel)
    for mode in ("bm25", "bmw"):
        got = gi.search_batch(qb, 11, mode)
        _check(gi, ora, qb, 11, got, kernel)
        ref = csr.search_batch(qb, 11, mode)
        assert got[0].tobytes() == ref[0].tobytes() and got[1].tobytes() == ref[1].tobytes()
    gi.close()
    csr.close()


def test_corrupt_postings_are_refused_at_load():
    """doc ids outside the segment or lists that do not ascend must not reach the scoring kernels"""
    good = segment_from_postings([([0, 3, 7], [1, 2, 1]), ([2, 3], [1, 1])], [5, 5, 5, 5, 5, 5, 5, 5])
    gi = GpuIndex(0)
    gi.load_segment(good)
    gi.close()
    for docs in ([0, 3, 8], [0, 7, 3], [3, 3, 7]):
        bad = segment_from_postings([(docs, [1, 2, 1]), ([2, 3], [1, 1])], [5, 5, 5, 5, 5, 5, 5, 5])
        gi = GpuIndex(0)
        with pytest.raises(SearchliteGpuError, match="outside the segment|ascending"):
            gi.load_segment(bad)
        gi.close()
        # the same through a posting image
        from oracle import slo
        o = slo.OracleIndex(bad)
        img, off = o.build_post_image()
        gi = GpuIndex(0)
        with pytest.raises(SearchliteGpuError, match="outside the segment|ascending"):
            gi.load_segment_post_image(bad, img, off)
        gi.close()


# ---- several segments in one handle: api/reader.rs:2670-2777 ----------------------------------------------
@pytest.mark.parametrize("kernel", ["cta", "warp", "reg", "reg-dense", "reg-nocol", "reg-small"])
def test_multi_segment_merge_order(kernel):
    from oracle import slo
    from searchlite_b200.shard import shard_ranges
    from tests.helpers import canonical_batch
    n_docs, vocab, world = 24_000, 1_500, 3
    qb = synth.generate_queries(70, vocab, seed=72, min_rank=2)
    gi = GpuIndex(0, kernel="reg", options=REG_OPTIONS[kernel], **REG_KW.get(kernel, {})) if kernel in REG_OPTIONS else GpuIndex(0, kernel=kernel)
    per_seg = []
    for r, (lo, hi) in enumerate(shard_ranges(n_docs, world)):
        spec = synth.CorpusSpec(n_docs=hi - lo, vocab=vocab, seed=71, len_lo=10, len_hi=60, segment_ord=r, doc_base=lo)
        seg = synth.generate_segment(spec, "cpu")
        gi.load_segment(seg)
        # the register-tile kernel's term order depends on which terms have a column in THIS segment
        per_seg.append(slo.OracleIndex(seg).search_batch(canonical_batch(gi, qb, r) if kernel.startswith("reg") else qb, 11, "bm25"))
    got_h, got_c = gi.search_batch(qb, 11, "bm25")
    for q in range(qb.n_queries):
        want = slo.merge_hits([h[q, : c[q]] for h, c in per_seg], 11)
        assert got_c[q] == len(want)
        g = got_h[q, : got_c[q]]
        assert np.array_equal(g["segment_ord"], want["segment_ord"]) and np.array_equal(g["doc_id"], want["doc_id"])
        assert np.array_equal(g["score"].view(np.uint32), want["score"].view(np.uint32))
    gi.close()


def test_merge_gathered_matches_sortkey_order():
    """api/reader.rs:2777 + query/sort.rs:80-93: score desc, segment_ord asc, doc_id asc"""
    import torch
    from oracle import slo
    rng = np.random.default_rng(8)
    n_shards, nq, k = 4, 50, 11
    hits = np.zeros((n_shards, nq, k), dtype=HIT_DTYPE)
    counts = rng.integers(0, k + 1, size=(n_shards, nq)).astype(np.uint32)
    for s in range(n_shards):
        for q in range(nq):
            c = counts[s, q]
            sc = np.sort(rng.choice(np.array([0.5, 1.0, 1.5, 2.0, 2.5, 3.0], dtype=np.float32), size=c))[::-1]
            docs = rng.choice(1000, size=c, replace=False)
            order = np.lexsort((docs, -sc))
            hits[s, q, :c]["segment_ord"] = s
            hits[s, q, :c]["doc_id"] = docs[order]
            hits[s, q, :c]["score"] = sc[order]
            hits[s, q, c:]["segment_ord"] = 0xFFFFFFFF
            hits[s, q, c:]["doc_id"] = 0xFFFFFFFF
    gi = GpuIndex(0)
    dh = torch.from_numpy(hits.view(np.uint8).reshape(-1).copy()).cuda()
    dc = torch.from_numpy(counts.view(np.int32).reshape(-1).copy()).cuda()
    torch.cuda.synchronize()
    mh, mc = gi.merge_gathered(dh.data_ptr(), dc.data_ptr(), n_shards, nq, k)
    for q in range(nq):
        want = slo.merge_hits([hits[s, q, : counts[s, q]] for s in range(n_shards)], k)
        assert mc[q] == len(want)
        assert mh[q, : mc[q]].tobytes() == want.tobytes()
    gi.close()


# ---- vectors + rerank: vectors/mod.rs:63-129, api/reader.rs:218-254 -------------------------------------
@pytest.mark.parametrize("metric", ["cosine", "l2"])
@pytest.mark.parametrize("bf16", [False, True])
def test_rerank_matches_oracle_formulae(metric, bf16):
    from oracle import slo
    L = slo.lib()
    spec = synth.CorpusSpec(n_docs=5_000, vocab=600, seed=81, len_lo=10, len_hi=40)
    seg = synth.generate_segment(spec, "cpu")
    rng = np.random.default_rng(10)
    dim, nq, k = 64, 25, 50
    have = rng.random(spec.n_docs) < 0.9
    offsets = np.full(spec.n_docs, 0xFFFFFFFF, dtype=np.uint32)
    offsets[have] = np.arange(int(have.sum()), dtype=np.uint32)
    vecs = rng.standard_normal((int(have.sum()), dim)).astype(np.float32)
    qv = rng.standard_normal((nq, dim)).astype(np.float32)
    if metric == "cosine":  # normalize_in_place, vectors/mod.rs:74-81
        for row in vecs:
            L.slo_normalize_in_place(row.ctypes.data, dim)
        for row in qv:
            L.slo_normalize_in_place(row.ctypes.data, dim)
    gi, _ = _gpu(seg)
    gi.load_vectors(0, offsets, vecs, store_bf16=bf16)
    qb = synth.generate_queries(nq, spec.vocab, seed=82, min_rank=2)
    cands, cc = gi.search_batch(qb, k, "bm25")
    alpha = 0.5
    out, vs = gi.rerank(qv, cands, cc, alpha, metric)
    m = 0 if metric == "cosine" else 1
    # bf16 storage is this build's choice: rows are rounded once, the arithmetic stays f32
    # f32 rows: the kernel folds in dimension order like the reference, so the score is the oracle's to the bit
    tol = 2e-2 if bf16 else 0.0
    for q in range(nq):
        n = int(cc[q])
        exp = {}
        for h in cands[q, :n]:
            d = int(h["doc_id"])
            if offsets[d] == 0xFFFFFFFF:
                s = L.slo_hybrid_score(float(h["score"]), 0, 0.0, alpha, m)
            else:
                v = np.ascontiguousarray(vecs[offsets[d]])
                sim = L.slo_metric_similarity(m, qv[q].ctypes.data, v.ctypes.data, dim)
                s = L.slo_hybrid_score(float(h["score"]), 1, sim, alpha, m)
            exp[d] = s
        got = out[q, :n]
        assert sorted(got["doc_id"].tolist()) == sorted(exp.keys())
        for h in got:
            e = exp[int(h["doc_id"])]
            assert abs(float(h["score"]) - e) <= tol * max(1.0, abs(e)), (q, int(h["doc_id"]), float(h["score"]), e)
        sc = got["score"]
        assert np.all(sc[:-1] >= sc[1:])  # re-ordered by blended score
        missing = [int(h["doc_id"]) for h in got if offsets[int(h["doc_id"])] == 0xFFFFFFFF]
        if missing and metric == "cosine":
            for h, v in zip(got, vs[q, :n]):
                if int(h["doc_id"]) in missing:
                    assert v == 0.0  # RankedHit.vector_score is None without a vector (api/reader.rs:253)
    # alpha shortcuts (api/reader.rs:241-247)
    out1, _ = gi.rerank(qv, cands, cc, 1.0, metric)
    for q in range(nq):
        assert out1[q, : cc[q]].tobytes() == cands[q, : cc[q]].tobytes()
    gi.close()


# ---- statistics (QueryStats, query/wand.rs:45-50) -----------------------------------------------------------
@pytest.mark.parametrize("kernel", ["cta", "warp", "warp-inplace", "reg", "reg-dense", "reg-nocol", "reg-small"])
def test_stats_count_scored_docs_and_postings(kernel):
    spec = synth.CorpusSpec(n_docs=15_000, vocab=1_200, seed=91, len_lo=10, len_hi=50)
    seg = synth.generate_segment(spec, "cpu")
    qb = synth.generate_queries(50, spec.vocab, seed=92, min_rank=2)
    gi, _ = _gpu(seg, kernel)
    _, _, st = gi.search_batch(qb, 11, "bm25", want_stats=True)
    off = seg.term_offsets.astype(np.int64)
    for q in range(qb.n_queries):
        terms = qb.terms["term_id"][qb.term_off[q]: qb.term_off[q + 1]]
        n_post = int(sum(off[t + 1] - off[t] for t in terms))
        n_docs = len(np.unique(np.concatenate([seg.post_docs[off[t]: off[t + 1]] for t in terms])))
        assert st["postings_advanced"][q] == n_post
        assert st["scored_docs"][q] == n_docs
    gi.close()


# ---- sharded runs: the per-query threshold exchange between the two parts of the posting scan ----------
@pytest.mark.parametrize("k", [11, 101])
@pytest.mark.parametrize("execution", ["bm25", "bmw"])
def test_two_step_scan_with_threshold_exchange(execution, k):
    """two shards (two handles on one device): first part of the scan, max of the shards' k-th keys imported into both, rest of
    the scan.  Every shard then prunes against the better bound; the merged result must stay the exact top k"""
    import torch
    from oracle import slo
    from searchlite_b200.shard import shard_ranges
    n_docs, vocab = 60_000, 4_000
    qb = synth.generate_queries(96, vocab, seed=232, min_rank=2)
    gis, oras, preps = [], [], []
    for r, (lo, hi) in enumerate(shard_ranges(n_docs, 2)):
        spec = synth.CorpusSpec(n_docs=hi - lo, vocab=vocab, seed=231, len_lo=20, len_hi=90, segment_ord=r, doc_base=lo)
        seg = synth.generate_segment(spec, "cpu")
        gi = GpuIndex(0, options={"scan_first_part": 64})
        gi.load_segment(seg)
        gis.append(gi)
        oras.append(slo.OracleIndex(seg))
    one_step = []
    for gi in gis:
        p = gi.prepare(qb, k, execution)
        p.run(sync=True)
        one_step.append(p.fetch())
        preps.append(p)
    # two-step: seeds on both, exchange, sweep on both
    for p in preps:
        assert p.run_seeds()
    torch.cuda.synchronize()
    from searchlite_b200.shard import ShardedSearcher
    keys = [ShardedSearcher(gi, qb.n_queries, k)._as_tensor(p.threshold_keys_ptr(), qb.n_queries * 8, torch.int64).clone() for gi, p in zip(gis, preps)]
    glob = torch.maximum(keys[0], keys[1])
    torch.cuda.synchronize()
    raised = 0
    for gi, p, own in zip(gis, preps, keys):
        raised += int(((glob >> 32) > (own >> 32)).sum())
        p.import_thresholds(glob.data_ptr())
        p.run_sweep(sync=True)
    if k <= 32:
        assert raised > 0  # the exchange did hand some shard a better bound (k > 32: a pool publishes its bound only once it overflows)
    two_step = [p.fetch() for p in preps]
    from tests.helpers import canonical_batch
    for qi in range(qb.n_queries):
        sub = qb.subset(qi, qi + 1)
        ref = slo.merge_hits([o.search_batch(canonical_batch(g, sub, r), k, "bm25")[0][0][: o.search_batch(canonical_batch(g, sub, r), k, "bm25")[1][0]]
                              for r, (o, g) in enumerate(zip(oras, gis))], k)
        got1 = slo.merge_hits([h[qi, : c[qi]] for h, c in one_step], k)
        got2 = slo.merge_hits([h[qi, : c[qi]] for h, c in two_step], k)
        assert got1.tobytes() == ref.tobytes(), qi
        assert got2.tobytes() == ref.tobytes(), qi
    for p in preps:
        p.free()
    for gi in gis:
        gi.close()


@pytest.mark.parametrize("k", [11, 101])
@pytest.mark.parametrize("execution", ["bm25", "bmw"])
def test_threshold_board_between_two_shards(execution, k):
    """slg_batch_set_threshold_board: two shards as two handles on one device, each other's board as the peer mapping.  Shard
    0 runs first and pushes its k-th scores into shard 1's board; shard 1 then prunes against them.  The merged result stays
    the exact top k, stale epochs are ignored, and shard 1 verifies fewer postings than it does alone"""
    import torch
    from oracle import slo
    from searchlite_b200.shard import shard_ranges
    from tests.helpers import canonical_batch
    n_docs, vocab = 80_000, 4_000
    qb = synth.generate_queries(128, vocab, seed=242, min_rank=2)
    gis, oras = [], []
    for r, (lo, hi) in enumerate(shard_ranges(n_docs, 2)):
        spec = synth.CorpusSpec(n_docs=hi - lo, vocab=vocab, seed=241, len_lo=20, len_hi=90, segment_ord=r, doc_base=lo)
        seg = synth.generate_segment(spec, "cpu")
        gi = GpuIndex(0)
        gi.load_segment(seg)
        gis.append(gi)
        oras.append(slo.OracleIndex(seg))
    boards = [torch.zeros(qb.n_queries, dtype=torch.int64, device="cuda") for _ in range(2)]
    torch.cuda.synchronize()
    preps = [gi.prepare(qb, k, execution) for gi in gis]
    alone = []
    for p in preps:
        p.run(sync=True)
        alone.append(p.fetch())
    verified_alone = gis[1].counters()["last_postings_verified"]
    for epoch in (1, 2):
        for r, p in enumerate(preps):
            p.set_threshold_board(boards[r].data_ptr(), [boards[1 - r].data_ptr()], epoch)
            p.run(sync=True)
        got = [p.fetch() for p in preps]
        verified_board = gis[1].counters()["last_postings_verified"]
        for qi in range(qb.n_queries):
            sub = qb.subset(qi, qi + 1)
            lists = []
            for r, (o, g) in enumerate(zip(oras, gis)):
                h, c = o.search_batch(canonical_batch(g, sub, r), k, "bm25")
                lists.append(h[0, : c[0]])
            ref = slo.merge_hits(lists, k)
            assert slo.merge_hits([h[qi, : c[qi]] for h, c in got], k).tobytes() == ref.tobytes(), (epoch, qi)
            assert slo.merge_hits([h[qi, : c[qi]] for h, c in alone], k).tobytes() == ref.tobytes(), qi
        assert int((boards[1] >> 32).max()) == epoch  # shard 0 pushed into shard 1's board under this epoch
        if k <= 32:
            assert verified_board < verified_alone
    # a board left over from an older epoch is ignored (and overwritten by the newer one)
    preps[1].set_threshold_board(boards[1].data_ptr(), [boards[0].data_ptr()], 7)
    preps[1].run(sync=True)
    h, c = preps[1].fetch()
    assert h.tobytes() == alone[1][0].tobytes() and c.tobytes() == alone[1][1].tobytes()
    for p in preps:
        p.free()
    for gi in gis:
        gi.close()


# ---- AND batches on the posting scan: Bool{must:[..]} driven by the rarest list ------------------------------------------
@pytest.mark.parametrize("k", [11, 101])
def test_and_batch_on_the_posting_scan(k):
    """every query Bool{must: t1, t2(, t3)} (C4's shape), with and without a root filter, two segments, terms with and without
    columns / bitmaps, a term one segment lacks: the scan walks the rarest list only and must return what the reference's
    OR-scan + reject returns (api/reader.rs:1527-1563)"""
    from oracle import slo
    from searchlite_b200.shard import shard_ranges
    from tests.helpers import canonical_batch
    segs, oras = [], []
    for r, (lo, hi) in enumerate(shard_ranges(70_000, 2)):
        spec = synth.CorpusSpec(n_docs=hi - lo, vocab=3_000, seed=251, len_lo=20, len_hi=90, segment_ord=r, doc_base=lo)
        seg = synth.generate_segment(spec, "cpu")
        names, lang, year = synth.fast_fields(spec)
        seg.fast_str["lang"] = (names, lang)
        seg.fast_i64["year"] = (year, None)
        segs.append(seg)
        oras.append(slo.OracleIndex(seg))
    gi = GpuIndex(0, options={"dense_den": 16, "dense_min_df": 64, "bitmap_den": 256})
    cols = {}
    for seg in segs:
        cols = gi.load_segment(seg)
    assert any(gi.term_has_column(0, t) for t in range(1, 30))
    rng = np.random.default_rng(8)
    qs = []
    for i in range(120):
        pool = np.arange(1, 40) if i % 3 == 0 else np.arange(1, 400)   # dense-only queries, mixed queries
        t = rng.choice(pool, size=3, replace=False).tolist()
        if i % 7 == 0:
            t[1] = int(rng.integers(2_000, 2_990))  # rare term: often absent from one of the segments
        qs.append({"must": t[: 2 + i % 2]})
    prog = lambda c: np.concatenate([node(F_AND, nc=2), node(F_KEYWORD_EQ, c["lang"], v=(0, 1)), node(F_I64_RANGE, c["year"], i=(2003, 2014))])
    fid = gi.compile_filter(prog(cols), ["en"])
    for filtered in (False, True):
        qb = QueryBatch.from_bool([dict(q, filter_id=fid) for q in qs] if filtered else qs)
        kw = dict(filter_nodes=prog(oras[0].columns), strings=["en"]) if filtered else {}
        launches0 = gi.counters()["kernel_launches"]
        got = {mode: gi.search_batch(qb, k, mode) for mode in ("bm25", "bmw")}
        assert got["bm25"][0].tobytes() == got["bmw"][0].tobytes() and got["bm25"][1].tobytes() == got["bmw"][1].tobytes()
        assert gi.counters()["last_postings_scattered"] > 0  # the posting scan ran (its work counters are filled)
        n_hits = 0
        for qi in range(qb.n_queries):
            sub = qb.subset(qi, qi + 1)
            ref_q = slo.merge_hits([h[0, : c[0]] for h, c in (o.search_batch(sub, k, "bm25", **kw) for o in oras)], k)
            ref_c = slo.merge_hits([h[0, : c[0]] for h, c in (o.search_batch(canonical_batch(gi, sub, r), k, "bm25", **kw) for r, o in enumerate(oras))], k)
            g = got["bm25"][0][qi, : got["bm25"][1][qi]]
            assert g.tobytes() == ref_c.tobytes(), (filtered, qi)  # declared order: bit for bit
            assert_parity(ref_q[None, :], np.array([len(ref_q)]), g[None, :], np.array([len(g)]), strict=False)
            n_hits += len(g)
        assert n_hits > 0
    # statistics keep such a batch on the matcher kernel; the hits are the same docs
    qb = QueryBatch.from_bool(qs)
    h, c, st = gi.search_batch(qb, k, "bm25", want_stats=True)
    for qi in range(qb.n_queries):
        assert sorted(h[qi, : c[qi]]["doc_id"].tolist()) == sorted(got["bm25"][0][qi, : got["bm25"][1][qi]]["doc_id"].tolist()) or filtered
    gi.close()


def test_keyword_filter_lowercases_unicode_like_the_reference():
    """case_insensitive_equals (index/fastfields.rs:475-481) with non-ASCII values: the engine's bitmaps equal the oracle's
    and Python's str.lower()"""
    from tests.test_oracle_golden import UNICODE_WORDS as words
    seg = token_corpus([[0]] * len(words), 1)
    seg.fast_str["tag"] = (words, np.arange(len(words), dtype=np.uint32))
    ora = _oracle(seg)
    gi, col = _gpu(seg)
    for probe in words:
        fid = gi.compile_filter(node(F_KEYWORD_EQ, col["tag"], v=(0, 1)), [probe])
        bits = gi.filter_bitmap(fid, 0, len(words))
        assert np.array_equal(bits, ora.filter_bitmap(node(F_KEYWORD_EQ, ora.columns["tag"], v=(0, 1)), [probe])), probe
        got = [d for d in range(len(words)) if (bits[d >> 5] >> (d & 31)) & 1]
        assert got == [d for d, w in enumerate(words) if w.lower() == probe.lower()], probe
    gi.close()
