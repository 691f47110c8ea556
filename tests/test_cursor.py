"""Search-after cursors: the cursor branch of the accept closure (api/reader.rs:3019-3028), `saw_cursor`
(:2663, :2747-2749) and the 21-byte PaginationCursor codec (:614-691, :821-841).

CPU part: codec known answers and the reference's rejection tests (searchlite-core/tests/smoke.rs:618-851),
pagination through the oracle (smoke.rs:500-598, :853-950 restated).  GPU part: every page of every kernel equals
the corresponding slice of the un-paged ranking, bit for bit, and equals the oracle's page."""
import numpy as np
import pytest

from searchlite_b200 import engine
from searchlite_b200.engine import HIT_DTYPE, QueryBatch
from tests.helpers import segment_from_postings, token_corpus
from tests.parity import assert_parity


def _slo():
    from oracle import slo
    slo.build()
    return slo


# ---- codec ----------------------------------------------------------------------------------------
def test_cursor_codec_known_answer_and_round_trip():
    # version 1 | generation 7 | bits(1.5) = 0x3fc00000 | segment 2 | doc 300 | returned 10, big-endian, lowercase hex
    raw = engine.cursor_encode(7, 10, (2, 300, 1.5))
    assert raw == "01" "00000007" "3fc00000" "00000002" "0000012c" "0000000a"
    key, returned = engine.cursor_decode(raw, 7)
    assert key == (2, 300, np.float32(1.5)) and returned == 10
    rng = np.random.default_rng(1)
    for _ in range(50):
        hit = (int(rng.integers(0, 2**32)), int(rng.integers(0, 2**32)), np.float32(rng.random() * 40))
        gen, ret = int(rng.integers(0, 2**32)), int(rng.integers(0, 50001))
        key, returned = engine.cursor_decode(engine.cursor_encode(gen, ret, hit), gen)
        assert key[:2] == hit[:2] and key[2].view(np.uint32) == hit[2].view(np.uint32) and returned == ret
    assert engine.cursor_decode(raw.upper(), 7)[1] == 10  # u8::from_str_radix takes either case


def test_cursor_codec_rejections_follow_the_reference():
    raw = engine.cursor_encode(3, 2, (0, 5, 2.25))
    with pytest.raises(engine.SearchliteGpuError, match="invalid cursor length: expected 42 hex chars, got 6"):
        engine.cursor_decode("abcdef", 3)                                   # api/reader.rs:653-658
    with pytest.raises(engine.SearchliteGpuError, match="decoding cursor at byte index 1"):
        engine.cursor_decode(raw[:2] + "zz" + raw[4:], 3)                    # smoke.rs:618 cursor_rejects_invalid_hex
    tampered = ("b" if raw[0] == "a" else "a") + raw[1:]                    # smoke.rs:984-990 tamper_cursor
    with pytest.raises(engine.SearchliteGpuError, match="unsupported cursor version 161"):
        engine.cursor_decode(tampered, 3)                                   # smoke.rs:792 cursor_rejects_mismatched_position
    with pytest.raises(engine.SearchliteGpuError, match="cursor requests 60000 hits, which exceeds max supported 50000"):
        engine.cursor_decode(engine.cursor_encode(3, 60_000, (0, 5, 2.25)), 3)  # smoke.rs:730 cursor_rejects_excessive_advance
    with pytest.raises(engine.SearchliteGpuError, match="stale cursor for this index generation: expected 4, got 3"):
        engine.cursor_decode(raw, 4)                                        # api/reader.rs:829-835


# ---- pagination -----------------------------------------------------------------------------------
def _segments(rng, n_segs=2, n_docs=1500, vocab=25):
    segs = []
    for so in range(n_segs):
        toks = [rng.integers(0, vocab, size=int(rng.integers(2, 6))).tolist() for _ in range(n_docs)]  # short docs: many score ties
        segs.append(token_corpus(toks, vocab, segment_ord=so))
    return segs


def _oracle_search(slo, oras, qb, k, exe="bm25"):
    """per-segment oracle searches merged by SortKey (api/reader.rs:2777); returns hits, counts, saw_cursor"""
    per = [o.search_batch(qb, k, exe, want_stats=True) for o in oras]
    hits = np.zeros((qb.n_queries, k), dtype=HIT_DTYPE)
    counts = np.zeros(qb.n_queries, dtype=np.uint32)
    saw = np.zeros(qb.n_queries, dtype=bool)
    for qi in range(qb.n_queries):
        m = slo.merge_hits([h[qi, : c[qi]] for h, c, _ in per], k)
        hits[qi, : len(m)] = m
        counts[qi] = len(m)
        no_cursor = qb.has_cursor is None or not qb.has_cursor[qi]  # saw_cursor starts true then (api/reader.rs:2663)
        saw[qi] = no_cursor or any(st[qi]["saw_cursor"] for _, _, st in per)
    return hits, counts, saw


def _paginate(search, qb, limit, full_n):
    """IndexReader::search paging: k = limit + 1 (api/reader.rs:2595-2619), next cursor = key of hit[limit-1] (:2838-2851)"""
    pages = [[] for _ in range(qb.n_queries)]
    cursors = [None] * qb.n_queries
    for _ in range((full_n + limit - 1) // limit + 1):
        batch = qb.subset(0, qb.n_queries).set_cursors(cursors)
        hits, counts, saw = search(batch, limit + 1)
        assert saw.all()
        for qi in range(qb.n_queries):
            n = int(counts[qi])
            pages[qi] += [hits[qi, i] for i in range(min(n, limit))]
            if n > 0:  # next cursor: the last hit returned (an exhausted query then pages into nothing)
                last = hits[qi, min(n, limit) - 1]
                cursors[qi] = (int(last["segment_ord"]), int(last["doc_id"]), last["score"])
    return pages


def test_oracle_pages_concatenate_to_the_full_ranking():
    """smoke.rs:500-598 (pages of 2 cover all 6 docs) and :853-950 (equal scores: segment, then doc order)"""
    slo = _slo()
    # equal-score docs in two segments, as cursor_orders_stably_across_segments builds them
    segs = [segment_from_postings([([0, 1, 2], [1, 1, 1])], [4, 4, 4], segment_ord=so) for so in range(2)]
    oras = [slo.OracleIndex(s) for s in segs]
    qb = QueryBatch.from_term_lists([[0]])
    got, cursor = [], None
    for _ in range(4):
        h, c, saw = _oracle_search(slo, oras, qb.subset(0, 1).set_cursors([cursor]), 3)
        assert saw.all()
        got += [(int(x["segment_ord"]), int(x["doc_id"])) for x in h[0, : min(2, c[0])]]
        if c[0] <= 2:
            break
        cursor = (int(h[0, 1]["segment_ord"]), int(h[0, 1]["doc_id"]), h[0, 1]["score"])
    assert got == [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (1, 2)]
    # a cursor that names no doc of the result set: "stale or invalid cursor for this result set" (api/reader.rs:2747-2749)
    h, c, saw = _oracle_search(slo, oras, qb.subset(0, 1).set_cursors([(0, 1, np.float32(123.0))]), 3)
    assert not saw[0]


def test_oracle_random_pagination():
    slo = _slo()
    rng = np.random.default_rng(9)
    segs = _segments(rng, n_docs=300)
    oras = [slo.OracleIndex(s) for s in segs]
    qb = QueryBatch.from_term_lists([rng.choice(25, size=2, replace=False).tolist() for _ in range(4)])
    full_h, full_c, _ = _oracle_search(slo, oras, qb, 601)
    limit = 37
    pages = _paginate(lambda b, k: _oracle_search(slo, oras, b, k), qb, limit, int(full_c.max()))
    for qi in range(qb.n_queries):
        want = full_h[qi, : full_c[qi]]
        assert len(pages[qi]) == len(want)
        assert np.array(pages[qi], dtype=HIT_DTYPE).tobytes() == want.tobytes()


# ---- GPU ------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["auto", "cta", "warp", "reg"])
@pytest.mark.parametrize("exe", ["bm25", "bmw"])
def test_gpu_pagination_matches_full_ranking_and_oracle(kernel, exe):
    from searchlite_b200 import GpuIndex
    slo = _slo()
    rng = np.random.default_rng(17)
    segs = _segments(rng)
    oras = [slo.OracleIndex(s) for s in segs]
    gi = GpuIndex(0, kernel=kernel)
    for s in segs:
        gi.load_segment(s)
    qb = QueryBatch.from_term_lists([rng.choice(25, size=int(rng.integers(1, 4)), replace=False).tolist() for _ in range(24)])
    limit = 5
    n_pages = 6  # k = 31 for the un-paged ranking: within the warp kernels' k <= 32
    full_h, full_c = gi.search_batch(qb, limit * n_pages + 1, exe)

    def search(batch, k):
        p = gi.prepare(batch, k, exe)
        p.run(sync=True)
        h, c = p.fetch()[:2]
        saw = p.cursor_seen()
        p.free()
        return h, c, saw

    cursors = [None] * qb.n_queries
    for page in range(n_pages):
        batch = qb.subset(0, qb.n_queries).set_cursors(cursors)
        h, c, saw = search(batch, limit + 1)
        assert saw.all(), (kernel, exe, page)
        oh, oc, osaw = _oracle_search(slo, oras, batch, limit + 1)
        assert osaw.all()
        assert_parity(oh, oc, h, c, strict=kernel in ("cta", "warp"))
        for qi in range(qb.n_queries):
            want = full_h[qi, page * limit: min(int(full_c[qi]), page * limit + limit + 1)]
            assert h[qi, : c[qi]].tobytes() == want.tobytes(), (kernel, exe, page, qi)
            if c[qi] > limit:
                last = h[qi, limit - 1]
                cursors[qi] = (int(last["segment_ord"]), int(last["doc_id"]), last["score"])
            elif c[qi] > 0:
                last = h[qi, c[qi] - 1]
                cursors[qi] = (int(last["segment_ord"]), int(last["doc_id"]), last["score"])
    # a cursor whose doc is not in the result set is reported (the reference fails the request)
    bad = qb.subset(0, 2).set_cursors([(0, 3, np.float32(77.0)), None])
    h, c, saw = search(bad, limit + 1)
    assert saw.tolist() == [False, True]
    gi.close()


@pytest.mark.gpu
def test_gpu_cursor_with_matcher_filter_and_plan():
    from searchlite_b200 import GpuIndex
    slo = _slo()
    rng = np.random.default_rng(23)
    toks = [rng.integers(0, 20, size=int(rng.integers(2, 8))).tolist() for _ in range(3000)]
    seg = token_corpus(toks, 20)
    seg.deleted_docs = np.sort(rng.choice(3000, size=200, replace=False)).astype(np.uint32)
    ora = slo.OracleIndex(seg)
    gi = GpuIndex(0)
    gi.load_segment(seg)
    queries = [{"must": [int(a)], "should": [int(b), int(c)], "must_not": [int(d)]}
               for a, b, c, d in (rng.choice(20, size=4, replace=False) for _ in range(12))]
    exprs = [("sum", [("leaf", 0), ("dismax", [("leaf", 1), ("leaf", 2)], 0.5)]) if i % 2 else None for i in range(12)]
    qb = QueryBatch.from_bool(queries).set_plans(exprs)
    full_h, full_c = gi.search_batch(qb, 41, "bm25")
    cursors = []
    for qi in range(12):
        n = int(full_c[qi])
        cursors.append(None if n < 8 else (0, int(full_h[qi, 6]["doc_id"]), full_h[qi, 6]["score"]))
    batch = qb.subset(0, 12).set_cursors(cursors)
    for exe in ("bm25", "wand"):
        p = gi.prepare(batch, 11, exe)
        p.run(sync=True)
        h, c = p.fetch()[:2]
        assert p.cursor_seen().all()
        p.free()
        oh, oc = ora.search_batch(batch, 11, "bm25")
        assert_parity(oh, oc, h, c, strict=True)
        for qi in range(12):
            lo = 0 if cursors[qi] is None else 7
            assert h[qi, : c[qi]].tobytes() == full_h[qi, lo: min(int(full_c[qi]), lo + 11)].tobytes(), (exe, qi)
    gi.close()


# ---- match counter (total_hits_estimate, api/reader.rs:3029-3031, :2825) ----------------------------
def test_oracle_total_hits_counts_across_pages():
    """smoke.rs:594-616 total_hits_estimate_counts_across_pages: 3 matching docs, limit 2 -> the second page still
    reports 3 = total_matches (docs after the cursor) + cursor.returned"""
    slo = _slo()
    seg = segment_from_postings([([0, 1, 2], [1, 1, 1])], [2, 3, 3])
    ora = slo.OracleIndex(seg)
    qb = QueryBatch.from_term_lists([[0]])
    h, c, st = ora.search_batch(qb, 3, "bm25", want_stats=True)
    assert st["total_matches"][0] == 3 and c[0] == 3
    cur = (0, int(h[0, 1]["doc_id"]), h[0, 1]["score"])
    h2, c2, st2 = ora.search_batch(qb.subset(0, 1).set_cursors([cur]), 3, "bm25", want_stats=True)
    assert c2[0] == 1 and st2["total_matches"][0] + 2 == 3


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["auto", "cta", "warp"])
def test_gpu_total_matches_equal_the_oracle_counter(kernel):
    from searchlite_b200 import GpuIndex
    slo = _slo()
    rng = np.random.default_rng(31)
    toks = [rng.integers(0, 30, size=int(rng.integers(2, 9))).tolist() for _ in range(5000)]
    seg = token_corpus(toks, 30)
    seg.deleted_docs = np.sort(rng.choice(5000, size=400, replace=False)).astype(np.uint32)
    ora = slo.OracleIndex(seg)
    gi = GpuIndex(0, kernel=kernel)
    gi.load_segment(seg)
    plain = QueryBatch.from_term_lists([rng.choice(30, size=int(rng.integers(1, 5)), replace=False).tolist() for _ in range(20)])
    boolq = QueryBatch.from_bool([{"must": [int(a)], "should": [int(b), int(c)], "must_not": [int(d)], "min_should": int(m)}
                                  for a, b, c, d, m in ((*rng.choice(30, size=4, replace=False), rng.integers(0, 2)) for _ in range(20))])
    for qb in (plain, boolq):
        h, c, st = gi.search_batch(qb, 11, "bm25", want_stats=True)
        oh, oc, ost = ora.search_batch(qb, 11, "bm25", want_stats=True)
        assert st["total_matches"].tolist() == ost["total_matches"].tolist()
        # second page: the counter only sees docs after the cursor; + cursor.returned gives the same total (:2825)
        cursors = [None if c[q] < 6 else (0, int(h[q, 4]["doc_id"]), h[q, 4]["score"]) for q in range(qb.n_queries)]
        page2 = qb.subset(0, qb.n_queries).set_cursors(cursors)
        h2, c2, st2 = gi.search_batch(page2, 11, "bm25", want_stats=True)
        oh2, oc2, ost2 = ora.search_batch(page2, 11, "bm25", want_stats=True)
        assert st2["total_matches"].tolist() == ost2["total_matches"].tolist()
        for q in range(qb.n_queries):
            if cursors[q] is not None:
                assert st2["total_matches"][q] + 5 == st["total_matches"][q]
        # pruned execution: an estimate from below, exact hits
        hp, cp, stp = gi.search_batch(qb, 11, "bmw", want_stats=True)
        assert hp.tobytes() == h.tobytes()
        assert (stp["total_matches"] <= st["total_matches"]).all() and (stp["total_matches"] >= cp).all()
    gi.close()
