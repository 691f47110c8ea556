"""GPU parity for SURVEY §8f rows 1-2: segments loaded from the reference's own on-disk files and phrase
matching on resident positions, both through the C ABI and both against the oracle."""
from __future__ import annotations

import numpy as np
import pytest

from searchlite_b200 import GpuIndex
from searchlite_b200.engine import ABSENT_TERM, FILTER_DTYPE, F_AND, F_I64_RANGE, F_KEYWORD_EQ, QueryBatch, SearchliteGpuError, SegmentData
from tests import segwriter as sw
from tests.parity import assert_parity

pytestmark = pytest.mark.gpu


def make_docs(rng, n_docs, vocab, lo=4, hi=60, zipf=1.3, shift=0):
    docs = []
    for _ in range(n_docs):
        n = int(rng.integers(lo, hi))
        toks = ((rng.zipf(zipf, size=n) + shift) % vocab).tolist()
        docs.append({"body": [f"w{t}" for t in toks], "title": [f"w{t}" for t in toks[:2]]})
    return docs


def two_segments(seed=21, n0=1800, n1=1300, vocab=220):
    rng = np.random.default_rng(seed)
    segs = []
    for sid, n, shift in (("s0", n0, 0), ("s1", n1, 37)):  # different shifts: each segment holds keys the other lacks
        docs = make_docs(rng, n, vocab if sid == "s0" else vocab + 40, shift=shift)
        langs = [None if i % 17 == 0 else ["en", "de", "fr", "ES"][int(rng.integers(0, 4))] for i in range(n)]
        years = [None if i % 19 == 0 else int(rng.integers(2000, 2026)) for i in range(n)]
        vecs = []
        for i in range(n):
            if i % 7 == 3:
                vecs.append(None)
            else:
                v = rng.standard_normal(16).astype(np.float32)
                vecs.append(v / np.linalg.norm(v))
        segs.append(sw.Segment(sid, docs, ["body", "title"], keywords={"lang": langs}, i64s={"year": years},
                               i64_lists={"tags": [[i % 7, (i * 3) % 11][: i % 3] for i in range(n)]},
                               keyword_lists={"labels": [["Red", "green", "BLUE", "red"][i % 2: i % 2 + i % 4] for i in range(n)]},
                               i64_nested={"sizes": [[[(i + j) % 9, 40 + j][: 1 + (i + j) % 2] for j in range(i % 3)] for i in range(n)]},
                               keyword_nested={"owners": [[["ann", "Bob"][: 1 + j] for j in range(i % 3)] for i in range(n)]},
                               vectors={"emb": ("cosine", vecs)},
                               deleted=[5, 77, n - 1] if sid == "s1" else []))
    return segs


def global_csr(gi, seg, ord_):
    """the segment as CSR SegmentData in the file-loaded handle's term space (+ positions)"""
    keys, toff, docs, tfs, poff, pos, lens = sw.csr_of(seg, "body")
    ids = [gi.term_lookup(k) for k in keys]
    assert ABSENT_TERM not in ids
    n_terms = max(ids) + 1
    order = np.argsort(ids)
    g_toff = np.zeros(n_terms + 1, dtype=np.uint64)
    g_docs, g_tfs, g_poff, g_pos = [], [], [0], []
    cur = 0
    next_id = 0
    for j in order:
        tid = ids[j]
        while next_id <= tid:
            g_toff[next_id] = cur
            next_id += 1
        a, b = int(toff[j]), int(toff[j + 1])
        g_docs.append(docs[a:b])
        g_tfs.append(tfs[a:b])
        for p in range(a, b):
            g_pos.append(pos[int(poff[p]):int(poff[p + 1])])
            g_poff.append(g_poff[-1] + int(poff[p + 1] - poff[p]))
        cur += b - a
    g_toff[next_id:] = cur
    sd = SegmentData(ord_, len(seg.docs), g_toff, np.concatenate(g_docs).astype(np.uint32), np.concatenate(g_tfs).astype(np.uint32),
                     lens, int(lens.sum()), deleted_docs=np.asarray(seg.deleted, dtype=np.uint32) if seg.deleted else None)
    return sd, np.asarray(g_poff, dtype=np.uint64), np.concatenate(g_pos).astype(np.uint32)


def random_queries(rng, gi, keys, n_q, with_absent=True):
    lists = []
    for _ in range(n_q):
        n = int(rng.integers(1, 5))
        ks = rng.choice(len(keys), size=n, replace=False)
        t = [gi.term_lookup(keys[i]) for i in ks]
        if with_absent and rng.random() < 0.1:
            t.append(gi.term_lookup("body:nosuchtoken"))
        lists.append(t)
    return QueryBatch.from_term_lists(lists)


@pytest.fixture(scope="module")
def index_dir(tmp_path_factory):
    root = tmp_path_factory.mktemp("refindex")
    segs = two_segments()
    # the manifest records the writer's paths; the directory has "moved" since
    sw.write_index(str(root), segs, stored_root="/var/lib/searchlite/idx")
    return str(root), segs


def oracle_merged(segs_csr, qb, k, **kw):
    from oracle import slo
    per = [slo.OracleIndex(sd).search_batch(qb, k, "bm25", **kw) for sd in segs_csr]
    hits = np.zeros((qb.n_queries, k), dtype=per[0][0].dtype)
    counts = np.zeros(qb.n_queries, dtype=np.uint32)
    for q in range(qb.n_queries):
        m = slo.merge_hits([h[q, : c[q]] for h, c in per], k)
        hits[q, : len(m)] = m
        counts[q] = len(m)
    return hits, counts


def test_index_dir_load_equals_csr_load_and_oracle(index_dir):
    root, segs = index_dir
    gi = GpuIndex(0, kernel="warp")
    assert gi.load_index_dir(root, "body") == 2
    keys_all = sorted({k for s in segs for k in s.postings() if k.startswith("body:")})
    assert gi.term_lookup("title:w1") == ABSENT_TERM and gi.term_lookup("body:nosuchtoken") == ABSENT_TERM
    csr = [global_csr(gi, s, i)[0] for i, s in enumerate(segs)]
    from oracle import slo
    for i, sd in enumerate(csr):
        st, o = gi.segment_stats(i), slo.OracleIndex(sd)
        assert st["n_postings"] == len(sd.post_docs)
        for name, want in (("avgdl", o.avgdl), ("live_docs", o.live_docs), ("min_doc_len", o.min_doc_len)):
            assert np.float32(st[name]).view(np.uint32) == np.float32(want).view(np.uint32), (i, name)
    rng = np.random.default_rng(4)
    qb = random_queries(rng, gi, keys_all, 90)
    k = 11
    got = gi.search_batch(qb, k, "bm25")
    assert_parity(*oracle_merged(csr, qb, k), *got, strict=True)
    # the same segments through the CSR path give the same bytes
    g2 = GpuIndex(0, kernel="warp")
    for sd in csr:
        g2.load_segment(sd)
    ref = g2.search_batch(qb, k, "bm25")
    assert ref[0].tobytes() == got[0].tobytes() and ref[1].tobytes() == got[1].tobytes()
    # pruned execution and the automatic kernel agree with the rule
    assert_parity(*oracle_merged(csr, qb, k), *gi.search_batch(qb, k, "bmw"), strict=True)
    ga = GpuIndex(0)
    ga.load_index_dir(root, "body")
    assert_parity(*oracle_merged(csr, qb, k), *ga.search_batch(qb, k, "bm25"), strict=False)
    # one term space per handle: file keys and caller-assigned ids do not mix
    with pytest.raises(SearchliteGpuError, match="caller-assigned"):
        gi.load_segment(csr[0])
    with pytest.raises(SearchliteGpuError, match="caller-assigned"):
        g2.load_index_dir(root, "body")
    with pytest.raises(SearchliteGpuError, match="same field list"):
        gi.load_index_dir(root, "title")
    for h in (gi, g2, ga):
        h.close()


def test_sharded_index_dir_equals_whole(index_dir):
    """one process per GPU loads the manifest segments of its rank (segment == shard); the merged local lists equal the
    single-handle result — term ids differ per handle, keys are resolved by each"""
    root, segs = index_dir
    from oracle import slo
    whole = GpuIndex(0, kernel="warp")
    assert whole.load_index_dir(root, "body") == 2
    parts = []
    for r in range(2):
        g = GpuIndex(0, kernel="warp")
        assert g.load_index_dir(root, "body", shard_rank=r, shard_world=2) == 1
        parts.append(g)
    keys_all = sorted({k for s in segs for k in s.postings() if k.startswith("body:")})
    rng = np.random.default_rng(14)
    qkeys = [[keys_all[i] for i in rng.choice(len(keys_all), size=int(rng.integers(1, 5)), replace=False)] for _ in range(50)]
    batch = lambda g: QueryBatch.from_term_lists([[g.term_lookup(k) for k in q] for q in qkeys])
    wh, wc = whole.search_batch(batch(whole), 11, "bm25")
    per = [g.search_batch(batch(g), 11, "bm25") for g in parts]
    assert set(np.unique(per[1][0]["segment_ord"][per[1][1] > 0][:, 0]).tolist()) <= {1}
    for q in range(len(qkeys)):
        want = slo.merge_hits([h[q, : c[q]] for h, c in per], 11)
        g = wh[q, : wc[q]]
        assert len(g) == len(want) and np.array_equal(g["segment_ord"], want["segment_ord"]) and np.array_equal(g["doc_id"], want["doc_id"])
        assert np.array_equal(g["score"].view(np.uint32), want["score"].view(np.uint32))
    with pytest.raises(SearchliteGpuError, match="shard"):
        GpuIndex(0).load_index_dir(root, "body", shard_rank=2, shard_world=2)
    for h in [whole] + parts:
        h.close()


def test_file_columns_filter_like_the_oracle(index_dir):
    root, segs = index_dir
    from oracle import slo
    gi = GpuIndex(0, kernel="warp")
    gi.load_index_dir(root, "body")
    lang, year = gi.column_lookup("lang"), gi.column_lookup("year")
    assert lang >= 0 and year >= 0 and gi.column_lookup("_len:body") == -1
    # list and nested columns of the files: any value of the doc (index/fastfields.rs:490-657)
    tagc, labc, sizc, ownc = (gi.column_lookup(x) for x in ("tags", "labels", "sizes", "owners"))
    assert min(tagc, labc, sizc, ownc) >= 0
    lnodes = np.zeros(1, dtype=FILTER_DTYPE)
    checks = [
        ((F_I64_RANGE, tagc, 5, 6, 0, 0, 0, 0, 0), [], lambda s, d: any(5 <= v <= 6 for v in s.i64_lists["tags"][d])),
        ((F_KEYWORD_EQ, labc, 0, 0, 0, 0, 0, 0, 1), ["RED"], lambda s, d: any(v.lower() == "red" for v in s.keyword_lists["labels"][d])),
        ((F_I64_RANGE, sizc, 40, 40, 0, 0, 0, 0, 0), [], lambda s, d: any(v == 40 for o in s.i64_nested["sizes"][d] for v in o)),
        ((F_KEYWORD_EQ, ownc, 0, 0, 0, 0, 0, 0, 1), ["bob"], lambda s, d: any(v.lower() == "bob" for o in s.keyword_nested["owners"][d] for v in o)),
    ]
    for row, strings, want in checks:
        lnodes[0] = row
        lf = gi.compile_filter(lnodes, strings)
        for i, s in enumerate(segs):
            bm = gi.filter_bitmap(lf, i, len(s.docs))
            got = [bool((bm[d >> 5] >> (d & 31)) & 1) for d in range(len(s.docs))]
            exp = [want(s, d) for d in range(len(s.docs))]
            assert got == exp and 0 < sum(exp) < len(exp), row
    nodes = np.zeros(3, dtype=FILTER_DTYPE)
    nodes[0] = (F_AND, -1, 0, 0, 0, 0, 2, 0, 0)
    nodes[1] = (F_KEYWORD_EQ, lang, 0, 0, 0, 0, 0, 0, 1)
    nodes[2] = (F_I64_RANGE, year, 2005, 2015, 0, 0, 0, 0, 0)
    fid = gi.compile_filter(nodes, ["es"])  # ASCII case-insensitive: matches "ES" (fastfields.rs:475-481)
    for i, s in enumerate(segs):
        bm = gi.filter_bitmap(fid, i, len(s.docs))
        langs, years = s.keywords["lang"], s.i64s["year"]
        for d in range(len(s.docs)):
            want = langs[d] is not None and langs[d].lower() == "es" and years[d] is not None and 2005 <= years[d] <= 2015
            assert bool((bm[d >> 5] >> (d & 31)) & 1) == want, (i, d)
    # filtered search against the oracle with the same columns
    csr = [global_csr(gi, s, i)[0] for i, s in enumerate(segs)]
    for sd, s in zip(csr, segs):
        dic = sorted({v for v in s.keywords["lang"] if v is not None})
        ords = np.array([0xFFFFFFFF if v is None else dic.index(v) for v in s.keywords["lang"]], dtype=np.uint32)
        sd.fast_str = {"lang": (dic, ords)}
        sd.fast_i64 = {"year": (np.array([0 if v is None else v for v in s.i64s["year"]], dtype=np.int64),
                                np.array([v is not None for v in s.i64s["year"]], dtype=np.uint8))}
    keys_all = sorted({k for s in segs for k in s.postings() if k.startswith("body:")})
    qb = random_queries(np.random.default_rng(6), gi, keys_all, 60)
    qb.filter_id = np.full(qb.n_queries, fid, dtype=np.int32)
    onodes = nodes.copy()
    per = []
    for sd in csr:
        o = slo.OracleIndex(sd)
        onodes[1]["column"], onodes[2]["column"] = o.columns["lang"], o.columns["year"]
        per.append(o.search_batch(qb, 11, "bm25", filter_nodes=onodes, strings=["es"]))
    got_h, got_c = gi.search_batch(qb, 11, "bm25")
    for q in range(qb.n_queries):
        want = slo.merge_hits([h[q, : c[q]] for h, c in per], 11)
        g = got_h[q, : got_c[q]]
        assert len(g) == len(want) and np.array_equal(g["doc_id"], want["doc_id"]) and np.array_equal(g["segment_ord"], want["segment_ord"])
        assert np.array_equal(g["score"].view(np.uint32), want["score"].view(np.uint32))
    gi.close()


def test_vector_file_rerank(index_dir):
    root, segs = index_dir
    from oracle import slo
    gi = GpuIndex(0)
    gi.load_index_dir(root, "body", vector_field="emb")
    keys_all = sorted({k for s in segs for k in s.postings() if k.startswith("body:")})
    qb = random_queries(np.random.default_rng(8), gi, keys_all, 12, with_absent=False)
    hits, counts = gi.search_batch(qb, 50, "bm25")
    qv = np.random.default_rng(9).standard_normal((12, 16)).astype(np.float32)
    qv /= np.linalg.norm(qv, axis=1, keepdims=True)
    out, vs = gi.rerank(qv, hits, counts, 0.5, "cosine")
    L = slo.lib()
    for q in range(12):
        want = []
        for h in hits[q, : counts[q]]:
            v = segs[int(h["segment_ord"])].vectors["emb"][1][int(h["doc_id"])]
            if v is None:
                sc = L.slo_hybrid_score(float(h["score"]), 0, 0.0, 0.5, 0)
            else:
                a = np.ascontiguousarray(qv[q])
                b = np.ascontiguousarray(v, dtype=np.float32)
                sc = L.slo_hybrid_score(float(h["score"]), 1, L.slo_metric_similarity(0, a.ctypes.data, b.ctypes.data, 16), 0.5, 0)
            want.append((int(h["segment_ord"]), int(h["doc_id"]), sc))
        got = {(int(h["segment_ord"]), int(h["doc_id"])): float(h["score"]) for h in out[q, : counts[q]]}
        for so, d, sc in want:
            assert abs(got[(so, d)] - sc) <= 2e-5 * max(1.0, abs(sc)), (q, so, d)
    gi.close()


def test_bad_files_fail_loudly(index_dir, tmp_path):
    root, segs = index_dir
    import shutil
    bad = tmp_path / "bad"
    shutil.copytree(root, bad)
    p = bad / "seg_s1.post"
    data = bytearray(p.read_bytes())
    data[len(data) // 3] ^= 0x10
    p.write_bytes(bytes(data))
    gi = GpuIndex(0)
    with pytest.raises(SearchliteGpuError, match="failed checksum for postings"):
        gi.load_index_dir(str(bad), "body")
    with pytest.raises(SearchliteGpuError, match="MANIFEST"):
        gi.load_index_dir(str(tmp_path / "nope"), "body")
    gi.close()


# ---- phrases ----
def phrase_docs(seed=31, n_docs=2500, vocab=30):
    rng = np.random.default_rng(seed)
    return sw.Segment("p0", make_docs(rng, n_docs, vocab, lo=1, hi=50, zipf=1.15), ["body", "title"])


def bitmap_list(bm, n):
    return [d for d in range(n) if (bm[d >> 5] >> (d & 31)) & 1]


@pytest.mark.parametrize("source", ["files", "csr"])
def test_phrase_bitmaps_match_the_oracle(source, tmp_path):
    from oracle import slo
    seg = phrase_docs()
    n = len(seg.docs)
    gi = GpuIndex(0, kernel="warp")
    if source == "files":
        sw.write_index(str(tmp_path), [seg])
        gi.load_index_dir(str(tmp_path), "body")
    else:
        keys, toff, docs, tfs, poff, pos, lens = sw.csr_of(seg, "body")
        gi.load_segment(SegmentData(0, n, toff, docs, tfs, lens, int(lens.sum())))
        with pytest.raises(SearchliteGpuError, match="no term positions"):
            gi.compile_phrase([0, 1], 0)
        gi.load_positions(0, toff, poff, pos)
    keys, toff, docs, tfs, poff, pos, lens = sw.csr_of(seg, "body")
    tid = (lambda k: gi.term_lookup(k)) if source == "files" else (lambda k: keys.index(k))
    # oracle in the engine's term space
    if source == "files":
        sd, g_poff, g_pos = global_csr(gi, seg, 0)
        o_toff, o_docs = sd.term_offsets, sd.post_docs
    else:
        o_toff, o_docs, g_poff, g_pos = toff, docs, poff, pos
    rng = np.random.default_rng(2)
    cases = [(["w1", "w2"], 0), (["w2", "w1"], 0), (["w1", "w2", "w3"], 0), (["w1", "w2", "w3"], 2), (["w1", "w1"], 0),
             (["w5"], 0), (["w1", "w9", "w2", "w4"], 6), (["w3", "w2"], 1)]
    for _ in range(30):
        m = int(rng.integers(2, 5))
        cases.append(([f"w{int(t)}" for t in rng.integers(1, 14, size=m)], int(rng.integers(0, 5))))
    n_nonempty = 0
    for toks, slop in cases:
        ids = [tid(f"body:{t}") for t in toks]
        fid = gi.compile_phrase(ids, slop)
        got = gi.filter_bitmap(fid, 0, n)
        want = slo.phrase_bitmap(n, o_toff, o_docs, g_poff, g_pos, ids, slop)
        assert np.array_equal(got, want), (toks, slop, bitmap_list(got, n)[:8], bitmap_list(want, n)[:8])
        n_nonempty += bool(want.any())
        gi.free_filter(fid)
    assert n_nonempty > 20
    # the whole batch in one launch gives the same bitmaps; freeing one id leaves its slab neighbours alone
    batch_ids = gi.compile_phrases([[tid(f"body:{t}") for t in toks] for toks, _ in cases] + [[tid("body:w1"), ABSENT_TERM]],
                                   [sl for _, sl in cases] + [0])
    assert len(set(batch_ids.tolist())) == len(cases) + 1
    gi.free_filter(int(batch_ids[1]))
    for (toks, slop), fid in list(zip(cases, batch_ids))[::3]:
        want = slo.phrase_bitmap(n, o_toff, o_docs, g_poff, g_pos, [tid(f"body:{t}") for t in toks], slop)
        assert np.array_equal(gi.filter_bitmap(int(fid), 0, n), want), (toks, slop)
    assert not gi.filter_bitmap(int(batch_ids[-1]), 0, n).any()
    with pytest.raises(SearchliteGpuError):
        gi.filter_bitmap(int(batch_ids[1]), 0, n)
    # a term the segment lacks: no doc matches (api/reader.rs:1690-1697)
    fid = gi.compile_phrase([tid("body:w1"), ABSENT_TERM], 0)
    assert not gi.filter_bitmap(fid, 0, n).any()
    gi.close()


def test_phrase_query_search_and_bitmap_algebra(tmp_path):
    """QueryString with a required phrase (api/reader.rs:1504-1508) behind a root filter: accept = phrase AND filter"""
    from oracle import slo
    seg = phrase_docs(seed=33)
    n = len(seg.docs)
    seg.i64s = {"year": [2000 + d % 20 for d in range(n)]}
    sw.write_index(str(tmp_path), [seg])
    gi = GpuIndex(0, kernel="warp")
    gi.load_index_dir(str(tmp_path), "body")
    sd, g_poff, g_pos = global_csr(gi, seg, 0)
    t1, t2, t3 = (gi.term_lookup(f"body:w{i}") for i in (1, 2, 3))
    ph = gi.compile_phrase([t1, t2], 1)
    nodes = np.zeros(1, dtype=FILTER_DTYPE)
    nodes[0] = (F_I64_RANGE, gi.column_lookup("year"), 2003, 2012, 0, 0, 0, 0, 0)
    flt = gi.compile_filter(nodes)
    both = gi.combine_filters("and", ph, flt)
    either = gi.combine_filters("or", ph, flt)
    minus = gi.combine_filters("and_not", flt, ph)
    a, b = gi.filter_bitmap(ph, 0, n), gi.filter_bitmap(flt, 0, n)
    assert np.array_equal(gi.filter_bitmap(both, 0, n), a & b)
    assert np.array_equal(gi.filter_bitmap(either, 0, n), a | b)
    assert np.array_equal(gi.filter_bitmap(minus, 0, n), b & ~a)
    for op, expect in (("and", a & b), ("or", a | b), ("and_not", b & ~a)):
        x, y = (flt, ph) if op == "and_not" else (ph, flt)
        out = gi.combine_filters_batch(op, [x, x, either], [y, y, either])
        assert np.array_equal(gi.filter_bitmap(int(out[0]), 0, n), expect) and np.array_equal(gi.filter_bitmap(int(out[1]), 0, n), expect)
        assert np.array_equal(gi.filter_bitmap(int(out[2]), 0, n), (a | b) if op != "and_not" else np.zeros_like(a))
    want_bm = slo.phrase_bitmap(n, sd.term_offsets, sd.post_docs, g_poff, g_pos, [t1, t2], 1) & b
    # query: w1 w2 w3 scored, phrase "w1 w2"~1 required
    qb = QueryBatch.from_term_lists([[t1, t2, t3], [t3], [t1, t2]])
    qb.filter_id = np.array([both, both, ph], dtype=np.int32)
    k = 11
    got_h, got_c = gi.search_batch(qb, k, "bm25")
    pr_h, pr_c = gi.search_batch(qb, k, "bmw")
    assert got_h.tobytes() == pr_h.tobytes() and got_c.tobytes() == pr_c.tobytes()
    o = slo.OracleIndex(sd)
    qo = QueryBatch.from_term_lists([[t1, t2, t3], [t3], [t1, t2]])
    full_h, full_c = o.search_batch(qo, n + 1, "bm25")
    for q, bm in enumerate((want_bm, want_bm, a)):
        keep = [h for h in full_h[q, : full_c[q]] if (bm[int(h["doc_id"]) >> 5] >> (int(h["doc_id"]) & 31)) & 1][:k]
        g = got_h[q, : got_c[q]]
        assert len(g) == len(keep) and len(keep) > 0
        assert [int(h["doc_id"]) for h in keep] == g["doc_id"].tolist()
        assert [np.float32(h["score"]).view(np.uint32) for h in keep] == g["score"].view(np.uint32).tolist()
    gi.free_filter(both)
    with pytest.raises(SearchliteGpuError, match="freed"):
        gi.search_batch(qb, k, "bm25")
    gi.close()


def test_positions_can_be_dropped(tmp_path):
    seg = phrase_docs(seed=35, n_docs=200)
    sw.write_index(str(tmp_path), [seg])
    gi = GpuIndex(0)
    gi.set_option("keep_positions", 0)
    gi.load_index_dir(str(tmp_path), "body")
    with pytest.raises(SearchliteGpuError, match="no term positions"):
        gi.compile_phrase([0, 1], 0)
    gi.close()


# ---- several text fields in one handle ----
def test_multi_field_keys_use_their_own_norms(index_dir):
    """QueryString over default fields [body, title] (query/planner.rs:284-312): per term one key per field; every key
    is scored with its field's lengths / avgdl / min length (api/reader.rs:2990-2994).  Expected values: the oracle
    scores each key alone in a single-field index of its field; the engine's sum is the f32 left fold in query order."""
    root, segs = index_dir
    from oracle import slo
    gi = GpuIndex(0, kernel="warp")
    assert gi.load_index_dir(root, "body,title") == 2
    per_field = {}
    for si, s in enumerate(segs):
        for fi, field in enumerate(("body", "title")):
            keys, toff, docs, tfs, poff, pos, lens = sw.csr_of(s, field)
            sd = SegmentData(si, len(s.docs), toff, docs, tfs, lens, int(lens.sum()),
                             deleted_docs=np.asarray(s.deleted, dtype=np.uint32) if s.deleted else None)
            o = slo.OracleIndex(sd)
            per_field[(si, field)] = (keys, o)
            st = gi.field_stats(si, fi)
            assert np.float32(st["avgdl"]).view(np.uint32) == np.float32(o.avgdl).view(np.uint32), (si, field)
            assert np.float32(st["min_doc_len"]).view(np.uint32) == np.float32(o.min_doc_len).view(np.uint32), (si, field)

    cache = {}

    def key_scores(si, key):
        """doc -> f32 score of one key alone in segment si (deleted docs absent)"""
        if (si, key) not in cache:
            cache[(si, key)] = _key_scores(si, key)
        return cache[(si, key)]

    def _key_scores(si, key):
        keys, o = per_field[(si, key.split(":")[0])]
        if key not in keys:
            return {}
        h, c = o.search_batch(QueryBatch.from_term_lists([[keys.index(key)]]), len(segs[si].docs) + 1, "bm25")
        return {int(x["doc_id"]): np.float32(x["score"]) for x in h[0, : c[0]]}

    rng = np.random.default_rng(12)
    queries = []
    for _ in range(40):
        toks = rng.choice(np.arange(1, 60), size=int(rng.integers(1, 4)), replace=False)
        queries.append([f"{f}:w{int(t)}" for t in toks for f in ("body", "title")])
    queries.append(["title:w1", "body:nosuch", "title:w2"])
    qb = QueryBatch.from_term_lists([[gi.term_lookup(k) for k in q] for q in queries])
    k = 11
    for exe in ("bm25", "bmw"):
        got_h, got_c = gi.search_batch(qb, k, exe)
        for qi, q in enumerate(queries):
            rows = []
            for si in range(len(segs)):
                total = {}
                for key in q:
                    for d, sc in key_scores(si, key).items():
                        total[d] = np.float32(total[d] + sc) if d in total else sc
                rows += [(-float(sc), si, d, sc) for d, sc in total.items()]
            rows.sort(key=lambda r: (r[0], r[1], r[2]))
            want = rows[:k]
            g = got_h[qi, : got_c[qi]]
            assert len(g) == len(want), (exe, qi)
            assert [(int(x["segment_ord"]), int(x["doc_id"])) for x in g] == [(r[1], r[2]) for r in want], (exe, qi, q)
            assert g["score"].view(np.uint32).tolist() == [int(np.float32(r[3]).view(np.uint32)) for r in want], (exe, qi)
    # the automatic kernel (column front end, resident scores) agrees within the rule
    ga = GpuIndex(0)
    ga.load_index_dir(root, "body,title")
    a_h, a_c = ga.search_batch(qb, k, "bm25")
    assert_parity(got_h, got_c, a_h, a_c, strict=False)
    # in-place scoring (no resident scores) takes the same per-field norms
    gp = GpuIndex(0, kernel="warp")
    gp.set_option("resident_scores", 0)
    gp.load_index_dir(root, "body,title")
    p_h, p_c = gp.search_batch(qb, k, "bm25")
    assert p_h.tobytes() == got_h.tobytes() and p_c.tobytes() == got_c.tobytes()
    with pytest.raises(SearchliteGpuError, match="same field list"):
        gi.load_index_dir(root, "body")
    for h in (gi, ga, gp):
        h.close()


def test_full_size_c4_phrase_properties():
    """BASELINE.json configs[3] (phrase = 2-term adjacent, slop 0) at full size: the 10 M-doc C2 corpus with resident
    positions.  Size-independent checks: phrase bitmaps equal a restatement that uses neither postings nor positions
    (the corpus is a pure function token(seed, doc, position), so "a at p and b at p+1" is recomputed from the generator),
    slop 1 is a superset of slop 0, the reversed phrase differs, every hit of a phrase query lies in its bitmap and holds
    both terms, pruned == exhaustive bytes."""
    import torch
    from searchlite_b200 import synth
    dev = "cuda:0"
    spec = synth.CorpusSpec(n_docs=10_000_000, vocab=1_000_000, seed=20260101)
    seg = synth.generate_segment(spec, dev)
    positions = synth.generate_positions(spec, dev)
    pos_off = torch.zeros(seg.post_tfs.shape[0] + 1, dtype=torch.int64, device=dev)
    pos_off[1:] = torch.cumsum(seg.post_tfs.to(torch.int64), 0)
    assert int(pos_off[-1]) == positions.shape[0] == seg.total_tokens
    torch.cuda.empty_cache()
    gi = GpuIndex(0)
    gi.load_segment(seg)
    gi.load_positions(0, seg.term_offsets, pos_off, positions)
    del seg, positions, pos_off
    torch.cuda.empty_cache()
    cdf = synth.zipf_cdf(spec.vocab, spec.zipf_s).to(dev)
    rng = np.random.default_rng(20260105)
    phrases = []
    while len(phrases) < 12:
        d = int(rng.integers(0, spec.n_docs))
        term, valid = synth.token_terms(spec, d, d + 1, cdf, dev)
        t = term[0][valid[0]].cpu().numpy()
        p = int(rng.integers(0, len(t) - 1))
        a, b = int(t[p]), int(t[p + 1])
        if a >= 9 and b >= 9 and a != b:
            phrases.append((a, b))
    n = spec.n_docs
    ids = gi.compile_phrases([[a, b] for a, b in phrases] + [[a, b] for a, b in phrases[:4]] + [[b, a] for a, b in phrases[:4]],
                             [0] * 12 + [1] * 4 + [0] * 4)
    n_check = 4
    want = [torch.zeros(n, dtype=torch.bool, device=dev) for _ in range(n_check)]
    has_both = [torch.zeros(n, dtype=torch.bool, device=dev) for _ in range(n_check)]
    chunk = 1 << 18
    for d0 in range(0, n, chunk):
        d1 = min(n, d0 + chunk)
        term, valid = synth.token_terms(spec, d0, d1, cdf, dev)
        for i in range(n_check):
            a, b = phrases[i]
            want[i][d0:d1] = ((term[:, :-1] == a) & (term[:, 1:] == b) & valid[:, 1:]).any(dim=1)
            has_both[i][d0:d1] = ((term == a) & valid).any(dim=1) & ((term == b) & valid).any(dim=1)
        del term, valid
    unpack = lambda bits: np.unpackbits(bits.view(np.uint8), bitorder="little")[:n].astype(bool)
    for i in range(n_check):
        got = unpack(gi.filter_bitmap(int(ids[i]), 0, n))
        w = want[i].cpu().numpy()
        assert w.any() and np.array_equal(got, w), (phrases[i], int(got.sum()), int(w.sum()))
        sloppy = unpack(gi.filter_bitmap(int(ids[12 + i]), 0, n))
        assert not (got & ~sloppy).any() and not (sloppy & ~has_both[i].cpu().numpy()).any()  # slop 0 within slop 1 within "holds both"
        rev = unpack(gi.filter_bitmap(int(ids[16 + i]), 0, n))
        assert not np.array_equal(rev, got)
    qb = QueryBatch.from_term_lists([[a, b] for a, b in phrases])
    qb.filter_id = np.asarray(ids[:12], dtype=np.int32)
    h, c = gi.search_batch(qb, 11, "bm25")
    hp, cp = gi.search_batch(qb, 11, "bmw")
    assert h.tobytes() == hp.tobytes() and c.tobytes() == cp.tobytes()
    assert c.min() >= 1
    for i in range(n_check):
        docs = torch.from_numpy(h[i, : c[i]]["doc_id"].astype(np.int64)).to(dev)
        assert bool(want[i][docs].all())
        sc = h[i, : c[i]]["score"]
        assert np.all(sc[:-1] >= sc[1:])
    gi.close()


def test_multi_match_best_fields_score_plan(index_dir):
    """MultiMatch best_fields over [body, title] (query/planner.rs:381-404): every query term has one key per field, all
    keys of a field add into that field's leaf (query/wand.rs:470-497), the score is DisMax(tie_breaker) over the field
    leaves (planner.rs:136-151).  Expected: per-key scores from single-field oracles, leaves folded in term order, the
    plan evaluated by the oracle's evaluator; segments merged by (score desc, segment, doc)."""
    root, segs = index_dir
    from oracle import slo
    from searchlite_b200.engine import PLAN_DTYPE, plan_postfix
    gi = GpuIndex(0)
    assert gi.load_index_dir(root, "body,title") == 2
    per_field = {}
    for si, s in enumerate(segs):
        for field in ("body", "title"):
            keys, toff, docs, tfs, poff, pos, lens = sw.csr_of(s, field)
            sd = SegmentData(si, len(s.docs), toff, docs, tfs, lens, int(lens.sum()),
                             deleted_docs=np.asarray(s.deleted, dtype=np.uint32) if s.deleted else None)
            per_field[(si, field)] = (keys, slo.OracleIndex(sd))
    cache = {}

    def key_scores(si, key):
        if (si, key) not in cache:
            keys, o = per_field[(si, key.split(":")[0])]
            if key not in keys:
                cache[(si, key)] = {}
            else:
                h, c = o.search_batch(QueryBatch.from_term_lists([[keys.index(key)]]), len(segs[si].docs) + 1, "bm25")
                cache[(si, key)] = {int(x["doc_id"]): np.float32(x["score"]) for x in h[0, : c[0]]}
        return cache[(si, key)]

    rng = np.random.default_rng(21)
    queries, exprs = [], []
    for qi in range(30):
        toks = rng.choice(np.arange(1, 60), size=int(rng.integers(1, 4)), replace=False)
        queries.append([f"{f}:w{int(t)}" for t in toks for f in ("body", "title")])
        exprs.append(("dismax", [("leaf", 0), ("leaf", 1)], float(rng.choice([0.0, 0.3, 1.0]))))
    qb = QueryBatch.from_term_lists([[gi.term_lookup(k) for k in q] for q in queries])
    qb.terms["leaf"] = np.concatenate([[0 if k.startswith("body:") else 1 for k in q] for q in queries]).astype(np.uint32)
    qb.set_plans(exprs)
    qb.leaf_count[:] = 2
    qb._structs = None
    k = 11
    L = slo.lib()
    base = None
    for exe in ("bm25", "bmw"):
        got_h, got_c = gi.search_batch(qb, k, exe)
        if base is None:
            base = (got_h.tobytes(), got_c.tobytes())
        else:
            assert (got_h.tobytes(), got_c.tobytes()) == base
        for qi, q in enumerate(queries):
            nodes = np.array(plan_postfix(exprs[qi]), dtype=PLAN_DTYPE)
            rows = []
            for si in range(len(segs)):
                leaves = {}
                for key in q:
                    leaf = 0 if key.startswith("body:") else 1
                    for d, sc in key_scores(si, key).items():
                        buf = leaves.setdefault(d, np.zeros(2, dtype=np.float32))
                        buf[leaf] = np.float32(buf[leaf] + sc)
                for d, buf in leaves.items():
                    sc = np.float32(L.slo_plan_evaluate(nodes.ctypes.data, len(nodes), buf.ctypes.data, 2))
                    rows.append((-float(sc), si, d, sc))
            rows.sort(key=lambda r: (r[0], r[1], r[2]))
            want = rows[:k]
            g = got_h[qi, : got_c[qi]]
            assert [(int(x["segment_ord"]), int(x["doc_id"])) for x in g] == [(r[1], r[2]) for r in want], (exe, qi, q)
            assert g["score"].view(np.uint32).tolist() == [int(np.float32(r[3]).view(np.uint32)) for r in want], (exe, qi)
    gi.close()
