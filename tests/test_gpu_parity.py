"""GPU parity: CUDA engine (through the C ABI) vs the CPU oracle on the same seeded inputs.


Float contract (include/searchlite_gpu.h): the CTA and warp kernels sum a doc's contributions in query
term order — bit-identical to the oracle's `bm25` mode (brute_force, query/wand.rs:527-548); the
items kernel (the automatic choice for plain OR queries) sums the terms without a dense column first,
then the terms with one — bit-identical to the oracle on that permutation of the query and within the
north-star 1e-5 rule of the query order."""
import numpy as np
import pytest

from searchlite_b200 import GpuIndex, QueryBatch, SearchliteGpuError, synth
from tests.helpers import assert_engine_parity
from tests.parity import assert_parity

pytestmark = pytest.mark.gpu

# small corpora: give every term with df >= N/8 (and >= 64 postings) a column so the column path runs
DENSE = {"dense_min_df": 64}


def _oracle(seg):
    from oracle import slo
    return slo.OracleIndex(seg)


@pytest.fixture(scope="module")
def small():
    spec = synth.CorpusSpec(n_docs=50_000, vocab=8_000, seed=11, len_lo=20, len_hi=80)
    seg = synth.generate_segment(spec, "cpu", chunk_docs=8192)
    qb = synth.generate_queries(300, spec.vocab, seed=12)
    return seg, qb


@pytest.mark.parametrize("kernel,tile_docs,sub_docs,k", [
    ("cta", 1024, 0, 11), ("cta", 16384, 0, 11), ("cta", 1024, 0, 101), ("cta", 16384, 0, 101),
    ("warp", 0, 128, 11), ("warp", 0, 2048, 11), ("warp", 0, 1024, 1), ("warp", 0, 2048, 32),
    ("warp-inplace", 0, 256, 11), ("warp-inplace", 0, 2048, 11),
])
def test_bm25_bit_exact_vs_oracle(small, kernel, tile_docs, sub_docs, k):
    seg, qb = small
    ora = _oracle(seg)
    ref_h, ref_c = ora.search_batch(qb, k, "bm25")
    gi = GpuIndex(0, tile_docs=tile_docs, sub_docs=sub_docs, kernel=kernel)
    gi.load_segment(seg)
    st = gi.segment_stats(0)
    assert st["avgdl"] == ora.avgdl and st["live_docs"] == ora.live_docs and st["min_doc_len"] == ora.min_doc_len
    got_h, got_c = gi.search_batch(qb, k, "bm25")
    assert_parity(ref_h, ref_c, got_h, got_c, strict=True)
    gi.close()


@pytest.mark.parametrize("k", [1, 11, 32])
@pytest.mark.parametrize("dense_den", [0, 8, 64])
@pytest.mark.parametrize("sub_docs", [128, 256, 2048, 4096])
def test_items_kernel_vs_oracle(small, k, dense_den, sub_docs):
    """the posting-driven items kernel.  dense_den 0: no columns (query order); 8: default; 64: most query terms
    are columns.  bm25 / wand / bmw return the same bytes, and so does the statistics run (warp kernel on the same
    term layout)."""
    seg, qb = small
    ora = _oracle(seg)
    gi = GpuIndex(0, kernel="items", sub_docs=sub_docs, options={**DENSE, "dense_den": dense_den})
    gi.load_segment(seg)
    n_col = sum(gi.term_has_column(0, int(t)) for t in np.unique(qb.terms["term_id"]))
    assert (n_col == 0) == (dense_den == 0)
    first = None
    for mode in ("bm25", "wand", "bmw"):
        got = gi.search_batch(qb, k, mode)
        assert_engine_parity(gi, ora, qb, k, got, exact_order=(dense_den == 0))
        if first is None:
            first = got
        assert got[0].tobytes() == first[0].tobytes() and got[1].tobytes() == first[1].tobytes(), mode
    with_stats = gi.search_batch(qb, k, "bm25", want_stats=True)
    assert with_stats[0].tobytes() == first[0].tobytes() and with_stats[1].tobytes() == first[1].tobytes()
    gi.close()


@pytest.mark.parametrize("maxscore_pct", [0, 35, 100])
def test_items_kernel_maxscore_settings_are_exact(small, maxscore_pct):
    """the MaxScore cap only moves work between scattering and exact rescoring: the bytes never change"""
    seg, qb = small
    ora = _oracle(seg)
    gi = GpuIndex(0, kernel="items", sub_docs=512, options={**DENSE, "dense_den": 32, "maxscore_pct": maxscore_pct})
    gi.load_segment(seg)
    full = gi.search_batch(qb, 11, "bm25")
    got = gi.search_batch(qb, 11, "bmw")
    assert got[0].tobytes() == full[0].tobytes() and got[1].tobytes() == full[1].tobytes()
    assert_engine_parity(gi, ora, qb, 11, got)
    gi.close()


def test_items_kernel_weights_and_filters(small):
    """weights != 1 (score_tf multiplies by the merged weight, query/wand.rs:284-285), deleted docs and a root
    filter on the items kernel, every execution"""
    seg, qb = small
    import copy
    seg2 = copy.copy(seg)
    seg2.deleted_docs = np.arange(0, seg.doc_count, 7, dtype=np.uint32)
    year = (np.arange(seg.doc_count) % 26 + 2000).astype(np.int64)
    seg2.fast_i64 = {"year": (year, None)}
    rng = np.random.default_rng(5)
    tl = [qb.terms["term_id"][int(qb.term_off[q]):int(qb.term_off[q + 1])].tolist() for q in range(qb.n_queries)]
    wq = QueryBatch.from_term_lists(tl, [[float(np.float32(rng.uniform(0.25, 3.0))) for _ in t] for t in tl])
    ora = _oracle(seg2)
    gi = GpuIndex(0, kernel="items", sub_docs=1024, options=DENSE)
    cols = gi.load_segment(seg2)
    from searchlite_b200.engine import FILTER_DTYPE, F_I64_RANGE
    from tests.helpers import canonical_batch

    def prog(c):
        node = np.zeros(1, dtype=FILTER_DTYPE)
        node[0]["op"], node[0]["column"], node[0]["i_min"], node[0]["i_max"] = F_I64_RANGE, c["year"], 2003, 2011
        return node
    fid = gi.compile_filter(prog(cols))
    wq.filter_id = np.full(wq.n_queries, fid, dtype=np.int32)
    for mode in ("bm25", "bmw", "wand"):
        got = gi.search_batch(wq, 11, mode)
        assert_parity(*ora.search_batch(canonical_batch(gi, wq), 11, "bm25", filter_nodes=prog(ora.columns)), *got, strict=True)
        assert_parity(*ora.search_batch(wq, 11, "bm25", filter_nodes=prog(ora.columns)), *got, strict=False)
    gi.close()


def test_automatic_kernel_is_the_items_kernel(small):
    seg, qb = small
    ora = _oracle(seg)
    ref = ora.search_batch(qb, 11, "bm25")
    # automatic choice for plain OR queries: the items kernel (column terms streamed from their columns)
    gi = GpuIndex(0, options=DENSE)
    gi.load_segment(seg)
    auto = gi.search_batch(qb, 11, "bm25")
    assert_engine_parity(gi, ora, qb, 11, auto)
    assert gi.counters()["last_items"] > 0
    r0 = GpuIndex(0, kernel="items", options=DENSE)
    r0.load_segment(seg)
    a = r0.search_batch(qb, 11, "bm25")
    assert auto[0].tobytes() == a[0].tobytes() and auto[1].tobytes() == a[1].tobytes()
    # the explicit warp kernel keeps the reference's summation order, bit for bit
    warp = GpuIndex(0, kernel="warp", options=DENSE)
    warp.load_segment(seg)
    assert_parity(*ref, *warp.search_batch(qb, 11, "bm25"), strict=True)
    # the column budget caps how many terms get a column; results stay within the contract
    capped = GpuIndex(0, options={**DENSE, "max_column_bytes": 3 * (57344 * 4)})
    capped.load_segment(seg)
    assert sum(capped.term_has_column(0, t) for t in range(200)) == 3
    assert_engine_parity(capped, ora, qb, 11, capped.search_batch(qb, 11, "bm25"))
    # a batch the items kernel cannot take is an error when it is required, a fallback when it is not
    bq = QueryBatch.from_bool([{"must": [3], "must_not": [5]}])
    with pytest.raises(SearchliteGpuError, match="items kernel handles plain OR"):
        r0.search_batch(bq, 11, "bm25")
    gi.search_batch(bq, 11, "bm25")
    # (a pure AND batch — one MUST term per group — rides the posting scan: accepted either way, same hits)
    aq = QueryBatch.from_bool([{"must": [3, 5]}])
    x, y = r0.search_batch(aq, 11, "bm25"), warp.search_batch(aq, 11, "bm25")
    assert x[1].tobytes() == y[1].tobytes() and sorted(x[0][0]["doc_id"][: x[1][0]].tolist()) == sorted(y[0][0]["doc_id"][: y[1][0]].tolist())
    for g in (gi, warp, r0, capped):
        g.close()


@pytest.mark.parametrize("kernel", ["cta", "warp"])
@pytest.mark.parametrize("execution", ["wand", "bmw"])
def test_pruned_modes_are_exact(small, execution, kernel):
    seg, qb = small
    ora = _oracle(seg)
    ref_h, ref_c = ora.search_batch(qb, 11, "bm25")
    wand_h, wand_c = ora.search_batch(qb, 11, "wand")
    gi = GpuIndex(0, tile_docs=2048, sub_docs=256, kernel=kernel)
    gi.load_segment(seg)
    got_h, got_c, stats = gi.search_batch(qb, 11, execution, want_stats=True)
    assert_parity(ref_h, ref_c, got_h, got_c, strict=True)      # same arithmetic order as bm25
    assert_parity(wand_h, wand_c, got_h, got_c, strict=False)   # reference default strategy, 1e-5 rule
    assert stats["blocks_skipped"].sum() > 0
    gi.close()


@pytest.mark.parametrize("execution", ["wand", "bmw"])
def test_items_kernel_pruning_skips_work_and_stays_exact(small, execution):
    seg, qb = small
    ora = _oracle(seg)
    gi = GpuIndex(0, kernel="items", sub_docs=256, options=DENSE)
    gi.load_segment(seg)
    full_h, full_c = gi.search_batch(qb, 11, "bm25")
    full_ctr = gi.counters()
    got_h, got_c = gi.search_batch(qb, 11, execution)
    ctr = gi.counters()
    assert got_h.tobytes() == full_h.tobytes() and got_c.tobytes() == full_c.tobytes()  # pruning never changes the result
    assert_engine_parity(gi, ora, qb, 11, (got_h, got_c))
    wand = ora.search_batch(qb, 11, "wand")
    assert_parity(*wand, got_h, got_c, strict=False)
    # MaxScore drops whole items of non-essential terms once a query's k-th score is known (on a corpus this small most
    # items start before that): never more work than the exhaustive run, and the exhaustive run drops nothing
    assert ctr["last_postings_scattered"] <= full_ctr["last_postings_scattered"]
    assert full_ctr["last_items_dropped"] == 0
    # per-query statistics come from the warp kernel on the same term layout: same bytes, counted skips
    st_h, st_c, st = gi.search_batch(qb, 11, execution, want_stats=True)
    assert st_h.tobytes() == full_h.tobytes() and st_c.tobytes() == full_c.tobytes()
    assert st["blocks_skipped"].sum() > 0
    gi.close()


def test_removed_sweep_kernel_is_refused():
    with pytest.raises(SearchliteGpuError, match="heavy_kernel 1 .* was removed"):
        GpuIndex(0, options={"heavy_kernel": 1})
    with pytest.raises(SearchliteGpuError, match="unknown option"):
        GpuIndex(0, options={"reg_tile_v": 8})


def test_division_sequence_is_ieee_exact():
    gi = GpuIndex(0)
    assert gi.selftest_div(200_000_000, seed=3) == 0
    gi.close()


def test_full_size_c2_properties():
    """BASELINE.json configs[1] at full size (10 M docs, 1 M-term vocabulary, 4096 queries, k = 11):
    size-independent properties — every execution strategy of a kernel returns the same bytes and a
    batch is re-runnable (idempotence), kernels with the same float contract agree bit for bit, kernels
    with different contracts agree within the 1e-5 rule, lists are ordered (score desc, doc asc) with
    unique docs, counts are full — plus oracle parity on a bounded sample of the batch."""
    import torch
    spec = synth.CorpusSpec(n_docs=10_000_000, vocab=1_000_000, seed=20260101)
    seg = synth.generate_segment(spec, "cuda:0")
    qb = synth.generate_queries(4096, spec.vocab, seed=20260102)
    k = 11
    results = {}
    canon = None
    for kernel, mode, opts in (("auto", "bm25", {}), ("auto", "bmw", {}), ("auto", "wand", {}), ("warp", "bm25", {}),
                               ("warp", "bmw", {}), ("warp-inplace", "bm25", {}), ("cta", "bm25", {})):
        gi = GpuIndex(0, kernel=kernel, options=opts)
        gi.load_segment(seg)
        p = gi.prepare(qb, k, mode)
        p.run()
        first = p.fetch()
        p.run()
        again = p.fetch()
        assert first[0].tobytes() == again[0].tobytes() and first[1].tobytes() == again[1].tobytes()  # re-runnable
        results[(kernel, mode, None)] = first
        if kernel == "auto" and canon is None:
            from tests.helpers import canonical_batch
            canon = canonical_batch(gi, qb.subset(0, 48))
            assert any(gi.term_has_column(0, int(t)) for t in qb.terms["term_id"][:200])
        p.free()
        gi.close()
    base_h, base_c = results[("auto", "bm25", None)]               # column order (automatic choice)
    w_h, w_c = results[("warp", "bm25", None)]                     # query order
    assert results[("auto", "bmw", None)][0].tobytes() == base_h.tobytes()
    assert results[("warp", "bmw", None)][0].tobytes() == w_h.tobytes()
    assert results[("auto", "wand", None)][0].tobytes() == base_h.tobytes() and results[("auto", "wand", None)][1].tobytes() == base_c.tobytes()
    for key in (("warp-inplace", "bm25", None), ("cta", "bm25", None)):
        assert results[key][0].tobytes() == w_h.tobytes() and results[key][1].tobytes() == w_c.tobytes(), key
    assert_parity(w_h, w_c, base_h, base_c, strict=False)
    assert np.all(base_c == k)
    for h in (base_h, w_h):
        sc, dc = h["score"], h["doc_id"].astype(np.int64)
        assert np.all(sc[:, :-1] >= sc[:, 1:])
        tie = sc[:, :-1] == sc[:, 1:]
        assert np.all(dc[:, :-1][tie] < dc[:, 1:][tie])
        assert all(len(set(row.tolist())) == k for row in dc[::64])
        assert np.all(dc < spec.n_docs) and np.all(h["segment_ord"] == 0)
    # oracle on a sample
    host = seg.to_host()
    del seg
    torch.cuda.empty_cache()
    from oracle import slo
    ora = slo.OracleIndex(host)
    n = 48
    ref_h, ref_c = ora.search_batch(qb.subset(0, n), k, "bm25_dense", threads=slo.max_threads())
    assert_parity(ref_h, ref_c, w_h[:n], w_c[:n], strict=True)
    assert_parity(ref_h, ref_c, base_h[:n], base_c[:n], strict=False)
    can_h, can_c = ora.search_batch(canon, k, "bm25_dense", threads=slo.max_threads())
    assert_parity(can_h, can_c, base_h[:n], base_c[:n], strict=True)


@pytest.mark.parametrize("k,dense_den", [(33, 8), (101, 8), (101, 0), (1001, 8), (2048, 64)])
def test_scan_kernel_large_k(small, k, dense_den):
    """k above the warp's sorted top-k (32): candidate pools + radix select on the flat posting scan.  bm25 / wand / bmw
    return the same bytes and match the oracle on the declared term order."""
    seg, qb = small
    ora = _oracle(seg)
    gi = GpuIndex(0, kernel="items", options={**DENSE, "dense_den": dense_den})
    gi.load_segment(seg)
    sub = qb.subset(0, 60)
    first = None
    for mode in ("bm25", "wand", "bmw"):
        got = gi.search_batch(sub, k, mode)
        assert_engine_parity(gi, ora, sub, k, got, exact_order=(dense_den == 0))
        if first is None:
            first = got
        assert got[0].tobytes() == first[0].tobytes() and got[1].tobytes() == first[1].tobytes(), mode
    c = gi.counters()
    assert c["last_items"] > 0  # the scan ran (not the CTA-per-item kernel)
    gi.close()
