"""GPU parity: CUDA engine (through the C ABI) vs the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from searchlite_b200 import GpuIndex, QueryBatch, synth
from tests.parity import assert_parity

pytestmark = pytest.mark.gpu


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


def test_division_sequence_is_ieee_exact():
    gi = GpuIndex(0)
    assert gi.selftest_div(200_000_000, seed=3) == 0
    gi.close()


def test_full_size_c2_properties():
    """BASELINE.json configs[1] at full size (10 M docs, 1 M-term vocabulary, 4096 queries, k = 11):
    size-independent properties — every kernel variant and every execution strategy returns the
    same bytes (idempotence across code paths), lists are ordered (score desc, doc asc) with unique
    docs, counts are full — plus oracle parity on a bounded sample of the batch."""
    import torch
    spec = synth.CorpusSpec(n_docs=10_000_000, vocab=1_000_000, seed=20260101)
    seg = synth.generate_segment(spec, "cuda:0")
    qb = synth.generate_queries(4096, spec.vocab, seed=20260102)
    k = 11
    results = {}
    for kernel, mode in (("auto", "bm25"), ("auto", "bmw"), ("warp-inplace", "bm25"), ("cta", "bm25")):
        gi = GpuIndex(0, kernel=kernel)
        gi.load_segment(seg)
        p = gi.prepare(qb, k, mode)
        p.run()
        first = p.fetch()
        p.run()
        again = p.fetch()
        assert first[0].tobytes() == again[0].tobytes() and first[1].tobytes() == again[1].tobytes()  # re-runnable
        results[(kernel, mode)] = first
        p.free()
        gi.close()
    base_h, base_c = results[("auto", "bm25")]
    for key, (h, c) in results.items():
        assert h.tobytes() == base_h.tobytes() and c.tobytes() == base_c.tobytes(), key
    assert np.all(base_c == k)
    sc, dc = base_h["score"], base_h["doc_id"].astype(np.int64)
    assert np.all(sc[:, :-1] >= sc[:, 1:])
    tie = sc[:, :-1] == sc[:, 1:]
    assert np.all(dc[:, :-1][tie] < dc[:, 1:][tie])
    assert all(len(set(row.tolist())) == k for row in dc[::64])
    assert np.all(dc < spec.n_docs) and np.all(base_h["segment_ord"] == 0)
    # oracle on a sample
    host = seg.to_host()
    del seg
    torch.cuda.empty_cache()
    from oracle import slo
    ora = slo.OracleIndex(host)
    n = 48
    ref_h, ref_c = ora.search_batch(qb.subset(0, n), k, "bm25_dense", threads=slo.max_threads())
    assert_parity(ref_h, ref_c, base_h[:n], base_c[:n], strict=True)
