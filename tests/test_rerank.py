"""Hybrid rerank: compute_hybrid_score for several clauses, the device pipeline
(slg_rerank_batch: BM25 top-k on the device -> exact similarities -> hybrid order -> segment / shard merge) and the
sequential summation order of metric_similarity.

Reference: searchlite-core/src/vectors/mod.rs:63-129, src/api/reader.rs:218-254 (compute_hybrid_score), :2379-2537
(collect_vector_maps / merge_vector_hits), tests/vector_search.rs:200-270 (hybrid_blends_text_and_vector).
"""
from __future__ import annotations

import numpy as np
import pytest

from searchlite_b200 import synth
from searchlite_b200.engine import HIT_DTYPE

gpu = pytest.mark.gpu


def _f32(x):
    return np.float32(x)


def hybrid_numpy(bm25, clause_scores, alphas, metrics):
    """api/reader.rs:226-254 restated with numpy f32 scalars; clause_scores[c] is None for a missing vector"""
    blended_sum, vector_sum, has = _f32(0), _f32(0), False
    for vs, a, m in zip(clause_scores, alphas, metrics):
        if vs is not None:
            vector_sum = _f32(vector_sum + _f32(vs))
            has = True
        v = _f32(vs) if vs is not None else (_f32(-1.0) if m == "cosine" else np.finfo(np.float32).min)
        if a >= 1.0:
            bl = _f32(bm25)
        elif a <= 0.0:
            bl = v
        else:
            bl = _f32(_f32(_f32(a) * _f32(bm25)) + _f32(_f32(_f32(1.0) - _f32(a)) * v))
        blended_sum = _f32(blended_sum + bl)
    return _f32(blended_sum / _f32(max(len(alphas), 1))), (vector_sum if has else None)


def test_oracle_multi_clause_hybrid_matches_numpy_restatement():
    from oracle import slo
    L = slo.lib()
    rng = np.random.default_rng(5)
    for _ in range(400):
        n = int(rng.integers(1, 9))
        has = rng.random(n) < 0.8
        vs = rng.standard_normal(n).astype(np.float32)
        alpha = rng.choice([0.0, 0.2, 0.5, 0.75, 1.0], n).astype(np.float32)
        metric = rng.integers(0, 2, n).astype(np.int32)
        bm25 = np.float32(rng.random() * 20)
        hv = np.ascontiguousarray(has.astype(np.int32))
        out_sum = np.zeros(1, np.float32)
        out_has = np.zeros(1, np.int32)
        got = L.slo_hybrid_score_clauses(float(bm25), n, hv.ctypes.data, vs.ctypes.data, alpha.ctypes.data, metric.ctypes.data,
                                         out_sum.ctypes.data, out_has.ctypes.data)
        exp, exp_sum = hybrid_numpy(bm25, [float(v) if h else None for v, h in zip(vs, has)], alpha.tolist(),
                                    ["cosine" if m == 0 else "l2" for m in metric])
        assert np.float32(got).tobytes() == np.float32(exp).tobytes()
        assert bool(out_has[0]) == (exp_sum is not None)
        if exp_sum is not None:
            assert out_sum[0].tobytes() == np.float32(exp_sum).tobytes()
        # one clause == the single-clause function
        if n == 1:
            one = L.slo_hybrid_score(float(bm25), int(has[0]), float(vs[0]), float(alpha[0]), int(metric[0]))
            assert np.float32(one).tobytes() == np.float32(got).tobytes()


def test_oracle_round_bf16_is_round_to_nearest_even():
    import torch
    from oracle import slo
    rng = np.random.default_rng(6)
    v = np.concatenate([rng.standard_normal(5000).astype(np.float32), np.array([0.0, -0.0, 1.0, 1.00390625, 1.01171875, 3.4e38, 1e-40], np.float32)])
    exp = torch.from_numpy(v).to(torch.bfloat16).to(torch.float32).numpy()
    assert slo.round_bf16(v).tobytes() == exp.tobytes()


def test_reference_hybrid_blends_text_and_vector_literal():
    """tests/vector_search.rs:200-270 through the oracle: 'long' (rust rust rust, embedding [0,1]) wins on BM25, 'short'
    (rust, [1,0]) wins once the vector [1,0] is blended with alpha 0.2"""
    from oracle import slo
    # doc ids follow the _id order the writer sorts by: "long" = 0, "short" = 1
    from tests.helpers import or_queries, segment_from_postings
    seg = segment_from_postings([([0, 1], [3, 1])], [3, 1])
    ora = slo.OracleIndex(seg)
    qb = or_queries([[0]])
    h, c = ora.search_batch(qb, 3, "bm25")
    assert c[0] == 2 and h[0, 0]["doc_id"] == 0
    offsets = np.array([0, 1], np.uint32)
    rows = np.array([[0.0, 1.0], [1.0, 0.0]], np.float32)
    qv = np.array([[1.0, 0.0]], np.float32)
    out, oc, vs = slo.rerank_batch(h, c, [(0, offsets, rows)], [(qv, 0.2, 1.0, "cosine")])
    assert oc[0] == 2 and out[0, 0]["doc_id"] == 1 and out[0, 0]["score"] > out[0, 1]["score"]
    assert vs[0, 0] == 1.0 and vs[0, 1] == 0.0
    same, _, _ = slo.rerank_batch(h, c, [(0, offsets, rows)], [(qv, 1.0, 1.0, "cosine")])
    assert same[0, :2].tobytes() == h[0, :2].tobytes()


# ---------------------------------------------------------------------------------------------------------------------
def _store(rng, n_docs, dim, metric, frac=0.9):
    from oracle import slo
    L = slo.lib()
    have = rng.random(n_docs) < frac
    offsets = np.full(n_docs, 0xFFFFFFFF, dtype=np.uint32)
    offsets[have] = rng.permutation(int(have.sum())).astype(np.uint32)
    vecs = rng.standard_normal((int(have.sum()), dim)).astype(np.float32)
    if metric == "cosine":
        for row in vecs:
            L.slo_normalize_in_place(row.ctypes.data, dim)
    return offsets, vecs


def _queries(rng, nq, dim, metric):
    from oracle import slo
    L = slo.lib()
    qv = rng.standard_normal((nq, dim)).astype(np.float32)
    if metric == "cosine":
        for row in qv:
            L.slo_normalize_in_place(row.ctypes.data, dim)
    return qv


@gpu
@pytest.mark.parametrize("dim,bf16", [(64, False), (768, False), (768, True), (128, True), (2, False), (20, True), (96, True)])
@pytest.mark.parametrize("metric", ["cosine", "l2"])
def test_rerank_host_candidates_bit_exact(dim, bf16, metric):
    """f32 rows: the reference's sequential fold, bit for bit.  bf16 rows: the same arithmetic on the once-rounded rows"""
    from oracle import slo
    from tests.helpers import gpu_index
    spec = synth.CorpusSpec(n_docs=6_000, vocab=700, seed=181, len_lo=10, len_hi=40)
    seg = synth.generate_segment(spec, "cpu")
    rng = np.random.default_rng(dim * 3 + bf16)
    offsets, vecs = _store(rng, spec.n_docs, dim, metric)
    nq, k = 21, 70
    qv = _queries(rng, nq, dim, metric)
    gi = gpu_index(seg)
    gi.load_vectors(0, offsets, vecs, store_bf16=bf16)
    qb = synth.generate_queries(nq, spec.vocab, seed=182, min_rank=2)
    cands, cc = gi.search_batch(qb, k, "bm25")
    out, vs = gi.rerank(qv, cands, cc, 0.5, metric)
    rows = slo.round_bf16(vecs) if bf16 else vecs
    exp_h, exp_c, exp_vs = slo.rerank_batch(cands, cc, [(0, offsets, rows)], [(qv, 0.5, 1.0, metric)])
    assert out.tobytes() == exp_h.tobytes()
    assert vs.tobytes() == exp_vs.tobytes()
    if bf16:  # the storage deviation against f32 rows stays inside the documented 2e-2
        f32_h, _, _ = slo.rerank_batch(cands, cc, [(0, offsets, vecs)], [(qv, 0.5, 1.0, metric)])
        for q in range(nq):
            a = {int(h["doc_id"]): float(h["score"]) for h in out[q, : cc[q]]}
            for h in f32_h[q, : cc[q]]:
                e = float(h["score"])
                assert abs(a[int(h["doc_id"])] - e) <= 2e-2 * max(1.0, abs(e))
    gi.close()


@gpu
@pytest.mark.parametrize("n_clauses", [2, 3, 5, 8])
def test_rerank_multi_clause_matches_oracle(n_clauses):
    from oracle import slo
    from tests.helpers import gpu_index
    spec = synth.CorpusSpec(n_docs=5_000, vocab=600, seed=191, len_lo=10, len_hi=40)
    seg = synth.generate_segment(spec, "cpu")
    rng = np.random.default_rng(n_clauses)
    dim, nq, k = 64, 17, 60
    offsets, vecs = _store(rng, spec.n_docs, dim, "cosine", frac=0.7)
    gi = gpu_index(seg)
    gi.load_vectors(0, offsets, vecs)
    qb = synth.generate_queries(nq, spec.vocab, seed=192, min_rank=2)
    cands, cc = gi.search_batch(qb, k, "bm25")
    alphas = [0.0, 0.3, 1.0, 0.5, 0.9, 0.1, 0.7, 0.25]
    clauses = [(_queries(rng, nq, dim, "cosine" if c % 2 == 0 else "l2"), alphas[c], 1.0 + 0.5 * c, "cosine" if c % 2 == 0 else "l2")
               for c in range(n_clauses)]
    out, oc, vs = gi.rerank_clauses(clauses, cands, cc)
    exp_h, exp_c, exp_vs = slo.rerank_batch(cands, cc, [(0, offsets, vecs)], clauses)
    assert oc.tobytes() == exp_c.tobytes()
    assert out.tobytes() == exp_h.tobytes()
    assert vs.tobytes() == exp_vs.tobytes()
    # an all-vector plan drops the candidates without a vector (api/reader.rs:2474-2476)
    vclauses = [(c[0], 0.0, c[2], c[3]) for c in clauses]
    out, oc, vs = gi.rerank_clauses(vclauses, cands, cc)
    exp_h, exp_c, exp_vs = slo.rerank_batch(cands, cc, [(0, offsets, vecs)], vclauses)
    assert (oc < cc).any() and oc.tobytes() == exp_c.tobytes() and out.tobytes() == exp_h.tobytes()
    gi.close()


@gpu
@pytest.mark.parametrize("execution", ["bm25", "bmw"])
@pytest.mark.parametrize("k", [11, 101, 1001])
def test_rerank_batch_pipeline_two_segments(execution, k):
    """device pipeline on a two-segment handle: per-segment top-k -> hybrid -> merge, against the oracle doing the same"""
    from oracle import slo
    from tests.helpers import gpu_index
    rng = np.random.default_rng(k)
    dim, nq = 64, 24
    segs, stores, oras = [], [], []
    for ord_, n_docs in enumerate([7_000, 9_000]):
        spec = synth.CorpusSpec(n_docs=n_docs, vocab=500, seed=200 + ord_, len_lo=10, len_hi=50, segment_ord=ord_)
        seg = synth.generate_segment(spec, "cpu")
        segs.append(seg)
        stores.append((ord_,) + _store(rng, n_docs, dim, "cosine"))
        oras.append(slo.OracleIndex(seg))
    gi = gpu_index(segs[0])
    gi.load_segment(segs[1])
    for ord_, offs, vecs in stores:
        gi.load_vectors(ord_, offs, vecs)
    qb = synth.generate_queries(nq, 500, seed=203, min_rank=2)
    qv = _queries(rng, nq, dim, "cosine")
    clauses = [(qv, 0.4, 2.0, "cosine")]
    p = gi.prepare(qb, k, execution)
    p.run(sync=True)
    p.rerank(clauses)
    got_h, got_c = p.fetch()
    got_vs = p.fetch_vector_scores()
    # oracle: per-segment BM25 top-k, each list rescored, lists merged by the hybrid key
    from tests.helpers import canonical_batch
    lists, vss = [], []
    for ora in oras:
        h, c = ora.search_batch(canonical_batch(gi, qb, len(lists)), k, "bm25")
        rh, rc, rv = slo.rerank_batch(h, c, stores, clauses)
        lists.append((rh, rc))
        vss.append(rv)
    for q in range(nq):
        m = slo.merge_hits([h[q, : c[q]] for h, c in lists], k)
        assert got_c[q] == len(m)
        assert got_h[q, : len(m)].tobytes() == m.tobytes(), q
        look = {}
        for (h, c), v in zip(lists, vss):
            for i in range(c[q]):
                look[(int(h[q, i]["segment_ord"]), int(h[q, i]["doc_id"]))] = v[q, i]
        for i in range(len(m)):
            assert got_vs[q, i] == look[(int(m[i]["segment_ord"]), int(m[i]["doc_id"]))]
    # a second run of the batch starts again from BM25 scores
    p.run(sync=True)
    again_h, _ = p.fetch()
    p.rerank(clauses)
    rer_h, _ = p.fetch()
    assert rer_h.tobytes() == got_h.tobytes() and again_h.tobytes() != got_h.tobytes()
    p.free()
    gi.close()


@gpu
def test_rerank_batch_rejects_what_the_reference_rejects():
    from searchlite_b200.engine import SearchliteGpuError
    from tests.helpers import gpu_index
    spec = synth.CorpusSpec(n_docs=2_000, vocab=300, seed=211, len_lo=10, len_hi=30)
    seg = synth.generate_segment(spec, "cpu")
    rng = np.random.default_rng(1)
    offsets, vecs = _store(rng, spec.n_docs, 32, "cosine")
    gi = gpu_index(seg)
    gi.load_vectors(0, offsets, vecs)
    qb = synth.generate_queries(4, spec.vocab, seed=212, min_rank=2)
    p = gi.prepare(qb, 11, "bm25")
    qv = _queries(rng, 4, 32, "cosine")
    with pytest.raises(SearchliteGpuError, match="follows slg_batch_run"):
        p.rerank([(qv, 0.5, 1.0, "cosine")])
    p.run(sync=True)
    with pytest.raises(SearchliteGpuError, match="alpha must be a finite value between 0 and 1"):
        p.rerank([(qv, 1.5, 1.0, "cosine")])
    with pytest.raises(SearchliteGpuError, match="boost must be finite and non-negative"):
        p.rerank([(qv, 0.5, -1.0, "cosine")])
    with pytest.raises(SearchliteGpuError, match="too many vector clauses: got 9, max supported 8"):
        p.rerank([(qv, 0.5, 1.0, "cosine")] * 9)
    with pytest.raises(SearchliteGpuError, match="expects dimension 32, got 16"):
        p.rerank([(qv[:, :16], 0.5, 1.0, "cosine")])
    p.free()
    # offsets that point past the store are refused at load
    bad = offsets.copy()
    bad[np.argmax(bad != 0xFFFFFFFF)] = len(vecs) + 3
    with pytest.raises(SearchliteGpuError, match="point past"):
        gi.load_vectors(0, bad, vecs)
    gi.close()


@gpu
def test_merge_kernel_many_sorted_lists():
    """slg_merge_gathered_packed / _hybrid: 8 shards x k = 1001 (C3 / C5 exchange shape) against the oracle's merge"""
    import torch
    from oracle import slo
    from tests.helpers import gpu_index
    spec = synth.CorpusSpec(n_docs=1_000, vocab=100, seed=221, len_lo=5, len_hi=10)
    gi = gpu_index(synth.generate_segment(spec, "cpu"))
    rng = np.random.default_rng(9)
    S, Q, k = 8, 6, 1001
    blocks, lists = [], []
    for s in range(S):
        h = np.zeros((Q, k), dtype=HIT_DTYPE)
        c = rng.integers(0, k + 1, Q).astype(np.uint32)
        c[0] = k
        vs = np.zeros((Q, k), dtype=np.float32)
        for q in range(Q):
            n = int(c[q])
            sc = np.sort(rng.choice(np.linspace(-2, 9, 300).astype(np.float32), n))[::-1]  # many ties, negatives too
            docs = rng.choice(50_000, n, replace=False).astype(np.uint32)
            rows = np.zeros(n, dtype=HIT_DTYPE)
            rows["segment_ord"], rows["doc_id"], rows["score"] = s, docs, sc
            rows = slo.merge_hits([rows], n) if n else rows
            h[q, :n] = rows
            h[q, n:]["segment_ord"] = 0xFFFFFFFF
            h[q, n:]["doc_id"] = 0xFFFFFFFF
            vs[q, :n] = (rows["doc_id"] % 997).astype(np.float32)
        lists.append((h, c))
        blocks.append(np.concatenate([h.view(np.uint8).reshape(-1), c.view(np.uint8), vs.view(np.uint8).reshape(-1)]))
    dev = torch.from_numpy(np.stack(blocks)).cuda()
    torch.cuda.synchronize()
    got_h, got_c, got_vs = gi.merge_gathered_hybrid(dev.data_ptr(), S, Q, k)
    plain_h, plain_c = gi.merge_gathered_packed(dev.data_ptr(), S, Q, k, shard_stride=dev.shape[1])
    for q in range(Q):
        m = slo.merge_hits([h[q, : c[q]] for h, c in lists], k)
        assert got_c[q] == len(m) == plain_c[q]
        assert got_h[q, : len(m)].tobytes() == m.tobytes()
        assert plain_h[q, : len(m)].tobytes() == m.tobytes()
        assert np.array_equal(got_vs[q, : len(m)], (m["doc_id"] % 997).astype(np.float32))
    gi.close()


def test_oracle_rerank_batch_matches_a_numpy_restatement():
    """slo_rerank_batch (the C5 checker) against pure numpy-f32 loops: sequential dot / squared distance
    (vectors/mod.rs:98-120), boost (api/reader.rs:2421), compute_hybrid_score (:226-254), the all-vector drop (:2474-2476)
    and the SortKey order (query/sort.rs:80-93)"""
    from oracle import slo
    rng = np.random.default_rng(17)
    nq, stride, dim, n_docs = 6, 23, 12, 200
    offsets = np.full(n_docs, 0xFFFFFFFF, dtype=np.uint32)
    have = rng.random(n_docs) < 0.75
    offsets[have] = rng.permutation(int(have.sum())).astype(np.uint32)
    rows = rng.standard_normal((int(have.sum()), dim)).astype(np.float32)
    cands = np.zeros((nq, stride), dtype=HIT_DTYPE)
    counts = rng.integers(5, stride + 1, nq).astype(np.uint32)
    for q in range(nq):
        cands[q, : counts[q]]["doc_id"] = rng.choice(n_docs, counts[q], replace=False)
        cands[q, : counts[q]]["score"] = np.round(rng.random(counts[q]).astype(np.float32) * 8, 2)  # ties on purpose
    for clauses in ([(rng.standard_normal((nq, dim)).astype(np.float32), 0.5, 1.0, "cosine")],
                    [(rng.standard_normal((nq, dim)).astype(np.float32), 0.3, 2.0, "l2"),
                     (rng.standard_normal((nq, dim)).astype(np.float32), 1.0, 1.0, "cosine"),
                     (rng.standard_normal((nq, dim)).astype(np.float32), 0.0, 0.5, "cosine")],
                    [(rng.standard_normal((nq, dim)).astype(np.float32), 0.0, 1.5, "cosine"),
                     (rng.standard_normal((nq, dim)).astype(np.float32), 0.0, 1.0, "l2")]):
        out, oc, vs = slo.rerank_batch(cands, counts, [(0, offsets, rows)], clauses)
        all_vec = all(c[1] <= 0.0 for c in clauses)
        for q in range(nq):
            exp = []
            for h in cands[q, : counts[q]]:
                d = int(h["doc_id"])
                row = rows[offsets[d]] if offsets[d] != 0xFFFFFFFF else None
                sims = []
                for qv, alpha, boost, metric in clauses:
                    if row is None:
                        sims.append(None)
                        continue
                    acc = _f32(0)
                    for x, y in zip(qv[q], row):
                        if metric == "cosine":
                            acc = _f32(acc + _f32(x * y))
                        else:
                            dd = _f32(x - y)
                            acc = _f32(acc + _f32(dd * dd))
                    s = (_f32(0) if np.isnan(acc) else acc) if metric == "cosine" else _f32(-np.sqrt(acc, dtype=np.float32))
                    sims.append(_f32(s * _f32(boost)))
                score, vsum = hybrid_numpy(h["score"], sims, [c[1] for c in clauses], [c[3] for c in clauses])
                if all_vec and vsum is None:
                    continue
                exp.append((float(score), d, float(vsum) if vsum is not None else 0.0))
            exp.sort(key=lambda t: (-t[0], t[1]))
            assert oc[q] == len(exp)
            got = [(float(h["score"]), int(h["doc_id"]), float(v)) for h, v in zip(out[q, : oc[q]], vs[q, : oc[q]])]
            assert [np.float32(a[0]).tobytes() + np.float32(a[2]).tobytes() for a in got] == \
                   [np.float32(a[0]).tobytes() + np.float32(a[2]).tobytes() for a in exp], q
            assert [a[1] for a in got] == [a[1] for a in exp], q
