"""ScorePlans: per-leaf accumulation (query/wand.rs:469-497) and ScoreExpr evaluation
(Sum / DisMax with tie_breaker, query/planner.rs:113-164).

CPU part: the oracle's postfix evaluator against a tree-recursive restatement of planner.rs:133-153 in numpy
f32, the reference's own multi-field ordering tests (searchlite-core/tests/multi_field.rs:105-192) restated on
the literal 5-doc corpus, and bm25 == wand under plans.  GPU part: the warp and CTA-per-item kernels with one
accumulator plane per leaf, bit-exact against the oracle's `bm25` execution."""
import ctypes as C

import numpy as np
import pytest

from searchlite_b200.engine import PLAN_DTYPE, QueryBatch, SegmentData, plan_postfix
from tests.helpers import segment_from_postings, token_corpus
from tests.parity import assert_parity

f32 = np.float32


def _slo():
    from oracle import slo
    slo.build()
    return slo


def tree_eval(expr, leaves):
    """ScoreExpr::evaluate, query/planner.rs:133-153, in f32 (every operation rounded once, no fusion)"""
    if expr[0] == "leaf":
        return f32(leaves[expr[1]]) if expr[1] < len(leaves) else f32(0)
    vals = [tree_eval(c, leaves) for c in expr[1]]
    if expr[0] == "sum":
        s = f32(0)
        for v in vals:
            s = f32(s + v)
        return s
    if not vals:
        return f32(0)
    mx, s = f32(-np.inf), f32(0)
    for v in vals:
        mx = max(mx, v)
        s = f32(s + v)
    return f32(mx + f32(f32(expr[2]) * f32(s - mx)))


def oracle_eval(slo, expr, leaves):
    nodes = np.array(plan_postfix(expr), dtype=PLAN_DTYPE)
    lv = np.asarray(leaves, dtype=np.float32)
    return f32(slo.lib().slo_plan_evaluate(nodes.ctypes.data, len(nodes), lv.ctypes.data, len(lv)))


def random_expr(rng, n_leaves, depth=0):
    """a tree reading every leaf exactly once (what the planner builds, planner.rs:284-460)"""
    def build(leaves, depth):
        if len(leaves) == 1 and (depth > 0 or rng.random() < 0.3):
            return ("leaf", leaves[0])
        n_parts = int(rng.integers(1, min(len(leaves), 3) + 1)) if depth < 2 else len(leaves)
        cuts = sorted(rng.choice(np.arange(1, len(leaves)), size=n_parts - 1, replace=False).tolist()) if n_parts > 1 else []
        parts = [leaves[a:b] for a, b in zip([0] + cuts, cuts + [len(leaves)])]
        kids = [build(p, depth + 1) if len(p) > 1 or depth < 2 else ("leaf", p[0]) for p in parts]
        if rng.random() < 0.5:
            return ("sum", kids)
        return ("dismax", kids, float(rng.choice([0.0, 0.1, 0.25, 0.5, 1.0])))
    return build(list(rng.permutation(n_leaves)), depth)


def test_plan_evaluate_matches_planner_restatement():
    slo = _slo()
    rng = np.random.default_rng(7)
    assert oracle_eval(slo, ("dismax", [], 0.5), [1.0]) == 0.0                       # planner.rs:140-142
    assert oracle_eval(slo, ("leaf", 5), [1.0, 2.0]) == 0.0                          # leaves.get(idx).unwrap_or(0.0)
    assert oracle_eval(slo, ("dismax", [("leaf", 0), ("leaf", 1)], 0.0), [1.5, 2.5]) == 2.5
    assert oracle_eval(slo, ("dismax", [("leaf", 0), ("leaf", 1)], 0.5), [1.0, 3.0]) == 3.5
    for _ in range(300):
        n = int(rng.integers(1, 9))
        leaves = (rng.random(n) * 12).astype(np.float32)
        leaves[rng.random(n) < 0.3] = 0.0
        e = random_expr(rng, n)
        want, got = tree_eval(e, leaves), oracle_eval(slo, e, leaves)
        assert want.view(np.uint32) == got.view(np.uint32), (e, leaves.tolist())


# ---- searchlite-core/tests/multi_field.rs: literal corpus (title, body), k1 0.9, b 0.4 ----
MF_DOCS = [("rust search", "fast"), ("rust", "search"), ("rust", "rust search"), ("boring", "rust"), ("none", "rust fast search")]


def _field_index(slo, field):
    vocab = sorted({w for d in MF_DOCS for w in d[field].split()})
    seg = token_corpus([[vocab.index(w) for w in d[field].split()] for d in MF_DOCS], len(vocab))
    return vocab, slo.OracleIndex(seg, k1=0.9, b=0.4)


def _key_scores(ix, word):
    vocab, o = ix
    if word not in vocab:
        return {}
    h, c = o.search_batch(QueryBatch.from_term_lists([[vocab.index(word)]]), len(MF_DOCS) + 1, "bm25")
    return {int(x["doc_id"]): f32(x["score"]) for x in h[0, : c[0]]}


def _plan_scores(slo, keys, leaf_of, n_leaves, expr):
    """keys: list of (field index, word) in search_segment's term order; leaves[leaf] += score per key"""
    ix = [_field_index(slo, 0), _field_index(slo, 1)]
    leaves = {}
    for (fi, w), leaf in zip(keys, leaf_of):
        for d, s in _key_scores(ix[fi], w).items():
            buf = leaves.setdefault(d, np.zeros(n_leaves, dtype=np.float32))
            buf[leaf] = f32(buf[leaf] + s)
    return {d: oracle_eval(slo, expr, buf) for d, buf in leaves.items()}


def test_reference_dis_max_tie_breaker_prefers_multi_field_hit():
    """multi_field.rs:169-192: DisMax[title:rust, body:rust], tie_breaker 0.5 -> doc-3 first"""
    slo = _slo()
    sc = _plan_scores(slo, [(0, "rust"), (1, "rust")], [0, 1], 2, ("dismax", [("leaf", 0), ("leaf", 1)], 0.5))
    best = sorted(sc.items(), key=lambda kv: (-float(kv[1]), kv[0]))[0][0]
    assert best == 2  # doc-3 (docs are ordered by _id)
    # tie_breaker 0 keeps only the better field: the body-only hit doc-4 scores its body leaf alone
    sc0 = _plan_scores(slo, [(0, "rust"), (1, "rust")], [0, 1], 2, ("dismax", [("leaf", 0), ("leaf", 1)], 0.0))
    assert sc0[3] == _key_scores(_field_index(slo, 1), "rust")[3]


def test_reference_most_fields_scores_above_best_fields():
    """multi_field.rs:105-167: "rust search" over [title, body]; best_fields = DisMax over one leaf per field
    (planner.rs:381-404, tie_breaker None -> 0), most_fields = one leaf for everything; doc-2 scores higher
    under most_fields"""
    slo = _slo()
    keys = [(0, "rust"), (1, "rust"), (0, "search"), (1, "search")]
    best = _plan_scores(slo, keys, [0, 1, 0, 1], 2, ("dismax", [("leaf", 0), ("leaf", 1)], 0.0))
    most = _plan_scores(slo, keys, [0, 0, 0, 0], 1, ("leaf", 0))
    assert 1 in best and 1 in most and most[1] > best[1]
    body_only = _plan_scores(slo, [(1, "rust"), (1, "search")], [0, 0], 1, ("dismax", [("leaf", 0)], 0.0))
    assert 2 in body_only  # doc-3


def _random_case(rng, n_docs=3000, vocab=40, n_queries=48, max_leaves=4, deleted=False):
    doc_tokens = [rng.integers(0, vocab, size=int(rng.integers(3, 30))).tolist() for _ in range(n_docs)]
    seg = token_corpus(doc_tokens, vocab)
    if deleted:
        seg.deleted_docs = np.sort(rng.choice(n_docs, size=n_docs // 10, replace=False)).astype(np.uint32)
    term_lists, leaf_lists, exprs = [], [], []
    for qi in range(n_queries):
        nt = int(rng.integers(1, 7))
        terms = rng.choice(vocab + 2, size=nt, replace=False).tolist()  # ids >= vocab are absent from the term space
        if qi % 7 == 3:  # no plan: the running sum
            leaves, e = list(range(nt)), None
        else:
            nl = int(rng.integers(1, min(nt, max_leaves) + 1))
            leaves = list(range(nl)) + rng.integers(0, nl, size=nt - nl).tolist()
            leaves = [leaves[i] for i in rng.permutation(nt)]
            e = random_expr(rng, nl)
        term_lists.append(terms)
        leaf_lists.append(leaves)
        exprs.append(e)
    return seg, term_lists, leaf_lists, exprs


def _plan_batch(term_lists, leaf_lists, exprs, weights=None):
    qb = QueryBatch.from_term_lists(term_lists, weights)
    qb.terms["leaf"] = np.concatenate([np.asarray(l, dtype=np.uint32) for l in leaf_lists]) if leaf_lists else []
    # leaf_count of a plan query = number of leaves its expression reads
    qb.set_plans(exprs)
    for qi, e in enumerate(exprs):
        if e is not None:
            qb.leaf_count[qi] = 1 + max(n[1] for n in plan_postfix(e) if n[0] == 0)
    qb._structs = None
    return qb


def test_oracle_plan_search_matches_manual_fold():
    slo = _slo()
    rng = np.random.default_rng(3)
    seg, tl, ll, ex = _random_case(rng, n_docs=300, vocab=12, n_queries=12)
    ora = slo.OracleIndex(seg)
    qb = _plan_batch(tl, ll, ex)
    k = seg.doc_count + 1
    hits, counts = ora.search_batch(qb, k, "bm25")
    for qi in range(len(tl)):
        leaves = {}
        nl = int(qb.leaf_count[qi]) if ex[qi] is not None else len(tl[qi])
        for t, leaf in zip(tl[qi], ll[qi]):
            if t >= 12:
                continue
            h, c = ora.search_batch(QueryBatch.from_term_lists([[t]]), k, "bm25")
            for x in h[0, : c[0]]:
                buf = leaves.setdefault(int(x["doc_id"]), np.zeros(nl, dtype=np.float32))
                buf[leaf] = f32(buf[leaf] + f32(x["score"]))
        e = ex[qi] if ex[qi] is not None else ("sum", [("leaf", i) for i in range(nl)])
        want = sorted(((-float(tree_eval(e, b)), d, tree_eval(e, b)) for d, b in leaves.items()))
        got = hits[qi, : counts[qi]]
        assert [int(x) for x in got["doc_id"]] == [w[1] for w in want], qi
        assert got["score"].view(np.uint32).tolist() == [int(w[2].view(np.uint32)) for w in want], qi


def test_oracle_wand_agrees_with_bm25_under_plans():
    """tests/pruning.rs:45-104 property with a ScorePlan attached: same ids, |dscore| < 1e-5 relative"""
    slo = _slo()
    rng = np.random.default_rng(11)
    seg, tl, ll, ex = _random_case(rng, n_docs=800, vocab=20, n_queries=24)
    ora = slo.OracleIndex(seg)
    qb = _plan_batch(tl, ll, ex)
    base = ora.search_batch(qb, 11, "bm25")
    got = ora.search_batch(qb, 11, "wand")
    assert_parity(*base, *got, strict=False)


# ---- GPU ------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["auto", "cta", "warp", "warp-inplace"])  # auto: warp kernel for k <= 32, CTA kernel above
@pytest.mark.parametrize("k", [11, 101])
@pytest.mark.parametrize("deleted", [False, True])
def test_gpu_plans_bit_exact(k, deleted, kernel):
    from searchlite_b200 import GpuIndex
    if k > 32 and kernel.startswith("warp"):
        pytest.skip("the warp kernel handles k <= 32")
    slo = _slo()
    rng = np.random.default_rng(100 + k + int(deleted))
    seg, tl, ll, ex = _random_case(rng, n_docs=6000, vocab=40, n_queries=64, max_leaves=6, deleted=deleted)
    weights = [[float(rng.choice([0.5, 1.0, 2.0, 3.5])) for _ in t] for t in tl]
    qb = _plan_batch(tl, ll, ex, weights)
    ora = slo.OracleIndex(seg)
    ref = ora.search_batch(qb, k, "bm25")
    gi = GpuIndex(0, kernel=kernel)
    gi.load_segment(seg)
    got = gi.search_batch(qb, k, "bm25")
    assert_parity(*ref, *got, strict=True)
    for exe in ("wand", "bmw"):  # pruning stays exact: a plan never exceeds the sum of its terms' bounds
        h, c = gi.search_batch(qb, k, exe)
        assert h.tobytes() == got[0].tobytes() and c.tobytes() == got[1].tobytes(), exe
    gi.close()
    gt = GpuIndex(0, tile_docs=1024, kernel=kernel)  # many tiles per query: plans across tile boundaries
    gt.load_segment(seg)
    h, c = gt.search_batch(qb, k, "bm25")
    assert h.tobytes() == got[0].tobytes() and c.tobytes() == got[1].tobytes()
    gt.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["auto", "cta", "warp"])
def test_gpu_plans_with_bool_matcher_and_segments(kernel):
    from searchlite_b200 import GpuIndex
    slo = _slo()
    rng = np.random.default_rng(5)
    vocab = 30
    segs = []
    for so in range(2):
        toks = [rng.integers(0, vocab, size=int(rng.integers(3, 25))).tolist() for _ in range(2500)]
        segs.append(token_corpus(toks, vocab, segment_ord=so))
    queries, exprs = [], []
    for qi in range(40):
        t = rng.choice(vocab, size=5, replace=False).tolist()
        queries.append({"must": t[:1], "should": t[1:4], "must_not": t[4:], "min_should": int(rng.integers(0, 3))})
        # leaves follow from_bool: must 0, should 1..3; Sum(must, DisMax(should...))
        exprs.append(("sum", [("leaf", 0), ("dismax", [("leaf", 1), ("leaf", 2), ("leaf", 3)], 0.25)]) if qi % 5 else None)
    qb = QueryBatch.from_bool(queries).set_plans(exprs)
    k = 21
    gi = GpuIndex(0, kernel=kernel)
    for s in segs:
        gi.load_segment(s)
    got = gi.search_batch(qb, k, "bm25")
    per_seg = [slo.OracleIndex(s).search_batch(qb, k, "bm25") for s in segs]
    want_h = np.zeros_like(got[0])
    want_c = np.zeros_like(got[1])
    for qi in range(qb.n_queries):
        m = slo.merge_hits([h[qi, : c[qi]] for h, c in per_seg], k)
        want_h[qi, : len(m)] = m
        want_c[qi] = len(m)
    assert_parity(want_h, want_c, *got, strict=True)
    h, c = gi.search_batch(qb, k, "bmw")
    assert h.tobytes() == got[0].tobytes() and c.tobytes() == got[1].tobytes()
    gi.close()


@pytest.mark.gpu
def test_gpu_plan_validation():
    from searchlite_b200 import GpuIndex, SearchliteGpuError
    seg = segment_from_postings([([0, 1, 2], [1, 2, 1]), ([1, 3], [1, 1])], [5, 6, 7, 8])
    gi = GpuIndex(0)
    gi.load_segment(seg)

    def run(expr, leaves=(0, 1), fix=None):
        qb = QueryBatch.from_term_lists([[0, 1]])
        qb.terms["leaf"] = np.asarray(leaves, dtype=np.uint32)
        qb.set_plans([expr])
        if fix:
            fix(qb)
        return gi.search_batch(qb, 3, "bm25")

    h, c = run(("dismax", [("leaf", 0), ("leaf", 1)], 1.0))
    assert c[0] == 3
    with pytest.raises(SearchliteGpuError, match="tie_breaker"):
        run(("dismax", [("leaf", 0), ("leaf", 1)], 1.5))
    with pytest.raises(SearchliteGpuError, match="leaf"):
        run(("sum", [("leaf", 0), ("leaf", 3)]))
    with pytest.raises(SearchliteGpuError, match="leaf 1 but the plan has 1"):
        run(("leaf", 0), fix=lambda qb: qb.leaf_count.__setitem__(0, 1))

    def two_values(qb):
        qb.plan_nodes = np.array([(0, 0, 0.0), (0, 1, 0.0)], dtype=PLAN_DTYPE)
        qb.plan_off = np.array([0, 2], dtype=np.int64)
    with pytest.raises(SearchliteGpuError, match="instead of one"):
        run(("leaf", 0), fix=two_values)
    gi.close()
    gw = GpuIndex(0, kernel="reg")
    gw.load_segment(seg)
    qb = QueryBatch.from_term_lists([[0, 1]]).set_plans([("sum", [("leaf", 0), ("leaf", 1)])])
    with pytest.raises(SearchliteGpuError, match="items kernel handles plain OR"):
        gw.search_batch(qb, 3, "bm25")
    gw.close()
