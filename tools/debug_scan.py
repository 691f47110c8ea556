import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
from searchlite_b200 import GpuIndex, synth
from oracle import slo
from tests.helpers import canonical_batch

spec = synth.CorpusSpec(n_docs=50_000, vocab=8_000, seed=11, len_lo=20, len_hi=80)
seg = synth.generate_segment(spec, "cpu", chunk_docs=8192)
qb = synth.generate_queries(300, spec.vocab, seed=12)
ora = slo.OracleIndex(seg)
dbg = int(os.environ.get("DBG", "0"))
gi = GpuIndex(0, kernel="items", sub_docs=2048, options={"dense_min_df": 64, "dense_den": 64, "dbg": dbg})
gi.load_segment(seg)
cb = canonical_batch(gi, qb)
for k in (1, 11):
    ref_h, ref_c = ora.search_batch(cb, k, "bm25")
    bad_runs = 0
    for it in range(40):
        got_h, got_c = gi.search_batch(qb, k, "bm25")
        bad = [q for q in range(qb.n_queries) if got_h[q].tobytes() != ref_h[q].tobytes()]
        if bad and not (dbg & 4):
            bad_runs += 1
            if bad_runs <= 2:
                q = bad[0]
                a, b = int(qb.term_off[q]), int(qb.term_off[q + 1])
                terms = qb.terms["term_id"][a:b]
                print("k", k, "iter", it, "bad", bad[:8], "terms", terms.tolist(), "col", [gi.term_has_column(0, int(t)) for t in terms])
                print("  ref", [(int(h["doc_id"]), round(float(h["score"]), 4)) for h in ref_h[q][: ref_c[q]]][:4], "got", [(int(h["doc_id"]), round(float(h["score"]), 4)) for h in got_h[q][: got_c[q]]][:4])
        if dbg & 4:
            # scan only: every returned hit must be a true hit with the right score; count docs that hold a sparse term and are missing
            pass
    print("DBG", dbg, "k", k, "bad runs", bad_runs, "of 40")
