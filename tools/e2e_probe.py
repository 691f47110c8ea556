import time, torch, sys
sys.path.insert(0, "/root/repo")
from searchlite_b200 import GpuIndex, synth
spec = synth.CorpusSpec(n_docs=10_000_000, vocab=1_000_000, seed=20260101)
seg = synth.generate_segment(spec, "cuda:0")
qb = synth.generate_queries(4096, spec.vocab, seed=20260102)
gi = GpuIndex(0)
gi.load_segment(seg)
del seg; torch.cuda.empty_cache()
for it in range(4):
    t0 = time.perf_counter(); s = qb.structs(); t1 = time.perf_counter()
    p = gi.prepare(qb, 11, "bm25"); t2 = time.perf_counter()
    p.run(sync=True); t3 = time.perf_counter()
    h = p.fetch(); t4 = time.perf_counter()
    p.free(); t5 = time.perf_counter()
    print(f"structs {1e3*(t1-t0):.1f} prepare {1e3*(t2-t1):.1f} run {1e3*(t3-t2):.1f} fetch {1e3*(t4-t3):.1f} free {1e3*(t5-t4):.1f} ms", gi.counters()["last_score_ms"])
for it in range(3):
    t0 = time.perf_counter(); gi.search_batch(qb, 11, "bm25"); t1 = time.perf_counter()
    print(f"search_batch {1e3*(t1-t0):.1f} ms", gi.counters()["last_score_ms"], gi.counters()["last_batch_ms"])
