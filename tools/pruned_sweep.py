"""time the pruned (bmw) execution at C2 for several MaxScore caps"""
import sys, time, torch
sys.path.insert(0, "/root/repo")
from searchlite_b200 import GpuIndex, synth
spec = synth.CorpusSpec(n_docs=10_000_000, vocab=1_000_000, seed=20260101)
seg = synth.generate_segment(spec, "cuda:0")
qb = synth.generate_queries(4096, spec.vocab, seed=20260102)
base = None
for pct in [int(a) for a in sys.argv[1:]] or [0, 20, 40, 60, 100]:
    gi = GpuIndex(0, options={"maxscore_pct": pct, "dense_den": 8})
    gi.load_segment(seg)
    p = gi.prepare(qb, 11, "bmw")
    for _ in range(3):
        p.run(sync=True)
    t0 = time.perf_counter()
    for _ in range(5):
        p.run(sync=True)
    ms = 1e3 * (time.perf_counter() - t0) / 5
    h, c = p.fetch()
    if base is None:
        base = h.tobytes()
    print(f"maxscore_pct {pct:3d}: {ms:7.2f} ms/step  {4096 / ms * 1e3:9.0f} q/s  identical {h.tobytes() == base}", flush=True)
    p.free(); gi.close()
