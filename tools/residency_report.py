import sys, json
sys.path.insert(0, "/root/repo")
from searchlite_b200 import GpuIndex, synth
seg = synth.generate_segment(synth.CorpusSpec(n_docs=10_000_000, vocab=1_000_000, seed=20260101), "cuda:0")
gi = GpuIndex(0); gi.load_segment(seg)
r = gi.segment_residency(0)
print(json.dumps(r)); print(sum(v for k, v in r.items() if not k.startswith("n_")), gi.counters()["resident_bytes"])
