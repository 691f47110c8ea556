"""C4 phrase shape (BASELINE.json configs[3]): 2-term adjacent phrases, slop 0, over the 10 M-doc C2 corpus with
resident positions.  Times position residency, phrase -> bitmap compilation and a batch of phrase queries,
and checks every compiled bitmap at FULL size against an engine-independent restatement: the corpus is a pure
function token(seed, doc, position), so "doc holds a at p and b at p+1" is recomputed with torch from the
generator alone (no postings, no positions arrays, no engine code).
Usage: python tools/phrase_bench.py [n_docs] [n_phrases]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from searchlite_b200 import GpuIndex, synth  # noqa: E402
from searchlite_b200.engine import QueryBatch  # noqa: E402

n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
n_phr = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = "cuda:0"
spec = synth.CorpusSpec(n_docs=n_docs, vocab=1_000_000, seed=20260101)

t0 = time.perf_counter()
seg = synth.generate_segment(spec, dev)
positions = synth.generate_positions(spec, dev)
torch.cuda.synchronize()
gen_s = time.perf_counter() - t0
pos_off = torch.zeros(seg.post_tfs.shape[0] + 1, dtype=torch.int64, device=dev)
pos_off[1:] = torch.cumsum(seg.post_tfs.to(torch.int64), 0)
assert int(pos_off[-1]) == positions.shape[0] == seg.total_tokens
torch.cuda.empty_cache()

gi = GpuIndex(0)
t0 = time.perf_counter()
gi.load_segment(seg)
torch.cuda.synchronize()
load_s = time.perf_counter() - t0
t0 = time.perf_counter()
gi.load_positions(0, seg.term_offsets, pos_off, positions)
torch.cuda.synchronize()
pos_s = time.perf_counter() - t0
n_post, n_pos = int(seg.post_docs.shape[0]), int(positions.shape[0])
df = (seg.term_offsets[1:] - seg.term_offsets[:-1]).cpu().numpy()
del seg, positions, pos_off
torch.cuda.empty_cache()
print(f"corpus {n_docs} docs: {n_post} postings, {n_pos} positions; generate {gen_s:.1f} s, load postings {load_s:.2f} s, "
      f"positions -> resident {pos_s:.2f} s ({(n_post * 8 + n_pos * 4) / pos_s / 1e9:.1f} GB/s of position index written)", flush=True)

# phrases that occur: adjacent token pairs sampled from the corpus itself (ranks >= 10, like the C2 queries)
cdf = synth.zipf_cdf(spec.vocab, spec.zipf_s).to(dev)
rng = np.random.default_rng(20260105)
phrases = []
while len(phrases) < n_phr:
    d = int(rng.integers(0, n_docs))
    term, valid = synth.token_terms(spec, d, d + 1, cdf, dev)
    t = term[0][valid[0]].cpu().numpy()
    p = int(rng.integers(0, len(t) - 1))
    a, b = int(t[p]), int(t[p + 1])
    if a >= 9 and b >= 9 and a != b:
        phrases.append((a, b))

torch.cuda.synchronize()
t0 = time.perf_counter()
ids = gi.compile_phrases([[a, b] for a, b in phrases]).tolist()  # one launch for the whole batch
torch.cuda.synchronize()
comp_s = time.perf_counter() - t0
driver_postings = sum(int(min(df[a], df[b])) for a, b in phrases)
print(f"{n_phr} phrases (2 adjacent terms, slop 0) compiled in one batch: {1e3 * comp_s:.2f} ms = {1e3 * comp_s / n_phr:.4f} ms per phrase, "
      f"{driver_postings / comp_s / 1e6:.0f} M driver postings/s", flush=True)
t0 = time.perf_counter()
one = [gi.compile_phrase([a, b], 0) for a, b in phrases[:32]]
torch.cuda.synchronize()
print(f"one call per phrase: {1e3 * (time.perf_counter() - t0) / 32:.3f} ms per phrase", flush=True)
for i, f in enumerate(one):
    assert np.array_equal(gi.filter_bitmap(f, 0, n_docs), gi.filter_bitmap(ids[i], 0, n_docs))
    gi.free_filter(f)

# full-size check of the first phrases against the generator
words = (n_docs + 31) // 32
n_check = min(8, n_phr)
want = [torch.zeros(n_docs, dtype=torch.bool, device=dev) for _ in range(n_check)]
chunk = 1 << 18
for d0 in range(0, n_docs, chunk):
    d1 = min(n_docs, d0 + chunk)
    term, valid = synth.token_terms(spec, d0, d1, cdf, dev)
    for i in range(n_check):
        a, b = phrases[i]
        hit = (term[:, :-1] == a) & (term[:, 1:] == b) & valid[:, 1:]
        want[i][d0:d1] = hit.any(dim=1)
    del term, valid
ok = 0
for i in range(n_check):
    bits = gi.filter_bitmap(ids[i], 0, n_docs)
    got = np.unpackbits(bits.view(np.uint8), bitorder="little")[:n_docs].astype(bool)
    w = want[i].cpu().numpy()
    assert np.array_equal(got, w), f"phrase {phrases[i]}: {int(got.sum())} docs, generator says {int(w.sum())}"
    ok += 1
    assert w.any()
print(f"full-size check: {ok} of {ok} phrase bitmaps equal the generator's adjacent-token test over all {n_docs} docs", flush=True)

# phrase queries: both terms scored, the phrase required (QueryString with a phrase group, api/reader.rs:1504-1508)
qb = QueryBatch.from_term_lists([[a, b] for a, b in phrases])
qb.filter_id = np.array(ids, dtype=np.int32)
for exe in ("bm25", "bmw"):
    p = gi.prepare(qb, 11, exe)
    p.run(sync=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        p.run(sync=True)
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / 3
    hits, counts = p.fetch()
    p.free()
    print(f"{n_phr} phrase queries, top-10, {exe}: {ms:.2f} ms/batch, {n_phr / ms * 1e3:.0f} q/s; hits/query min {int(counts.min())}", flush=True)
    # every returned doc is in its phrase's bitmap
    for i in range(n_check):
        w = want[i]
        docs = torch.from_numpy(hits[i, : counts[i]]["doc_id"].astype(np.int64)).to(dev)
        assert bool(w[docs].all())
c = gi.counters()
print(f"resident {c['resident_bytes'] / 1e9:.1f} GB; kernels launched {c['kernel_launches']}", flush=True)
gi.close()
