"""Residency from the reference's on-disk files at scale: writes a synthetic corpus as a
searchlite index directory (MANIFEST.json, seg_*.terms/.post/.fast/.meta, positions ON like every index the
reference's tests and benches write), loads it with slg_load_index_dir and checks, at that size, that
  * a batch of OR queries returns byte-identical hits to the same corpus loaded through the CSR path,
  * phrase bitmaps from the file-decoded positions equal those from CSR-loaded positions.
The `.post` image is encoded by the oracle's restatement of PostingsWriter::write_term (C++), the rest by
tests/segwriter.py's layout code.  Usage: python tools/fileload_bench.py [n_docs] [n_segments]"""
import json
import os
import struct
import sys
import tempfile
import time
import zlib

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from oracle import slo  # noqa: E402  (the checker's codec writes the test image)
from searchlite_b200 import GpuIndex, synth  # noqa: E402
from searchlite_b200.engine import QueryBatch  # noqa: E402
from searchlite_b200.shard import shard_ranges  # noqa: E402
from tests import segwriter as sw  # noqa: E402

n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n_segs = int(sys.argv[2]) if len(sys.argv) > 2 else 2
vocab = 200_000
dev = "cuda:0" if torch.cuda.is_available() else "cpu"
root = tempfile.mkdtemp(prefix="slg_index_")

segs, metas, total_bytes = [], [], 0
t_write = time.perf_counter()
for ord_, (lo, hi) in enumerate(shard_ranges(n_docs, n_segs)):
    spec = synth.CorpusSpec(n_docs=hi - lo, vocab=vocab, seed=20260101, segment_ord=ord_, doc_base=lo)
    seg = synth.generate_segment(spec, dev).to_host()
    positions = synth.generate_positions(spec, dev).cpu().numpy().view(np.uint32)
    pos_off = np.zeros(len(seg.post_tfs) + 1, dtype=np.uint64)
    pos_off[1:] = np.cumsum(seg.post_tfs.astype(np.uint64))
    o = slo.OracleIndex(seg)
    o.set_positions(pos_off, positions)
    img, off = o.build_post_image()
    df = np.diff(seg.term_offsets.astype(np.int64))
    present = np.nonzero(df)[0]  # the writer only stores terms that occur
    body = bytearray()
    for t in present.tolist():
        kb = b"body:%d" % t
        body += sw.varint(len(kb)) + kb + struct.pack("<Q", int(off[t]))
    terms = struct.pack("<Q", len(present)) + bytes(body) + struct.pack("<I", zlib.crc32(bytes(body)))
    n = hi - lo
    year = (2000 + (np.arange(n) * 7) % 26).astype("<i8")
    fast = b"FFV1" + struct.pack("<I", 2)
    for name, vals in ((b"year", year), (b"_len:body", seg.field_lengths.astype("<i8"))):
        fast += struct.pack("<I", len(name)) + name + bytes([0]) + struct.pack("<I", n) + bytes([1]) * n + vals.tobytes()
    avg = float(np.float32(seg.total_tokens) / np.float32(n))
    meta = json.dumps({"doc_offsets": [], "doc_ids": [], "avg_field_lengths": {"body": sw.f32_json(avg)}, "use_zstd": False},
                      indent=2).encode()
    files = {"terms": terms, "postings": img.tobytes(), "docstore": b"", "fast": fast, "meta": meta}
    names = {"terms": f"seg_{ord_}.terms", "postings": f"seg_{ord_}.post", "docstore": f"seg_{ord_}.docs",
             "fast": f"seg_{ord_}.fast", "meta": f"seg_{ord_}.meta"}
    for k, data in files.items():
        with open(os.path.join(root, names[k]), "wb") as f:
            f.write(data)
        total_bytes += len(data)
    metas.append({"id": str(ord_), "generation": 1, "paths": {k: os.path.join("/elsewhere", v) for k, v in names.items()},
                  "doc_count": n, "max_doc_id": n - 1, "blockmax": True, "deleted_docs": [],
                  "avg_field_lengths": {"body": sw.f32_json(avg)}, "checksums": {k: zlib.crc32(v) for k, v in files.items()}})
    segs.append((seg, pos_off, positions, present))
    del o
with open(os.path.join(root, "MANIFEST.json"), "w") as f:
    json.dump({"version": 1, "uuid": "00000000-0000-4000-8000-000000000000", "segments": metas,
               "committed_at": "2026-01-01T00:00:00+00:00", "schema": {}}, f, indent=2)
n_post = sum(len(s[0].post_docs) for s in segs)
n_pos = sum(len(s[2]) for s in segs)
print(f"index written: {n_docs} docs in {n_segs} segments, {n_post} postings, {n_pos} positions, {total_bytes / 1e9:.2f} GB of files "
      f"({time.perf_counter() - t_write:.1f} s to generate + encode)", flush=True)
if not torch.cuda.is_available():  # CPU dry run: the host-side validation pass only
    from searchlite_b200.engine import inspect_segment_files
    rd = lambda n: open(os.path.join(root, n), "rb").read()
    for m in metas:
        i = m["id"]
        print(inspect_segment_files(m["doc_count"], rd(f"seg_{i}.terms"), rd(f"seg_{i}.post"), rd(f"seg_{i}.fast"), rd(f"seg_{i}.meta"), "body",
                                    checksums=[m["checksums"][k] for k in ("terms", "postings", "fast", "meta")]))
    sys.exit(0)
torch.cuda.empty_cache()

for keep in (1, 0):
    gi = GpuIndex(0, kernel="warp")
    gi.set_option("keep_positions", keep)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    assert gi.load_index_dir(root, "body") == n_segs
    torch.cuda.synchronize()
    s = time.perf_counter() - t0
    print(f"slg_load_index_dir, keep_positions={keep}: {s:.2f} s = {total_bytes / s / 1e9:.2f} GB/s of file bytes "
          f"(read + crc32 + parse + copy + decode + derived arrays); resident {gi.counters()['resident_bytes'] / 1e9:.2f} GB", flush=True)
    if keep:
        files_gi = gi
    else:
        gi.close()

# the same corpus through the CSR path, in the file-loaded handle's term space
csr = GpuIndex(0, kernel="warp")
n_space = 0
ids_of = {}
for seg, pos_off, positions, present in segs:
    ids = np.array([files_gi.term_lookup("body:%d" % t) for t in present.tolist()], dtype=np.int64)
    n_space = max(n_space, int(ids.max()) + 1)
    ids_of[seg.segment_ord] = ids
for seg, pos_off, positions, present in segs:
    ids = ids_of[seg.segment_ord]
    df = np.diff(seg.term_offsets.astype(np.int64))
    g_df = np.zeros(n_space, dtype=np.int64)
    g_df[ids] = df[present]
    g_off = np.zeros(n_space + 1, dtype=np.uint64)
    g_off[1:] = np.cumsum(g_df)
    # permutation of postings: global term order
    order = np.argsort(ids, kind="stable")
    src_lo = seg.term_offsets.astype(np.int64)[present][order]
    cnt = df[present][order]
    idx = np.repeat(src_lo - np.concatenate(([0], np.cumsum(cnt)[:-1])), cnt) + np.arange(int(cnt.sum()))
    docs, tfs = seg.post_docs[idx], seg.post_tfs[idx]
    g_poff = np.zeros(len(idx) + 1, dtype=np.uint64)
    g_poff[1:] = np.cumsum(tfs.astype(np.uint64))
    plen = np.diff(pos_off.astype(np.int64))[idx]
    pidx = np.repeat(pos_off.astype(np.int64)[idx] - np.concatenate(([0], np.cumsum(plen)[:-1])), plen) + np.arange(int(plen.sum()))
    g_pos = positions[pidx]
    from searchlite_b200.engine import SegmentData
    csr.load_segment(SegmentData(seg.segment_ord, seg.doc_count, g_off, docs, tfs, seg.field_lengths, seg.total_tokens))
    csr.load_positions(seg.segment_ord, g_off, g_poff, g_pos)

rng = np.random.default_rng(7)
all_present = np.unique(np.concatenate([sg[3] for sg in segs]))
cdf = synth.zipf_cdf(vocab, 1.0).numpy()
lists = []
for _ in range(1024):
    n = int(rng.integers(2, 6))
    t = np.unique(np.searchsorted(cdf, rng.random(n) * (1 - cdf[8]) + cdf[8]))
    lists.append([files_gi.term_lookup("body:%d" % int(x)) for x in t])
qb = QueryBatch.from_term_lists(lists)
for exe in ("bm25", "bmw"):
    a, b = files_gi.search_batch(qb, 11, exe), csr.search_batch(qb, 11, exe)
    assert a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes(), exe
print(f"1024 OR queries, bm25 and bmw: file-loaded == CSR-loaded, byte for byte (hits/query min {int(a[1].min())})", flush=True)

phr = []
spec0 = synth.CorpusSpec(n_docs=segs[0][0].doc_count, vocab=vocab, seed=20260101)
cdf_d = synth.zipf_cdf(vocab, 1.0).to(dev)
while len(phr) < 64:
    d = int(rng.integers(0, spec0.n_docs))
    term, valid = synth.token_terms(spec0, d, d + 1, cdf_d, dev)
    t = term[0][valid[0]].cpu().numpy()
    p = int(rng.integers(0, len(t) - 2))
    phr.append([files_gi.term_lookup("body:%d" % int(x)) for x in t[p:p + int(rng.integers(2, 4))]])
slops = rng.integers(0, 3, size=len(phr)).tolist()
fa, fb = files_gi.compile_phrases(phr, slops), csr.compile_phrases(phr, slops)
nonempty = 0
for x, y in zip(fa, fb):
    for seg, *_ in segs:
        ba, bb = files_gi.filter_bitmap(int(x), seg.segment_ord, seg.doc_count), csr.filter_bitmap(int(y), seg.segment_ord, seg.doc_count)
        assert np.array_equal(ba, bb)
        nonempty += bool(ba.any())
print(f"{len(phr)} phrases (2-3 terms, slop 0-2): bitmaps from file-decoded positions == from CSR positions ({nonempty} non-empty)", flush=True)
files_gi.close()
csr.close()
import shutil
shutil.rmtree(root)
