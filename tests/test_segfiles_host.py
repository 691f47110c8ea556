"""CPU tests for SURVEY §8f rows 1-2: the host-side readers of the reference's segment files (no device
needed: slg_inspect_segment_files), the test writer against the oracle's codec, and the oracle's
matches_phrase against the reference's literal tests (query/phrase.rs:50-118)."""
from __future__ import annotations

import ctypes as C
import json
import struct
import zlib

import numpy as np
import pytest

from oracle import slo
from searchlite_b200.engine import SearchliteGpuError, inspect_segment_files
from tests import segwriter as sw
from tests.helpers import f32_bits


def small_segment(seed: int = 3, n_docs: int = 300, vocab: int = 60):
    rng = np.random.default_rng(seed)
    docs = []
    for _ in range(n_docs):
        n = int(rng.integers(3, 40))
        toks = (rng.zipf(1.4, size=n) % vocab).tolist()
        docs.append({"body": [f"t{t}" for t in toks], "title": [f"t{t}" for t in toks[:3]]})
    langs = [None if i % 11 == 0 else ["en", "de", "FR"][i % 3] for i in range(n_docs)]
    years = [None if i % 13 == 0 else 2000 + (i * 7) % 26 for i in range(n_docs)]
    return sw.Segment("a1", docs, ["body", "title"], keywords={"lang": langs}, i64s={"year": years},
                      f64s={"price": [float(i) * 0.5 for i in range(n_docs)]},
                      i64_lists={"tags": [[i, i + 1][: i % 3] for i in range(n_docs)]},
                      keyword_lists={"labels": [["Red", "green", "BLUE"][: i % 4] for i in range(n_docs)]},
                      i64_nested={"sizes": [[[i % 5, 9][: j + 1] for j in range(i % 3)] for i in range(n_docs)]})


def files_of(seg):
    post, terms = seg.post_and_terms()
    return terms, post, seg.fast(), seg.meta()


def test_inspect_reports_what_the_writer_wrote():
    seg = small_segment()
    terms, post, fast, meta = files_of(seg)
    crcs = [zlib.crc32(x) for x in (terms, post, fast, meta)]
    info = inspect_segment_files(len(seg.docs), terms, post, fast, meta, "body", checksums=crcs)
    table = seg.postings()
    body = {k: v for k, v in table.items() if k.startswith("body:")}
    assert info["n_terms_total"] == len(table)
    assert info["n_terms_field"] == len(body)
    assert info["n_postings"] == sum(len(v) for v in body.values())
    assert f32_bits(info["avgdl"]) == f32_bits(np.float32(seg.avg_field_lengths()["body"]))
    assert info["has_positions"] == 1 and info["has_length_column"] == 1
    assert info["n_fast_columns"] == 8 and info["n_scalar_columns"] == 5  # _len:body, _len:title, year, price, lang
    assert info["n_list_columns"] == 3  # tags (I64List), labels (StrList), sizes (I64Nested)
    # crc32 == crc32fast (IEEE): util/checksum.rs:3-7
    assert [info["crc_terms"], info["crc_postings"], info["crc_fast"], info["crc_meta"]] == crcs
    other = inspect_segment_files(len(seg.docs), terms, post, fast, meta, "nosuchfield")
    assert other["n_terms_field"] == 0 and other["avgdl"] == 0.0 and other["has_length_column"] == 0


def test_corrupt_files_are_rejected_like_the_reference():
    seg = small_segment(n_docs=50)
    terms, post, fast, meta = files_of(seg)
    n = len(seg.docs)
    # index/terms.rs:96-108 invalid_checksum_errors: last byte bumped
    bad = bytearray(terms)
    bad[-1] = (bad[-1] + 1) & 0xFF
    with pytest.raises(SearchliteGpuError, match="failed checksum validation"):
        inspect_segment_files(n, bytes(bad), post, fast, meta, "body")
    with pytest.raises(SearchliteGpuError, match="truncated"):  # terms.rs:29-31
        inspect_segment_files(n, terms[:8], post, fast, meta, "body")
    # verify_checksums, index/segment.rs:1140-1160
    crcs = [zlib.crc32(x) for x in (terms, post, fast, meta)]
    flipped = bytearray(post)
    flipped[len(flipped) // 2] ^= 0x40
    with pytest.raises(SearchliteGpuError, match="failed checksum for postings"):
        inspect_segment_files(n, terms, bytes(flipped), fast, meta, "body", checksums=crcs)
    with pytest.raises(SearchliteGpuError, match="FFV1"):
        inspect_segment_files(n, terms, post, b"XXXX" + fast[4:], meta, "body")
    with pytest.raises(SearchliteGpuError, match="ended unexpectedly"):
        inspect_segment_files(n, terms, post, fast[: len(fast) // 2], meta, "body")
    with pytest.raises(SearchliteGpuError, match="rows"):
        inspect_segment_files(n + 1, terms, post, fast, meta, "body")
    with pytest.raises(SearchliteGpuError, match="JSON"):
        inspect_segment_files(n, terms, post, fast, b"[1, 2]", "body")


def test_writer_codec_matches_the_oracle_codec():
    """two independent restatements of PostingsWriter::write_term (index/postings.rs:78-129) agree byte for byte,
    with and without positions, across the 128-posting block boundary"""
    rng = np.random.default_rng(5)
    L = slo.lib()
    for n in (0, 1, 127, 128, 129, 400):
        docs = np.sort(rng.choice(1 << 22, size=n, replace=False)).astype(np.uint32)
        tfs = rng.integers(1, 6, size=n).astype(np.uint32)
        pos_lists = [np.sort(rng.choice(300, size=int(t), replace=False)).astype(np.uint32) for t in tfs]
        poff = np.zeros(n + 1, dtype=np.uint32)
        poff[1:] = np.cumsum([len(p) for p in pos_lists])
        flat = np.concatenate(pos_lists).astype(np.uint32) if n else np.zeros(0, dtype=np.uint32)
        for keep in (0, 1):
            mine = sw.encode_postings([(int(d), int(t), p.tolist()) for d, t, p in zip(docs, tfs, pos_lists)], bool(keep))
            size = L.slo_postings_encode(docs.ctypes.data, tfs.ctypes.data, n, keep, poff.ctypes.data, flat.ctypes.data, None, 0)
            buf = np.zeros(max(size, 1), dtype=np.uint8)
            L.slo_postings_encode(docs.ctypes.data, tfs.ctypes.data, n, keep, poff.ctypes.data, flat.ctypes.data, buf.ctypes.data, size)
            assert bytes(buf[:size]) == mine, (n, keep)


def test_terms_file_layout_roundtrip():
    """index/terms.rs:77-93 roundtrips_terms_file: alpha/beta/gamma with offsets 10/20/30"""
    body = b"".join(sw.varint(len(k)) + k + struct.pack("<Q", o) for k, o in ((b"alpha", 10), (b"beta", 20), (b"gamma", 30)))
    terms = struct.pack("<Q", 3) + body + struct.pack("<I", zlib.crc32(body))
    # offsets must point into the posting image: give it 64 bytes of empty lists
    post = bytes(64)
    fast = b"FFV1" + struct.pack("<I", 0)
    info = inspect_segment_files(0, terms, post, fast, b"{}", "alpha")
    assert info["n_terms_total"] == 3 and info["n_terms_field"] == 0  # no "alpha:" key: keys are whole strings


# ---- matches_phrase: the reference's own unit tests, query/phrase.rs:50-118 ----
def test_phrase_matches_consecutive_positions():
    assert slo.matches_phrase_positions([[1, 4], [2], [3]], 0)


def test_phrase_rejects_non_consecutive_positions():
    assert not slo.matches_phrase_positions([[1], [3]], 0)


def test_phrase_allows_sloppy_phrase():
    assert not slo.matches_phrase_positions([[1], [4], [6]], 0)
    assert slo.matches_phrase_positions([[1], [4], [6]], 3)
    assert not slo.matches_phrase_positions([[1], [4], [6]], 2)


def test_phrase_edge_cases():
    assert slo.matches_phrase_positions([], 0)                 # phrase.rs:5-7
    assert not slo.matches_phrase_positions([[1], []], 5)      # :16-18
    assert slo.matches_phrase_positions([[9]], 0)              # :19-21
    assert not slo.matches_phrase_positions([[5], [4]], 10)    # order matters: pos <= prev is skipped
    assert slo.matches_phrase_positions([[0, 7], [3, 8]], 0)   # a later start succeeds
    assert slo.matches_phrase_positions([[2], [2, 3]], 0)      # the same term twice ("a a"): strictly increasing


def _greedy(lists, slop):
    """the closed form the CUDA kernel uses (slg_phrase.cuh): min over starts of the greedy chain's span"""
    if not lists:
        return True
    if any(len(l) == 0 for l in lists):
        return False
    if len(lists) == 1:
        return True
    for p0 in lists[0]:
        prev, ok = p0, True
        for l in lists[1:]:
            nxt = [p for p in l if p > prev]
            if not nxt:
                ok = False
                break
            prev = nxt[0]
        if ok and prev - p0 - (len(lists) - 1) <= slop:
            return True
    return False


def test_greedy_chain_equals_the_recursive_search():
    rng = np.random.default_rng(11)
    for _ in range(3000):
        n = int(rng.integers(1, 5))
        lists = [sorted(set(rng.integers(0, 24, size=int(rng.integers(0, 6))).tolist())) for _ in range(n)]
        slop = int(rng.integers(0, 6))
        assert slo.matches_phrase_positions(lists, slop) == _greedy(lists, slop), (lists, slop)


def test_phrase_bitmap_over_a_segment():
    seg = small_segment(seed=9, n_docs=120, vocab=12)
    keys, toff, docs, tfs, poff, pos, lens = sw.csr_of(seg, "body")
    a, b = keys.index("body:t1"), keys.index("body:t2")
    bm = slo.phrase_bitmap(len(seg.docs), toff, docs, poff, pos, [a, b], 0)
    for d, doc in enumerate(seg.docs):
        toks = doc["body"]
        want = any(toks[i] == "t1" and toks[i + 1] == "t2" for i in range(len(toks) - 1))
        assert bool((bm[d >> 5] >> (d & 31)) & 1) == want, d
    assert not slo.phrase_bitmap(len(seg.docs), toff, docs, poff, pos, [a, len(keys)], 0).any()  # absent term


def test_manifest_and_meta_are_plain_json(tmp_path):
    seg = small_segment(n_docs=20)
    man = sw.write_index(str(tmp_path), [seg])
    on_disk = json.load(open(tmp_path / "MANIFEST.json"))
    assert on_disk["segments"][0]["doc_count"] == 20 and on_disk == man
    meta = json.load(open(tmp_path / "seg_a1.meta"))
    assert set(meta["avg_field_lengths"]) == {"body", "title"}


def test_synthetic_positions_follow_the_posting_order():
    """synth.generate_positions: positions of posting p = the token indexes of its term in its doc, ascending"""
    from searchlite_b200 import synth
    spec = synth.CorpusSpec(n_docs=3000, vocab=500, seed=5, len_lo=5, len_hi=40)
    seg = synth.generate_segment(spec, "cpu")
    pos = synth.generate_positions(spec, "cpu", chunk_docs=1000).numpy()
    assert len(pos) == seg.total_tokens == int(seg.post_tfs.sum())
    off = np.zeros(len(seg.post_tfs) + 1, dtype=np.int64)
    off[1:] = np.cumsum(seg.post_tfs.astype(np.int64))
    term, valid = synth.token_terms(spec, 0, spec.n_docs, synth.zipf_cdf(spec.vocab, spec.zipf_s), "cpu")
    term, valid = term.numpy(), valid.numpy()
    toff = seg.term_offsets.astype(np.int64)
    for t in (0, 1, 7, 100, 499):
        for p in range(toff[t], min(toff[t + 1], toff[t] + 40)):
            d, ps = seg.post_docs[p], pos[off[p]:off[p + 1]]
            assert (np.diff(ps) > 0).all() and (term[d, ps] == t).all()
            assert int((term[d][valid[d]] == t).sum()) == len(ps) == seg.post_tfs[p]


def test_oracle_image_with_positions_equals_the_writer_image():
    """slo_index_build_post_image with positions on (tools/fileload_bench.py's encoder) == tests/segwriter.py's .post bytes"""
    from searchlite_b200.engine import SegmentData
    seg = small_segment(seed=4, n_docs=400, vocab=50)
    seg.text_fields = ["body"]
    keys, toff, docs, tfs, poff, pos, lens = sw.csr_of(seg, "body")
    o = slo.OracleIndex(SegmentData(0, len(seg.docs), toff, docs, tfs, lens, int(lens.sum())))
    o.set_positions(poff, pos)
    img, off = o.build_post_image()
    post, _ = seg.post_and_terms()  # keys sorted == term id order of csr_of
    assert img.tobytes() == post and off[-1] == len(post)


def test_parallel_crc32_of_a_large_file_equals_zlib():
    """files above a few MB are checksummed on all host cores and the parts combined (GF(2) shift): same value as one pass"""
    rng = np.random.default_rng(1)
    terms = struct.pack("<Q", 0) + struct.pack("<I", zlib.crc32(b""))
    fast = b"FFV1" + struct.pack("<I", 0)
    for n in (50_000_003, 4 * (4 << 20), 8_388_609):
        post = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        info = inspect_segment_files(0, terms, post, fast, b"{}", "body")
        assert info["crc_postings"] == zlib.crc32(post), n
