"""CPU test of the N > 1 path: world_size 2 over gloo.  Each rank holds one doc-range segment, scores
every query on it (with the oracle standing in for the CUDA engine — there is no GPU here), the
per-rank top-k lists are exchanged with ShardedSearcher's all-gather and merged in SortKey order
(searchlite-core/src/api/reader.rs:2777).  The result must equal the single-process search over
the same two segments, on every rank."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import slo
from searchlite_b200 import synth
from searchlite_b200.shard import ShardedSearcher, shard_ranges

N_DOCS, VOCAB, NQ, K = 6000, 800, 40, 11


def _segment(rank, world):
    lo, hi = shard_ranges(N_DOCS, world)[rank]
    spec = synth.CorpusSpec(n_docs=hi - lo, vocab=VOCAB, seed=21, len_lo=10, len_hi=50, segment_ord=rank, doc_base=lo)
    return synth.generate_segment(spec, "cpu")


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ora = slo.OracleIndex(_segment(rank, world))
        qb = synth.generate_queries(NQ, VOCAB, seed=22)
        h, c = ora.search_batch(qb, K, "bm25")
        s = ShardedSearcher(None, NQ, K, host_merge=slo.merge_hits)
        mh, mc = s.merge_cpu(h, c)
        np.save(os.path.join(out_dir, f"hits{rank}.npy"), mh)
        np.save(os.path.join(out_dir, f"counts{rank}.npy"), mc)
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_and_merge_equals_single_process(tmp_path):
    world = 2
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    # single process: both segments sequentially, then hits.sort_by(SortKey)
    qb = synth.generate_queries(NQ, VOCAB, seed=22)
    per_seg = []
    for r in range(world):
        ora = slo.OracleIndex(_segment(r, world))
        per_seg.append(ora.search_batch(qb, K, "bm25"))
    for r in range(world):
        mh = np.load(tmp_path / f"hits{r}.npy")
        mc = np.load(tmp_path / f"counts{r}.npy")
        for q in range(NQ):
            want = slo.merge_hits([h[q, : c[q]] for h, c in per_seg], K)
            got = mh[q, : mc[q]]
            assert mc[q] == len(want)
            assert np.array_equal(got["segment_ord"], want["segment_ord"]) and np.array_equal(got["doc_id"], want["doc_id"])
            assert np.array_equal(got["score"].view(np.uint32), want["score"].view(np.uint32))
    # both segment ordinals show up in the merged lists (the exchange really happened)
    assert set(np.unique(np.load(tmp_path / "hits0.npy")["segment_ord"][:, 0]).tolist()) <= {0, 1}
    assert len(np.unique(np.load(tmp_path / "hits1.npy")["segment_ord"])) >= 2
