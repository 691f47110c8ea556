"""Where the time of a sharded step goes (torchrun, N ranks): local scan in one launch vs in two parts with the threshold
exchange, piece by piece (CUDA events on the handle's stream, max over ranks is NOT taken: rank 0's view)."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from searchlite_b200 import GpuIndex, synth  # noqa: E402
from searchlite_b200.shard import ShardedSearcher, shard_ranges  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n_docs, Q, k = int(os.environ.get("DOCS", 10_000_000)), 4096, 11
lo, hi = shard_ranges(n_docs, world)[rank]
seg = synth.generate_segment(synth.CorpusSpec(n_docs=hi - lo, vocab=1_000_000, seed=20260101, segment_ord=rank, doc_base=lo), dev)
qb = synth.generate_queries(Q, 1_000_000, seed=20260102)
opts = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in sys.argv[1:]}
gi = GpuIndex(local, options=opts)
gi.load_segment(seg)
del seg
torch.cuda.empty_cache()
stream = torch.cuda.ExternalStream(gi.stream_ptr(), device=dev)
s = ShardedSearcher(gi, Q, k)
p = gi.prepare(qb, k, "bmw")


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record(stream)
    return e


if rank == 0:
    print("board", s.board is not None, getattr(s, "board_error", None), flush=True)
for mode in ("one-step", "board", "two-step"):
    acc = {}
    for it in range(8):
        dist.barrier()
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            marks = [("start", ev())]
            t0 = time.perf_counter()
            if mode == "one-step":
                p.set_threshold_board(0, [], 0)
                p.run(sync=False)
                marks.append(("scan", ev()))
            elif mode == "board":
                s.attach_board(p)
                p.run(sync=False)
                marks.append(("scan", ev()))
            else:
                p.set_threshold_board(0, [], 0)
                assert p.run_seeds()
                marks.append(("first part", ev()))
                keys = s._as_tensor(p.threshold_keys_ptr(), Q * 8, torch.int64)
                glob = keys.clone()
                dist.all_reduce(glob, op=dist.ReduceOp.MAX)
                marks.append(("all-reduce", ev()))
                p.import_thresholds(glob.data_ptr())
                p.run_sweep(sync=False)
                marks.append(("rest", ev()))
            host_ms = 1e3 * (time.perf_counter() - t0)
            s._exchange(p)
            marks.append(("gather+merge", ev()))
        torch.cuda.synchronize()
        if it >= 3:
            for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
                acc[name] = acc.get(name, 0.0) + a.elapsed_time(b) / 5
            acc["host enqueue"] = acc.get("host enqueue", 0.0) + host_ms / 5
    p.fetch()
    c = gi.counters()
    if rank == 0:
        print(mode, {n: round(v, 3) for n, v in acc.items()}, "scanned", c["last_postings_scattered"], "verified", c["last_postings_verified"], flush=True)
    dist.barrier()
p.free()
dist.destroy_process_group()
