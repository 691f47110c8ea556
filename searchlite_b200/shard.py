"""Doc-range / segment sharding across the GPUs of one box.

One process per GPU, each holding one segment (== shard).  Per batch every rank scores all
queries against its own segment, the per-rank top-k lists (Q x k x 12 B) are exchanged with ONE
all-gather over NCCL/NVLink, and every rank merges the gathered lists with the device merge
kernel (slg_merge_gathered) in the reference's SortKey order — score desc, segment_ord asc,
doc_id asc (searchlite-core/src/api/reader.rs:2777, src/query/sort.rs:80-93).  Because N, df
and avgdl are per segment in the reference (api/reader.rs:2985,2994), the N-GPU result equals
the reference run on the same N-segment index with no statistics exchange.

The same code runs with the `gloo` backend on CPU tensors (tests): then the merge is done by
the callable passed as `host_merge` because the CUDA merge kernel needs a device.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .engine import HIT_DTYPE

HIT_BYTES = HIT_DTYPE.itemsize


def shard_ranges(n_docs: int, world: int) -> list:
    """contiguous doc ranges, sizes differing by at most one"""
    base, rem = divmod(n_docs, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def gather_hits(local_hits: torch.Tensor, local_counts: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """all-gather of per-rank results.  local_hits: uint8 [Q*k*12], local_counts: int32 [Q].
    Returns ([world, Q*k*12] uint8, [world, Q] int32), rank-major."""
    world = dist.get_world_size(group)
    # flat outputs (rank-major concatenation): the form both NCCL and gloo accept
    hits = torch.empty(world * local_hits.numel(), dtype=local_hits.dtype, device=local_hits.device)
    counts = torch.empty(world * local_counts.numel(), dtype=local_counts.dtype, device=local_counts.device)
    dist.all_gather_into_tensor(hits, local_hits.reshape(-1), group=group)
    dist.all_gather_into_tensor(counts, local_counts.reshape(-1), group=group)
    return hits.view(world, -1), counts.view(world, -1)


class ShardedSearcher:
    """search over world_size shards; every rank ends up with the merged result"""

    def __init__(self, index, n_queries: int, k: int, group=None, host_merge: Optional[Callable] = None):
        self.index, self.q, self.k, self.group = index, n_queries, k, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.host_merge = host_merge
        self.stream = None
        if index is not None:
            dev = torch.device("cuda", index.device)
            self.stream = torch.cuda.ExternalStream(index.stream_ptr(), device=dev)
            self.send_hits = torch.empty(n_queries * k * HIT_BYTES, dtype=torch.uint8, device=dev)
            self.send_counts = torch.empty(n_queries, dtype=torch.int32, device=dev)

    def exchange_and_merge(self, prepared) -> Tuple[np.ndarray, np.ndarray]:
        """after prepared.run(): gather every rank's top-k and merge on the device"""
        if self.world == 1:
            return prepared.fetch()
        with torch.cuda.stream(self.stream):
            prepared.copy_results_to(self.send_hits.data_ptr(), self.send_counts.data_ptr())
            hits, counts = gather_hits(self.send_hits, self.send_counts, self.group)
            # NCCL work is ordered after the current (= the handle's) stream and the stream waits for it
            return self.index.merge_gathered(hits.data_ptr(), counts.data_ptr(), self.world, self.q, self.k)

    def merge_cpu(self, local_hits: np.ndarray, local_counts: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """gloo path used by the CPU tests: same exchange, merge by `host_merge`"""
        lh = torch.from_numpy(np.ascontiguousarray(local_hits).view(np.uint8).reshape(-1))
        lc = torch.from_numpy(np.ascontiguousarray(local_counts).astype(np.int32))
        hits, counts = gather_hits(lh, lc, self.group)
        hits = hits.numpy().view(HIT_DTYPE).reshape(self.world, self.q, self.k)
        counts = counts.numpy().astype(np.uint32)
        out_h = np.zeros((self.q, self.k), dtype=HIT_DTYPE)
        out_c = np.zeros(self.q, dtype=np.uint32)
        for qi in range(self.q):
            lists = [hits[r, qi, : counts[r, qi]] for r in range(self.world)]
            merged = self.host_merge(lists, self.k)
            out_h[qi, : len(merged)] = merged
            out_c[qi] = len(merged)
        return out_h, out_c
