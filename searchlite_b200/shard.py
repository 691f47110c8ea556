"""Doc-range / segment sharding across the GPUs of one box.

One process per GPU, each holding one segment (== shard).  Per batch every rank scores all
queries against its own segment, the per-rank result blocks (Q x k x 12 B of hits + Q x 4 B of counts)
are exchanged with ONE all-gather over NCCL/NVLink, and every rank merges the gathered lists with the device merge
kernel (slg_merge_gathered) in the reference's SortKey order — score desc, segment_ord asc,
doc_id asc (searchlite-core/src/api/reader.rs:2777, src/query/sort.rs:80-93).  Because N, df
and avgdl are per segment in the reference (api/reader.rs:2985,2994), the N-GPU result equals
the reference run on the same N-segment index with no statistics exchange.

Threshold board (on by default at N > 1 on CUDA): n_queries x 8 bytes of symmetric memory per rank, mapped into every
peer over NVLink; while the posting scan runs, a shard that raises a query's k-th score pushes it into its peers' boards and
prunes against the best score any shard has found (slg_batch_set_threshold_board).  Results are exact either way.

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


def gather_blocks(local_block: torch.Tensor, group=None) -> torch.Tensor:
    """ONE all-gather of the per-rank result blocks (uint8, Q*k*12 bytes of hits followed by Q*4 bytes of counts — the
    layout of slg_batch_packed_results).  Returns [world, block bytes], rank-major."""
    world = dist.get_world_size(group)
    out = torch.empty(world * local_block.numel(), dtype=local_block.dtype, device=local_block.device)
    dist.all_gather_into_tensor(out, local_block.reshape(-1), group=group)
    return out.view(world, -1)


class ShardedSearcher:
    """search over world_size shards; every rank ends up with the merged result"""

    def __init__(self, index, n_queries: int, k: int, group=None, host_merge: Optional[Callable] = None, threshold_board: bool = True):
        self.index, self.q, self.k, self.group = index, n_queries, k, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.host_merge = host_merge
        self.stream = None
        self.block_bytes = n_queries * k * HIT_BYTES + n_queries * 4
        self.board = None
        self.epoch = 0
        if index is not None:
            self.dev = torch.device("cuda", index.device)
            self.stream = torch.cuda.ExternalStream(index.stream_ptr(), device=self.dev)
            if self.world > 1 and self.world <= 8 and threshold_board:
                self._open_board()

    def _open_board(self) -> None:
        """the threshold board: n_queries x 8 bytes of symmetric memory per rank, every rank's buffer mapped into every other
        rank's address space over NVLink (torch symmetric memory: CUDA VMM + handle exchange through the store).  Stays off —
        the shards then prune against their own thresholds only — where the platform cannot map peers."""
        try:
            import torch.distributed._symmetric_memory as symm_mem
            buf = symm_mem.empty(self.q, dtype=torch.int64, device=self.dev)
            buf.zero_()
            hdl = symm_mem.rendezvous(buf, group=self.group if self.group is not None else dist.group.WORLD)
            torch.cuda.synchronize()
            hdl.barrier()
            rank = dist.get_rank(self.group)
            self.board = (buf, hdl, int(hdl.buffer_ptrs[rank]), [int(p) for r, p in enumerate(hdl.buffer_ptrs) if r != rank])
        except Exception as e:  # noqa: BLE001
            self.board = None
            self.board_error = repr(e)

    def _as_tensor(self, ptr: int, nbytes: int, dtype=torch.uint8) -> torch.Tensor:
        """zero-copy view of device memory the engine owns (valid until the batch is re-run or freed)"""
        itemsize = torch.empty((), dtype=dtype).element_size()

        class _Arr:  # __cuda_array_interface__ v3
            pass
        a = _Arr()
        typestr = {torch.uint8: "|u1", torch.int64: "<i8"}[dtype]
        a.__cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": typestr, "data": (ptr, False), "version": 3,
                                      "strides": None}
        return torch.as_tensor(a, device=self.dev)

    def run(self, prepared, exchange_thresholds: bool = True):
        """one sharded search: local scoring (asynchronous), the exchange, the merge.  Pruned executions run in two steps
        with an all-reduce(max) of the per-query k-th keys after the seeds, so every shard prunes against the global
        bound."""
        if self.world == 1:
            prepared.run(sync=False)
            return prepared.fetch()
        with torch.cuda.stream(self.stream):
            if exchange_thresholds and self.board is not None:
                # thresholds travel between the shards INSIDE the scan (peer pushes over NVLink): one launch, no collective
                self.attach_board(prepared)
                prepared.run(sync=False)
            elif exchange_thresholds == "two-step" and prepared.two_step_ok and prepared.run_seeds():
                keys = self._as_tensor(prepared.threshold_keys_ptr(), self.q * 8, torch.int64)  # positive-score keys: top bit clear
                glob = keys.clone()
                dist.all_reduce(glob, op=dist.ReduceOp.MAX, group=self.group)
                prepared.import_thresholds(glob.data_ptr())
                prepared.run_sweep(sync=False)
            else:
                prepared.run(sync=False)
            return self._exchange(prepared)

    def attach_board(self, prepared) -> None:
        """a fresh epoch of the threshold board for the next run of `prepared` (every rank calls this the same number of times)"""
        if self.board is not None:
            self.epoch += 1
            try:
                prepared.set_threshold_board(self.board[2], self.board[3], self.epoch)
            except Exception as e:  # noqa: BLE001  (several segments per handle: the shards then keep to their own thresholds)
                self._board_keep = self.board  # (peers may still push into this rank's buffer: it stays mapped)
                self.board = None
                self.board_error = repr(e)

    def _exchange(self, prepared):
        ptr, nbytes = prepared.packed_results()
        assert nbytes == self.block_bytes
        blocks = gather_blocks(self._as_tensor(ptr, nbytes), self.group)
        # NCCL work is ordered after the current (= the handle's) stream and the stream waits for it
        return self.index.merge_gathered_packed(blocks.data_ptr(), self.world, self.q, self.k)

    def exchange_and_merge(self, prepared) -> Tuple[np.ndarray, np.ndarray]:
        """after prepared.run(): gather every rank's top-k with one all-gather and merge on the device"""
        if self.world == 1:
            return prepared.fetch()
        with torch.cuda.stream(self.stream):
            return self._exchange(prepared)

    def merge_cpu(self, local_hits: np.ndarray, local_counts: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """gloo path used by the CPU tests: same single-block exchange, merge by `host_merge`"""
        lh = np.ascontiguousarray(local_hits).view(np.uint8).reshape(-1)
        lc = np.ascontiguousarray(local_counts).astype(np.uint32).view(np.uint8).reshape(-1)
        blocks = gather_blocks(torch.from_numpy(np.concatenate([lh, lc])), self.group).numpy()
        hb = self.q * self.k * HIT_BYTES
        hits = np.ascontiguousarray(blocks[:, :hb]).view(HIT_DTYPE).reshape(self.world, self.q, self.k)
        counts = np.ascontiguousarray(blocks[:, hb:]).view(np.uint32).reshape(self.world, self.q)
        out_h = np.zeros((self.q, self.k), dtype=HIT_DTYPE)
        out_c = np.zeros(self.q, dtype=np.uint32)
        for qi in range(self.q):
            lists = [hits[r, qi, : counts[r, qi]] for r in range(self.world)]
            merged = self.host_merge(lists, self.k)
            out_h[qi, : len(merged)] = merged
            out_c[qi] = len(merged)
        return out_h, out_c
