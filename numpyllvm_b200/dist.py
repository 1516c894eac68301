"""One-process-per-GPU plumbing (torchrun) — control plane only.

torch.distributed (gloo) carries the rendezvous, barriers and the tiny host-side
exchanges (timings, the NCCL unique id); the data path never goes through torch:
reductions are finished inside their kernel over NVLink mailboxes
(nl_launch_pipeline_allreduce) or by the library's own NCCL communicator
(nl_finish_reduce) and filters by the exclusive scan of the shard counts
(nl_finish_filter).  The helpers below are the host logic of the N > 1 path and
are exercised on CPU by tests/test_dist_gloo.py with world_size = 2.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence, Tuple

import numpy as np


class Ranks:
    def __init__(self, n_gpus: int = 1, backend: str = "gloo"):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            if not dist.is_initialized():
                dist.init_process_group(backend, rank=self.rank, world_size=self.world)
            self.dist = dist
        elif n_gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")

    # -- control plane ---------------------------------------------------------
    def barrier(self):
        if self.dist:
            self.dist.barrier()

    def _reduce(self, x: float, op) -> float:
        if not self.dist:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64)
        self.dist.all_reduce(t, op=op)
        return float(t[0])

    def max(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.MAX) if self.dist else x

    def sum(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.SUM) if self.dist else x

    def bcast_bytes(self, b: bytes) -> bytes:
        if not self.dist:
            return b
        obj = [b]
        self.dist.broadcast_object_list(obj, src=0)
        return obj[0]

    def all_gather_int(self, v: int) -> List[int]:
        if not self.dist:
            return [int(v)]
        import torch
        out = [torch.zeros(1, dtype=torch.int64) for _ in range(self.world)]
        self.dist.all_gather(out, torch.tensor([int(v)], dtype=torch.int64))
        return [int(t[0]) for t in out]

    def all_gather_array(self, a: np.ndarray) -> List[np.ndarray]:
        if not self.dist:
            return [a]
        out = [None] * self.world
        self.dist.all_gather_object(out, a)
        return out

    def finish(self):
        if self.dist:
            self.dist.barrier()
            self.dist.destroy_process_group()
            self.dist = None

    # -- data-path communicator of the library -------------------------------------
    def init_library_comm(self, lib):
        """Rank 0 creates the NCCL id, everybody joins: local device 0 becomes rank `rank`."""
        from . import abi
        if self.world <= 1:
            return
        ident = C.create_string_buffer(abi.NL_UNIQUE_ID_BYTES)
        if self.rank == 0:
            abi.check(lib.nl_comm_unique_id(ident))
        raw = self.bcast_bytes(bytes(ident.raw))
        abi.check(lib.nl_comm_init(C.create_string_buffer(raw, abi.NL_UNIQUE_ID_BYTES), self.world, self.rank))
        self.init_mailboxes(lib)

    def init_mailboxes(self, lib) -> bool:
        """Exchange the CUDA IPC handles of the ranks' NVLink mailboxes (one tiny host all-gather)
        so that reductions can finish inside their kernel (nl_launch_pipeline_allreduce).  Returns
        whether the fused finish is available; the NCCL finish stays as the other path."""
        from . import abi
        if self.world <= 1:
            return bool(lib.nl_comm_peer_ready())
        mine = C.create_string_buffer(abi.NL_MAILBOX_HANDLE_BYTES)
        ok = lib.nl_comm_mailbox_export(0, mine) == 0
        gathered = self.all_gather_array(np.frombuffer(mine.raw if ok else b"", dtype=np.uint8).copy())
        if any(len(g) != abi.NL_MAILBOX_HANDLE_BYTES for g in gathered):
            return False                      # some rank could not export: everybody stays on NCCL
        blob = b"".join(bytes(g.tobytes()) for g in gathered)
        rc = lib.nl_comm_mailbox_import(C.create_string_buffer(blob, len(blob)))
        oks = self.all_gather_int(1 if rc == 0 else 0)
        return all(oks) and bool(lib.nl_comm_peer_ready())


# ---- host logic of the sharded path (pure functions) -------------------------------
def shard_range(size: int, world: int, rank: int, align: int = 1024) -> Tuple[int, int]:
    """Contiguous shard of [0, size) owned by `rank`: the reference's range partition
    (scheduler.cpp:133-147) with interior boundaries rounded down to `align` elements so every
    shard starts 128-bit aligned.  Mirrors nl_partition()."""
    if align < 1:
        align = 1
    inc = (size // world) // align * align
    start = inc * rank
    end = size if rank == world - 1 else inc * (rank + 1)
    return start, end


def exclusive_scan(counts: Sequence[int]) -> List[int]:
    """Where each shard's compacted run starts in the concatenated result — what the
    memmove gather of compiler.cpp:36-48 computes; len(counts)+1 entries, last = total."""
    out, run = [], 0
    for c in counts:
        out.append(run)
        run += int(c)
    out.append(run)
    return out


def combine_partials(op: str, partials: Sequence, dtype) -> np.generic:
    """Fold per-shard reduction partials in shard order (llvm_finalize_sum_data,
    thunk_methods.cpp:133-151; max/min alike)."""
    dt = np.dtype(dtype)
    arr = np.asarray(partials, dtype=dt)
    if op == "sum":
        acc = dt.type(0)
        with np.errstate(over="ignore"):
            for v in arr:
                acc = dt.type(acc + v)
        return acc
    if op == "max":
        return arr.max() if not (dt.kind == "f" and np.isnan(arr).any()) else dt.type(np.nan)
    if op == "min":
        return arr.min() if not (dt.kind == "f" and np.isnan(arr).any()) else dt.type(np.nan)
    raise ValueError(op)
