"""Thin Pythonic wrappers over the C-ABI (device buffers, pipeline launches, events).

Host plumbing only: every call forwards to libnlkernels.  Used by the C-ABI
level tests and by bench.py; the user-facing surface is the `jit` extension.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import abi


class DevArray:
    """A device allocation on one local device (owned)."""

    def __init__(self, rt: "Runtime", dev: int, dtype, size: int):
        self.rt, self.dev, self.dtype, self.size = rt, dev, np.dtype(dtype), int(size)
        self.nbytes = self.size * self.dtype.itemsize
        p = C.c_void_p()
        abi.check(rt.lib.nl_malloc(dev, max(self.nbytes, 16), C.byref(p)))
        self.ptr = p.value

    def free(self):
        if self.ptr:
            self.rt.lib.nl_free(self.dev, C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Runtime:
    _instance: Optional["Runtime"] = None

    def __init__(self, n_devices: int = 0):
        self.lib = abi.load_library()
        abi.check(self.lib.nl_init(n_devices, None))
        self.n_devices = self.lib.nl_device_count()

    @classmethod
    def get(cls, n_devices: int = 0) -> "Runtime":
        if cls._instance is None:
            cls._instance = cls(n_devices)
        return cls._instance

    # -- memory ------------------------------------------------------------
    def empty(self, dtype, size: int, dev: int = 0) -> DevArray:
        return DevArray(self, dev, dtype, size)

    def upload(self, arr: np.ndarray, dev: int = 0, stream: int = 0, offset_elems: int = 0) -> DevArray:
        arr = np.ascontiguousarray(arr)
        d = DevArray(self, dev, arr.dtype, arr.size + offset_elems)
        if arr.size:
            abi.check(self.lib.nl_memcpy_h2d(dev, stream, C.c_void_p(d.ptr + offset_elems * arr.dtype.itemsize),
                                             C.c_void_p(arr.ctypes.data), arr.nbytes))
            abi.check(self.lib.nl_stream_sync(dev, stream))
        return d

    def download(self, d: DevArray, count: Optional[int] = None, stream: int = 0, offset_elems: int = 0) -> np.ndarray:
        n = d.size - offset_elems if count is None else int(count)
        out = np.empty(n, dtype=d.dtype)
        if n:
            abi.check(self.lib.nl_memcpy_d2h(d.dev, stream, C.c_void_p(out.ctypes.data),
                                             C.c_void_p(d.ptr + offset_elems * d.dtype.itemsize), out.nbytes))
        abi.check(self.lib.nl_stream_sync(d.dev, stream))
        return out

    def fill(self, d: DevArray, first_index: int, seed: int, kind: int, stream: int = 0, count: Optional[int] = None):
        abi.check(self.lib.nl_fill(d.dev, stream, C.c_void_p(d.ptr), abi.nl_dtype_of(d.dtype), first_index,
                                   d.size if count is None else count, seed, kind))

    def sync(self, dev: int = 0, stream: int = 0):
        abi.check(self.lib.nl_stream_sync(dev, stream))

    # -- events --------------------------------------------------------------
    def event(self, dev: int = 0):
        ev = C.c_void_p()
        abi.check(self.lib.nl_event_create(dev, C.byref(ev)))
        return ev

    def record(self, ev, stream: int = 0):
        abi.check(self.lib.nl_event_record(ev, stream))

    def elapsed_ms(self, e0, e1) -> float:
        abi.check(self.lib.nl_event_sync(e1))
        ms = C.c_float()
        abi.check(self.lib.nl_event_elapsed_ms(e0, e1, C.byref(ms)))
        return float(ms.value)

    # -- the hot path ----------------------------------------------------------
    def launch(self, plan: abi.Plan, outputs: Sequence[int], inputs: Sequence[Optional[int]], start: int, end: int,
               sizes_ptr: int, thread_nr: int = 0, dev: int = 0, stream: int = 0):
        abi.check(self.lib.nl_launch_pipeline(C.byref(plan), abi.void_array(outputs), abi.void_array(inputs),
                                              start, end, C.c_void_p(sizes_ptr), thread_nr, dev, stream))

    def launch_allreduce(self, plan: abi.Plan, outputs: Sequence[int], inputs: Sequence[Optional[int]], start: int,
                         end: int, sizes_ptr: int, dev: int = 0):
        """Reduction pipeline with the multi-GPU finish fused into the kernel (NVLink mailboxes)."""
        abi.check(self.lib.nl_launch_pipeline_allreduce(C.byref(plan), abi.void_array(outputs), abi.void_array(inputs),
                                                        start, end, C.c_void_p(sizes_ptr), dev))

    def launch_allgather(self, plan: abi.Plan, outputs: Sequence[int], inputs: Sequence[Optional[int]], start: int,
                         end: int, sizes_ptr: int, offsets_ptr: int, dev: int = 0):
        """Filter pipeline whose kernel also exchanges the shard counts (NVLink mailboxes) and leaves
        their exclusive scan in `offsets` (world + 1 int64 on the device)."""
        abi.check(self.lib.nl_launch_pipeline_allgather(C.byref(plan), abi.void_array(outputs), abi.void_array(inputs),
                                                        start, end, C.c_void_p(sizes_ptr), C.c_void_p(offsets_ptr), dev))

    def sort(self, d: "DevArray", stream: int = 0):
        """In-place device radix sort of a column (the sort() full breaker)."""
        tmp = DevArray(self, d.dev, d.dtype, max(d.size, 1))
        abi.check(self.lib.nl_sort(d.dev, stream, abi.nl_dtype_of(d.dtype), C.c_void_p(d.ptr), C.c_void_p(tmp.ptr), d.size))
        abi.check(self.lib.nl_stream_sync(d.dev, stream))
        tmp.free()

    def peer_ready(self) -> bool:
        return bool(self.lib.nl_comm_peer_ready())

    def last_launch(self):
        name = C.create_string_buffer(128)
        g, b, v = C.c_int(), C.c_int(), C.c_int()
        self.lib.nl_last_launch_info(name, len(name), C.byref(g), C.byref(b), C.byref(v))
        return name.value.decode(), g.value, b.value, v.value


def output_layout(expr: abi.Expr, plan: abi.Plan):
    return [(abi.np_dtype_of(plan.out[k].dtype), plan.out[k].index_space) for k in range(plan.n_outputs)]


def run_pipeline(rt: Runtime, expr: abi.Expr, inputs: Sequence[Optional[np.ndarray]], size: int, dev: int = 0,
                 inplace: Optional[dict] = None, start: int = 0, misalign_elems: int = 0,
                 thread_nr: int = 0, n_slots: int = 1) -> List[np.ndarray]:
    """Run one pipeline on one device through the C-ABI and fetch its outputs, trimmed to
    their cardinality — the device twin of oracle.run().  `inplace` maps an output index to
    the INPUT index whose buffer it aliases (masked assignment)."""
    plan = abi.compile_plan(expr)
    dins: List[Optional[DevArray]] = []
    for k, a in enumerate(inputs):
        if a is None or plan.in_[k].kind == abi.IN_SCALAR_IMM:
            dins.append(None)
        elif plan.in_[k].kind == abi.IN_SCALAR_DEV:
            dins.append(rt.upload(np.asarray(a).reshape(1), dev))
        else:
            dins.append(rt.upload(a, dev, offset_elems=misalign_elems))
    layout = output_layout(expr, plan)
    douts: List[DevArray] = []
    for k, (dt, space) in enumerate(layout):
        if space == abi.IDX_WHERE:
            douts.append(dins[inplace[k]])
        elif space == abi.IDX_SHARD:
            douts.append(rt.empty(dt, n_slots, dev))
        else:
            douts.append(rt.empty(dt, max(size + misalign_elems, 1), dev))
    sizes = rt.empty(np.int64, max(1, plan.n_outputs), dev)

    def ptr(d, dt):
        return None if d is None else d.ptr
    in_ptrs = [None if d is None else d.ptr for d in dins]
    out_ptrs = [d.ptr for d in douts]
    rt.launch(plan, out_ptrs, in_ptrs, start + misalign_elems, size + misalign_elems, sizes.ptr, thread_nr, dev)
    counts = rt.download(sizes)
    res = []
    for k, (dt, space) in enumerate(layout):
        if space == abi.IDX_SHARD:
            res.append(rt.download(douts[k])[thread_nr:thread_nr + 1])
        elif space == abi.IDX_COMPACT:
            n = int(counts[k]) - (start + misalign_elems)
            res.append(rt.download(douts[k], n, offset_elems=start + misalign_elems))
        else:
            n = int(counts[k]) - (start + misalign_elems)
            res.append(rt.download(douts[k], n, offset_elems=start + misalign_elems))
    return res
