"""CPU oracle for the thunk-pipeline path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package, and only as the checker or the timed
CPU baseline.  The product (numpyllvm_b200) never imports it.
See oracle/nl_oracle.h for the parity-pinning status.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

from numpyllvm_b200 import abi  # ABI struct layouts only (the mirror of nl_abi.h)

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "libnloracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement when a source is newer than the library.  Several ranks of one bench run may get
    here at the same moment (a header touched after the last build): the check and the build sit behind a file lock,
    and make writes to a private name that is renamed into place, so nobody ever loads a half-written file."""
    import fcntl
    srcs = [os.path.join(_DIR, f) for f in ("nl_oracle.c", "nl_oracle_eval.inc", "nl_oracle.h")]
    srcs.append(os.path.join(_DIR, "..", "include", "nl_abi.h"))

    def stale():
        return (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs if os.path.exists(s))
    if not (force or stale()):
        return _SO
    with open(os.path.join(_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or stale():
                tmp = f"libnloracle.{os.getpid()}.tmp.so"
                subprocess.run(["make", "-C", _DIR, "-B", "-s", tmp], check=True, stdout=subprocess.DEVNULL)
                os.replace(os.path.join(_DIR, tmp), _SO)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        VP, I64 = C.c_void_p, C.c_int64
        _lib.nlo_run.restype = C.c_int
        _lib.nlo_run.argtypes = [C.POINTER(abi.ExprNode), C.c_int, C.POINTER(abi.InputDesc), C.c_int,
                                 C.POINTER(VP), C.POINTER(VP), C.c_int, I64, C.c_int, C.c_int, C.POINTER(I64)]
        _lib.nlo_fill.restype = C.c_int
        _lib.nlo_fill.argtypes = [VP, C.c_int, I64, I64, C.c_uint64, C.c_int]
        _lib.nlo_partition.restype = None
        _lib.nlo_partition.argtypes = [I64, C.c_int, C.c_int, C.POINTER(I64), C.POINTER(I64)]
        _lib.nlo_assign_fill.restype = C.c_int
        _lib.nlo_assign_fill.argtypes = [VP, C.c_int, I64, I64, I64, C.c_uint64]
        _lib.nlo_assign_copy.restype = C.c_int
        _lib.nlo_assign_copy.argtypes = [VP, C.c_int, I64, I64, I64, VP]
        _lib.nlo_fast_mul2_f32.argtypes = [VP, VP, VP, I64, C.c_int, C.c_int]
        _lib.nlo_fast_mul6_f64.argtypes = [C.POINTER(VP), VP, I64, C.c_int, C.c_int]
        _lib.nlo_fast_filter_gt_f64.argtypes = [VP, C.c_double, VP, I64, C.c_int, C.c_int, C.POINTER(I64)]
        _lib.nlo_fast_max_f64.argtypes = [VP, I64, C.c_int, C.c_int, C.POINTER(C.c_double)]
        _lib.nlo_fast_max_i64.argtypes = [VP, I64, C.c_int, C.c_int, C.POINTER(I64)]
        _lib.nlo_fast_sum_i64.argtypes = [VP, I64, C.c_int, C.c_int, C.POINTER(I64)]
        _lib.nlo_fast_filter_mul_max_f32.argtypes = [VP, C.c_float, C.c_float, I64, C.c_int, C.c_int, C.POINTER(C.c_float)]
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(None)


def output_layout(expr: abi.Expr):
    """(dtype, is_reduce) of each materialised output, in output order."""
    outs = {}
    for nd in expr.nodes:
        if nd.output >= 0:
            is_bool = (abi.OP_GE <= nd.op <= abi.OP_NOT) or (
                nd.op == abi.OP_INPUT and expr.inputs[nd.input].dtype == abi.NL_BOOL)
            outs[nd.output] = (np.dtype(np.bool_) if is_bool else expr.dtype,
                               nd.op in (abi.OP_MAX, abi.OP_MIN, abi.OP_SUM), nd.op == abi.OP_WHERE_STORE)
    return [outs[k] for k in range(expr.n_outputs)]


def run(expr: abi.Expr, inputs: Sequence[Optional[np.ndarray]], size: int, thread_count: int = 8,
        n_workers: int = 1, inplace: Optional[dict] = None) -> List[np.ndarray]:
    """Evaluate one pipeline the reference's way; returns each output trimmed to its
    final cardinality.  `inplace` maps output index -> ndarray updated in place
    (NL_OP_WHERE_STORE targets)."""
    L = lib()
    ins = [None if a is None else np.ascontiguousarray(a) for a in inputs]
    layout = output_layout(expr)
    outs = []
    for k, (dt, is_red, is_where) in enumerate(layout):
        if is_where:
            outs.append(inplace[k])
        else:
            outs.append(np.empty(1 if is_red else max(size, 1), dtype=dt))
    in_arr = (C.c_void_p * max(1, len(ins)))(*[_ptr(a) for a in ins])
    out_arr = (C.c_void_p * max(1, len(outs)))(*[_ptr(a) for a in outs])
    counts = (C.c_int64 * max(1, len(outs)))()
    rc = L.nlo_run(expr.node_array(), len(expr.nodes), expr.input_array(), len(expr.inputs),
                   in_arr, out_arr, len(outs), size, thread_count, n_workers, counts)
    if rc != 0:
        raise RuntimeError(f"oracle nlo_run failed: {rc}")
    return [o[: counts[k]] if not layout[k][2] else o for k, o in enumerate(outs)]


def fill(dtype, first_index: int, count: int, seed: int = 42, kind: int = 0) -> np.ndarray:
    out = np.empty(count, dtype=dtype)
    rc = lib().nlo_fill(_ptr(out), abi.nl_dtype_of(dtype), first_index, count, seed, kind)
    if rc != 0:
        raise RuntimeError(f"nlo_fill failed: {rc}")
    return out


def partition(size: int, thread_count: int, j: int):
    s, e = C.c_int64(), C.c_int64()
    lib().nlo_partition(size, thread_count, j, C.byref(s), C.byref(e))
    return s.value, e.value


def assign_fill(dst: np.ndarray, lo: int, step: int, count: int, value) -> None:
    lib().nlo_assign_fill(_ptr(dst), abi.nl_dtype_of(dst.dtype), lo, step, count, abi.scalar_bits(value, dst.dtype))


def assign_copy(dst: np.ndarray, lo: int, step: int, count: int, src: np.ndarray) -> None:
    src = np.ascontiguousarray(src, dtype=dst.dtype)
    lib().nlo_assign_copy(_ptr(dst), abi.nl_dtype_of(dst.dtype), lo, step, count, _ptr(src))


# ---- fixed-shape compiled loops (the timed CPU baseline) --------------------
def fast_mul2_f32(a, b, c, thread_count=8, n_workers=1):
    lib().nlo_fast_mul2_f32(_ptr(a), _ptr(b), _ptr(c), a.size, thread_count, n_workers)
    return c


def fast_mul6_f64(ins6, out, thread_count=8, n_workers=1):
    arr = (C.c_void_p * 6)(*[_ptr(a) for a in ins6])
    lib().nlo_fast_mul6_f64(arr, _ptr(out), out.size, thread_count, n_workers)
    return out


def fast_filter_gt_f64(a, t, out, thread_count=8, n_workers=1):
    cnt = C.c_int64()
    lib().nlo_fast_filter_gt_f64(_ptr(a), float(t), _ptr(out), a.size, thread_count, n_workers, C.byref(cnt))
    return out[: cnt.value]


def fast_max_f64(a, thread_count=8, n_workers=1):
    r = C.c_double()
    lib().nlo_fast_max_f64(_ptr(a), a.size, thread_count, n_workers, C.byref(r))
    return r.value


def fast_max_i64(a, thread_count=8, n_workers=1):
    r = C.c_int64()
    lib().nlo_fast_max_i64(_ptr(a), a.size, thread_count, n_workers, C.byref(r))
    return r.value


def fast_sum_i64(a, thread_count=8, n_workers=1):
    r = C.c_int64()
    lib().nlo_fast_sum_i64(_ptr(a), a.size, thread_count, n_workers, C.byref(r))
    return r.value


def fast_filter_mul_max_f32(a, t, k, thread_count=8, n_workers=1):
    r = C.c_float()
    lib().nlo_fast_filter_mul_max_f32(_ptr(a), float(t), float(k), a.size, thread_count, n_workers, C.byref(r))
    return r.value
