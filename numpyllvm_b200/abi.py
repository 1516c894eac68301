"""ctypes mirror of include/nl_abi.h — the C-ABI of libnlkernels.

This is the binding a Python host would add on the reference side (see
INTEGRATION.md).  It holds no compute: every entry point forwards to the CUDA
library and raises when the library or a device is missing (there is no CPU
fallback).  The `Expr` builder constructs the reference's per-pipeline
operation tree (parser.hpp:25-87) in the ABI's node format.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

NL_ABI_VERSION = 1
NL_MAX_INPUTS = 8
NL_MAX_OUTPUTS = 4
NL_MAX_NODES = 48
NL_MAX_STEPS = 64
NL_UNIQUE_ID_BYTES = 128
NL_MAILBOX_HANDLE_BYTES = 64

# error codes
NL_OK, NL_ERR_INVALID, NL_ERR_UNSUPPORTED, NL_ERR_SPLIT = 0, -1, -2, -3
NL_ERR_NO_DEVICE, NL_ERR_CUDA, NL_ERR_NCCL, NL_ERR_NOMEM = -4, -5, -6, -7
NL_ERR_SPLIT_FILTER = -8

# dtypes
NL_BOOL, NL_I32, NL_I64, NL_F32, NL_F64 = 0, 1, 2, 3, 4
NL_I8, NL_I16, NL_U8, NL_U16, NL_U32, NL_U64 = 5, 6, 7, 8, 9, 10
_NP2NL = {np.dtype(np.bool_): NL_BOOL, np.dtype(np.int32): NL_I32, np.dtype(np.int64): NL_I64,
          np.dtype(np.float32): NL_F32, np.dtype(np.float64): NL_F64,
          np.dtype(np.int8): NL_I8, np.dtype(np.int16): NL_I16, np.dtype(np.uint8): NL_U8, np.dtype(np.uint16): NL_U16,
          np.dtype(np.uint32): NL_U32, np.dtype(np.uint64): NL_U64}
_NL2NP = {v: k for k, v in _NP2NL.items()}

# ops
OP_INPUT, OP_MUL, OP_ADD, OP_SUB = 0, 1, 2, 3
OP_DIV, OP_FLOORDIV = 4, 5
OP_GE, OP_GT, OP_LE, OP_LT, OP_EQ, OP_NE = 8, 9, 10, 11, 12, 13
OP_AND, OP_OR, OP_XOR, OP_NOT = 16, 17, 18, 19
OP_FILTER = 24
OP_MAX, OP_MIN, OP_SUM = 32, 33, 34
OP_WHERE_STORE = 40
OP_NAMES = {OP_INPUT: "in", OP_MUL: "*", OP_ADD: "+", OP_SUB: "-", OP_DIV: "/", OP_FLOORDIV: "//", OP_GE: ">=", OP_GT: ">",
            OP_LE: "<=", OP_LT: "<", OP_EQ: "==", OP_NE: "!=", OP_AND: "&", OP_OR: "|",
            OP_XOR: "^", OP_NOT: "~", OP_FILTER: "[]", OP_MAX: "max", OP_MIN: "min",
            OP_SUM: "sum", OP_WHERE_STORE: "where="}

IN_ARRAY, IN_SCALAR_IMM, IN_SCALAR_DEV = 0, 1, 2
IDX_LOOP, IDX_COMPACT, IDX_SHARD, IDX_WHERE = 0, 1, 2, 3
PLAN_HAS_FILTER, PLAN_HAS_COMPACT, PLAN_HAS_REDUCE, PLAN_HAS_WHERE = 1, 2, 4, 8


def nl_dtype_of(dt) -> int:
    dt = np.dtype(dt)
    if dt not in _NP2NL:
        raise TypeError(f"dtype {dt} is outside the kernel library (bool, int8..int64, uint8..uint64, float32, float64)")
    return _NP2NL[dt]


def np_dtype_of(nl: int) -> np.dtype:
    return _NL2NP[nl]


class ExprNode(C.Structure):
    _fields_ = [("op", C.c_int32), ("lhs", C.c_int32), ("rhs", C.c_int32),
                ("input", C.c_int32), ("output", C.c_int32)]


class InputDesc(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("kind", C.c_int32), ("imm_bits", C.c_uint64)]


class OutputDesc(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("index_space", C.c_int32)]


class Plan(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("dtype", C.c_int32), ("n_inputs", C.c_int32),
                ("n_outputs", C.c_int32), ("n_steps", C.c_int32), ("flags", C.c_int32),
                ("reduce_op", C.c_int32), ("n_temps", C.c_int32),
                ("in_", InputDesc * NL_MAX_INPUTS), ("out", OutputDesc * NL_MAX_OUTPUTS),
                ("step", C.c_uint32 * NL_MAX_STEPS)]


def scalar_bits(value, dtype) -> int:
    """Bit pattern of `value` converted to `dtype` (getLLVMConstant, compiler.cpp:125-154)."""
    arr = np.array([value]).astype(np.dtype(dtype))
    raw = arr.tobytes() + b"\0" * (8 - arr.itemsize)
    return int.from_bytes(raw, "little")


class Expr:
    """Builds one pipeline's operation tree in the ABI layout (post-order node list).

    >>> e = Expr(np.int64); a = e.array(); r = e.filter(a, e.ge(e.mul(a, e.scalar(2)), e.scalar(50)))
    >>> e.materialize(r)
    """

    def __init__(self, dtype):
        self.dtype = np.dtype(dtype)
        self.nodes: List[ExprNode] = []
        self.inputs: List[InputDesc] = []
        self.n_outputs = 0

    # -- leaves ---------------------------------------------------------
    def _leaf(self, desc: InputDesc) -> int:
        self.inputs.append(desc)
        self.nodes.append(ExprNode(OP_INPUT, -1, -1, len(self.inputs) - 1, -1))
        return len(self.nodes) - 1

    def array(self, dtype=None) -> int:
        return self._leaf(InputDesc(nl_dtype_of(dtype or self.dtype), IN_ARRAY, 0))

    def scalar(self, value) -> int:
        return self._leaf(InputDesc(nl_dtype_of(self.dtype), IN_SCALAR_IMM, scalar_bits(value, self.dtype)))

    def scalar_dev(self, dtype=None) -> int:
        return self._leaf(InputDesc(nl_dtype_of(dtype or self.dtype), IN_SCALAR_DEV, 0))

    def ref(self, node: int) -> int:
        """Second use of an existing leaf: the reference loads the column again
        (each ObjectOperation is its own node, parser.cpp:84-90) but shares the input slot."""
        nd = self.nodes[node]
        assert nd.op == OP_INPUT
        self.nodes.append(ExprNode(OP_INPUT, -1, -1, nd.input, -1))
        return len(self.nodes) - 1

    # -- operators --------------------------------------------------------
    def binop(self, op: int, l: int, r: int) -> int:
        self.nodes.append(ExprNode(op, l, r, -1, -1))
        return len(self.nodes) - 1

    def unop(self, op: int, l: int) -> int:
        self.nodes.append(ExprNode(op, l, -1, -1, -1))
        return len(self.nodes) - 1

    def mul(self, l, r): return self.binop(OP_MUL, l, r)
    def add(self, l, r): return self.binop(OP_ADD, l, r)
    def sub(self, l, r): return self.binop(OP_SUB, l, r)
    def div(self, l, r): return self.binop(OP_DIV, l, r)
    def floordiv(self, l, r): return self.binop(OP_FLOORDIV, l, r)
    def ge(self, l, r): return self.binop(OP_GE, l, r)
    def gt(self, l, r): return self.binop(OP_GT, l, r)
    def le(self, l, r): return self.binop(OP_LE, l, r)
    def lt(self, l, r): return self.binop(OP_LT, l, r)
    def eq(self, l, r): return self.binop(OP_EQ, l, r)
    def ne(self, l, r): return self.binop(OP_NE, l, r)
    def and_(self, l, r): return self.binop(OP_AND, l, r)
    def or_(self, l, r): return self.binop(OP_OR, l, r)
    def xor(self, l, r): return self.binop(OP_XOR, l, r)
    def not_(self, l): return self.unop(OP_NOT, l)
    def filter(self, value, mask): return self.binop(OP_FILTER, value, mask)
    def max(self, l): return self.unop(OP_MAX, l)
    def min(self, l): return self.unop(OP_MIN, l)
    def sum(self, l): return self.unop(OP_SUM, l)
    def where_store(self, value, mask): return self.binop(OP_WHERE_STORE, value, mask)

    def materialize(self, node: int) -> int:
        """Operation::materialize (parser.hpp:37): the node's value is stored to the next output."""
        self.nodes[node].output = self.n_outputs
        self.n_outputs += 1
        return self.n_outputs - 1

    # -- ABI views ----------------------------------------------------------
    def node_array(self):
        return (ExprNode * len(self.nodes))(*self.nodes)

    def input_array(self):
        return (InputDesc * max(1, len(self.inputs)))(*self.inputs)


# ---------------------------------------------------------------------------
_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NL_KERNELS_LIB") or os.path.join(_PKG_DIR, "libnlkernels.so")     # (the override is for A/B builds)
_lib = None


class NLError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libnlkernels error {code}: {msg}")
        self.code = code


_VP = C.c_void_p
_I64P = C.POINTER(C.c_int64)
_SIGNATURES = {
    # name: (restype, argtypes)
    "nl_abi_version": (C.c_int, []),
    "nl_last_error": (C.c_char_p, []),
    "nl_dtype_size": (C.c_size_t, [C.c_int]),
    "nl_plan_compile": (C.c_int, [C.POINTER(ExprNode), C.c_int, C.POINTER(InputDesc), C.c_int, C.POINTER(Plan)]),
    "nl_plan_dump": (C.c_int, [C.POINTER(Plan), C.c_char_p, C.c_size_t]),
    "nl_partition": (C.c_int, [C.c_int64, C.c_int, C.c_int64, C.c_int, _I64P, _I64P]),
    "nl_init": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    "nl_shutdown": (C.c_int, []),
    "nl_device_count": (C.c_int, []),
    "nl_device_info": (C.c_int, [C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.POINTER(C.c_int), C.POINTER(C.c_size_t)]),
    "nl_device_cuda_id": (C.c_int, [C.c_int]),
    "nl_graph_begin": (C.c_int, [C.c_int]),
    "nl_graph_end": (C.c_int, [C.c_int, C.POINTER(_VP)]),
    "nl_graph_launch": (C.c_int, [_VP]),
    "nl_graph_destroy": (C.c_int, [_VP]),
    "nl_malloc": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(_VP)]),
    "nl_free": (C.c_int, [C.c_int, _VP]),
    "nl_pool_trim": (C.c_int, [C.c_int]),
    "nl_memset": (C.c_int, [C.c_int, C.c_int, _VP, C.c_int, C.c_size_t]),
    "nl_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_VP)]),
    "nl_host_free": (C.c_int, [_VP]),
    "nl_memcpy_h2d": (C.c_int, [C.c_int, C.c_int, _VP, _VP, C.c_size_t]),
    "nl_memcpy_d2h": (C.c_int, [C.c_int, C.c_int, _VP, _VP, C.c_size_t]),
    "nl_memcpy_d2d": (C.c_int, [C.c_int, C.c_int, _VP, _VP, C.c_size_t]),
    "nl_memcpy_peer": (C.c_int, [C.c_int, C.c_int, _VP, C.c_int, _VP, C.c_size_t]),
    "nl_stream_sync": (C.c_int, [C.c_int, C.c_int]),
    "nl_device_sync": (C.c_int, [C.c_int]),
    "nl_stream_wait": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "nl_host_is_pinned": (C.c_int, [_VP]),
    "nl_event_create": (C.c_int, [C.c_int, C.POINTER(_VP)]),
    "nl_event_create_sync": (C.c_int, [C.c_int, C.POINTER(_VP)]),
    "nl_event_query": (C.c_int, [_VP]),
    "nl_stream_wait_event": (C.c_int, [C.c_int, C.c_int, _VP]),
    "nl_event_destroy": (C.c_int, [_VP]),
    "nl_event_record": (C.c_int, [_VP, C.c_int]),
    "nl_event_sync": (C.c_int, [_VP]),
    "nl_event_elapsed_ms": (C.c_int, [_VP, _VP, C.POINTER(C.c_float)]),
    "nl_launch_pipeline": (C.c_int, [C.POINTER(Plan), C.POINTER(_VP), C.POINTER(_VP), C.c_int64, C.c_int64,
                                     _VP, C.c_int, C.c_int, C.c_int]),
    "nl_launch_pipeline_allreduce": (C.c_int, [C.POINTER(Plan), C.POINTER(_VP), C.POINTER(_VP), C.c_int64, C.c_int64,
                                               _VP, C.c_int]),
    "nl_launch_pipeline_allgather": (C.c_int, [C.POINTER(Plan), C.POINTER(_VP), C.POINTER(_VP), C.c_int64, C.c_int64,
                                               _VP, _VP, C.c_int]),
    "nl_finalize_partials": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _VP, C.c_int, _VP]),
    "nl_finish_reduce": (C.c_int, [C.c_int, C.c_int, C.POINTER(_VP), C.c_int]),
    "nl_finish_filter": (C.c_int, [C.POINTER(_VP), C.POINTER(_VP), C.c_int]),
    "nl_assign_fill": (C.c_int, [C.c_int, C.c_int, _VP, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_uint64]),
    "nl_assign_copy": (C.c_int, [C.c_int, C.c_int, _VP, C.c_int, C.c_int64, C.c_int64, C.c_int64, _VP]),
    "nl_fill": (C.c_int, [C.c_int, C.c_int, _VP, C.c_int, C.c_int64, C.c_int64, C.c_uint64, C.c_int]),
    "nl_cast": (C.c_int, [C.c_int, C.c_int, _VP, C.c_int, _VP, C.c_int, C.c_int64]),
    "nl_checksum": (C.c_int, [C.c_int, C.c_int, _VP, C.c_size_t, _VP]),
    "nl_sort": (C.c_int, [C.c_int, C.c_int, C.c_int, _VP, _VP, C.c_int64]),
    "nl_sort_sample": (C.c_int, [C.c_int, C.c_int, C.c_int, _VP, C.c_int64, C.c_int, C.POINTER(C.c_uint64)]),
    "nl_sort_split": (C.c_int, [C.c_int, C.c_int, C.c_int, _VP, _VP, C.c_int64, C.POINTER(C.c_uint64), C.c_int, _I64P]),
    "nl_sort_split_count": (C.c_int, [C.c_int, C.c_int, C.c_int, _VP, C.c_int64, C.POINTER(C.c_uint64), C.c_int, _I64P]),
    "nl_sort_split_scatter": (C.c_int, [C.c_int, C.c_int, C.c_int, _VP, C.c_int64, C.POINTER(C.c_uint64), C.c_int, C.POINTER(_VP)]),
    "nl_gather": (C.c_int, [C.c_int, C.c_int, C.c_int, _VP, _VP, C.c_int64, C.POINTER(_VP), C.POINTER(C.c_int64), C.c_int, C.c_int64, _VP]),
    "nl_comm_unique_id": (C.c_int, [_VP]),
    "nl_comm_init": (C.c_int, [_VP, C.c_int, C.c_int]),
    "nl_comm_world_size": (C.c_int, []),
    "nl_comm_destroy": (C.c_int, []),
    "nl_comm_mailbox_export": (C.c_int, [C.c_int, _VP]),
    "nl_comm_mailbox_import": (C.c_int, [_VP]),
    "nl_comm_peer_ready": (C.c_int, []),
    "nl_comm_abort": (C.c_int, [C.c_int]),
    "nl_plan_kernel": (C.c_int, [C.POINTER(Plan), C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_int)]),
    "nl_set_static_kernels": (C.c_int, [C.c_int]),
    "nl_tune": (C.c_int, [C.c_char_p, C.c_int]),
    "nl_launch_count": (C.c_uint64, []),
    "nl_last_launch_info": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def load_library(path: Optional[str] = None):
    """dlopen libnlkernels.so and attach prototypes.  Raises if it is missing:
    the product has no other execution path."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ImportError(f"{p} not found: build it with `python __graft_entry__.py build` "
                          "(numpyllvm_b200 has no CPU fallback)")
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.nl_abi_version() != NL_ABI_VERSION:
        raise ImportError("libnlkernels ABI version mismatch")
    if path is None:
        _lib = lib
    return lib


def check(rc: int) -> int:
    if rc < 0:
        lib = load_library()
        raise NLError(rc, (lib.nl_last_error() or b"").decode())
    return rc


def compile_plan(expr: Expr) -> Plan:
    lib = load_library()
    plan = Plan()
    check(lib.nl_plan_compile(expr.node_array(), len(expr.nodes), expr.input_array(), len(expr.inputs), C.byref(plan)))
    return plan


def dump_plan(plan: Plan) -> str:
    lib = load_library()
    buf = C.create_string_buffer(8192)
    check(lib.nl_plan_dump(C.byref(plan), buf, len(buf)))
    return buf.value.decode()


def plan_kernel(plan: Plan, vec16: bool = True):
    """(kernel name, is_static) the library selects for this plan."""
    lib = load_library()
    buf = C.create_string_buffer(160)
    st = C.c_int()
    check(lib.nl_plan_kernel(C.byref(plan), int(vec16), buf, len(buf), C.byref(st)))
    return buf.value.decode(), bool(st.value)


def void_array(ptrs: Sequence[Optional[int]]):
    return (_VP * max(1, len(ptrs)))(*[_VP(p) if p else _VP(None) for p in ptrs])
