"""The reference's LLVM code generator, restated on llvmlite — TEST INFRASTRUCTURE ONLY.

The reference cannot be built here (CPython 2 + LLVM 3.9 ORC v1), but its hot path IS an
IR emitter, and llvmlite (MCJIT on LLVM 20) is in the image: this module emits, for one pipeline, the same
function the reference's `CompilePipeline` emits — same signature, same basic blocks, same allocas, same
per-operator instructions — JIT-compiles it on the host and drives it the way the scheduler and the gather do.
It is the closest runnable stand-in for "the reference's LLVM-JIT CPU path" and a SECOND, independent oracle:
tests/test_oracle_llvmlite.py checks the C restatement (nl_oracle.c) against it on the reference's own test
vectors and on random pipelines.  Nothing in the product imports it.

What follows the reference (file:line in /root/reference):
  function `i64 f(i8** outputs, i8** inputs, i64 start, i64 end, i64* result_sizes, i64 thread_nr)`
      — argument order of the IR, compiler.cpp:340-347; blocks entry / for.cond / for.body / for.inc / for.end,
      compiler.cpp:362-366; one alloca per column pointer and per scalar argument, compiler.cpp:396-444;
      `index < end` signed, compiler.cpp:452-457; index++ in for.inc, compiler.cpp:478-485; result_sizes[k] =
      the counter the output was stored with (it->index_addr), else 1, compiler.cpp:488-507.
  operands: a column is `load column[index]`; a size-1 source is an immediate (getLLVMConstant),
      compiler.cpp:202-222.  A materialised node stores `result[index] = value`, integer-cast signed when the
      types differ, and records the index counter in use (compiler.cpp:244-266).
  `*`  CreateMul                         thunk_as_number.cpp:45-47
  `>=` CreateICmpSGE                     thunk_as_number.cpp:9-12
  `a[mask]`  CondBr(mask, branch, for.inc); own counter alloca initialised from `start` in the entry block;
             load counter, store counter+1, later stores use the loaded value as index
                                         thunk_as_mapping.cpp:9-29
  `sum` i64 accumulator alloca = 0; acc += l; index := thread_nr     thunk_methods.cpp:114-131
  `max` i64 accumulator alloca = LLONG_MIN; `l > acc` (SGT) -> branch: store l; index := thread_nr
                                         thunk_methods.cpp:70-91
  schedule: 8 chunks `inc = size/8`, the last to `size` (scheduler.cpp:133-147); outputs zero-filled at the
      upper-bound size (scheduler.cpp:114-118), sum's output 8 partials (thunk_methods.cpp:164-173);
      gather: memmove of every chunk's run behind the previous one (compiler.cpp:30-49, with the real item
      size — the reference hard-codes 8); sum finalize adds the partials (thunk_methods.cpp:133-151).

Extensions the reference has no emitter for, emitted as the obvious siblings so that the float and
compare configs can be cross-checked too: fmul/fadd/fsub and `+`/`-` (CreateAdd/CreateSub), the other five compares
(icmp/fcmp ordered; `!=` unordered), f64 accumulators for float sum/max, `min`.  `max` here is the INTENDED one
(partials combined with max), since the reference's is unfinished.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from numpyllvm_b200 import abi

try:
    import llvmlite.binding as llvm
    from llvmlite import ir
    HAVE_LLVMLITE = True
except Exception:  # pragma: no cover - the image has llvmlite; gate for boxes that do not
    HAVE_LLVMLITE = False

_inited = False
_engines = []   # keep every MCJIT engine alive as long as its function pointers are


def _init():
    global _inited
    if not _inited:
        try:
            llvm.initialize()
        except Exception:
            pass   # newer llvmlite initialises on import
        llvm.initialize_native_target()
        llvm.initialize_native_asmprinter()
        _inited = True


_I64_MIN = -(1 << 63)


def _llvm_type(dt: np.dtype):
    """getLLVMType (compiler.cpp:156-182): bool/i8 -> i8, i16, i32, i64, f32, f64."""
    dt = np.dtype(dt)
    if dt == np.bool_:
        return ir.IntType(8)
    if dt.kind in "iu":
        return ir.IntType(8 * dt.itemsize)
    return ir.FloatType() if dt.itemsize == 4 else ir.DoubleType()


class _Info:
    """JITInformation (compiler.hpp): the cursor the emitters move."""
    def __init__(self):
        self.builder = None
        self.function = None
        self.loop_inc = None
        self.index = None        # Value used as subscript of the next store
        self.index_addr = None   # alloca it was loaded from
        self.thread_addr = None


def compile_pipeline(expr: abi.Expr):
    """CompilePipeline (compiler.cpp:309-538) for one abi.Expr.  Returns (ctypes function, list of output dtypes,
    list of 'kind' per output: 'index' | 'filter' | 'reduce')."""
    _init()
    dt = expr.dtype
    is_float = dt.kind == "f"
    col_t = _llvm_type(dt)
    i8, i64, i1 = ir.IntType(8), ir.IntType(64), ir.IntType(1)
    i8pp = i8.as_pointer().as_pointer()
    module = ir.Module(name="PipelineFunction")
    fnty = ir.FunctionType(i64, [i8pp, i8pp, i64, i64, i64.as_pointer(), i64])
    fn = ir.Function(module, fnty, name="PipelineFunction0")
    outputs_arg, inputs_arg, start_arg, end_arg, sizes_arg, thread_arg = fn.args
    entry = fn.append_basic_block("entry")
    cond = fn.append_basic_block("for.cond")
    body = fn.append_basic_block("for.body")
    inc = fn.append_basic_block("for.inc")
    end = fn.append_basic_block("for.end")
    b = ir.IRBuilder(entry)
    info = _Info()
    info.builder, info.function, info.loop_inc = b, fn, inc

    nodes, ins = expr.nodes, expr.inputs
    n_out = expr.n_outputs
    out_dtype: List[Optional[np.dtype]] = [None] * n_out
    out_kind: List[str] = ["index"] * n_out
    for nd in nodes:
        if nd.output >= 0:
            is_bool = abi.OP_GE <= nd.op <= abi.OP_NOT
            out_dtype[nd.output] = np.dtype(np.bool_) if is_bool else dt

    # entry: one alloca per output column pointer, per input column pointer, then start/end/result_sizes/thread_nr
    out_addr = []
    for k in range(n_out):
        t = _llvm_type(out_dtype[k]).as_pointer()
        vp = b.load(b.gep(outputs_arg, [ir.Constant(i64, k)]), name="voidptr")
        a = b.alloca(t, name=f"outputs{k}")
        b.store(b.bitcast(vp, t), a)
        out_addr.append(a)
    in_addr = []
    for k, desc in enumerate(ins):
        if desc.kind != abi.IN_ARRAY:
            in_addr.append(None)
            continue
        t = _llvm_type(abi.np_dtype_of(desc.dtype)).as_pointer() if hasattr(abi, "np_dtype_of") else col_t.as_pointer()
        vp = b.load(b.gep(inputs_arg, [ir.Constant(i64, k)]), name="voidptr")
        a = b.alloca(t, name=f"inputs{k}")
        b.store(b.bitcast(vp, t), a)
        in_addr.append(a)
    start_addr = b.alloca(i64, name="start"); b.store(start_arg, start_addr)
    end_addr = b.alloca(i64, name="end"); b.store(end_arg, end_addr)
    sizes_addr = b.alloca(i64.as_pointer(), name="result_sizes"); b.store(sizes_arg, sizes_addr)
    thread_addr = b.alloca(i64, name="thread_nr"); b.store(thread_arg, thread_addr)
    info.index_addr, info.thread_addr = start_addr, thread_addr

    # PerformInitialization (compiler.cpp:185-200): post-order over the tree, still in the entry block
    extra = {}
    acc_t = ir.DoubleType() if is_float else i64
    for idx, nd in enumerate(nodes):   # the node list is post-order already
        if nd.op == abi.OP_FILTER:
            a = b.alloca(i64)
            b.store(b.load(info.index_addr), a)                 # thunk_as_mapping.cpp:23-29
            extra[idx] = a
        elif nd.op == abi.OP_SUM:
            a = b.alloca(acc_t)
            b.store(ir.Constant(acc_t, 0.0 if is_float else 0), a)   # thunk_methods.cpp:124-131
            extra[idx] = a
        elif nd.op in (abi.OP_MAX, abi.OP_MIN):
            a = b.alloca(acc_t)
            if is_float:
                init = float("-inf") if nd.op == abi.OP_MAX else float("inf")
            elif dt.kind == "u":
                init = 0 if nd.op == abi.OP_MAX else (1 << 64) - 1         # unsigned columns (extension): unsigned order
            else:
                init = _I64_MIN if nd.op == abi.OP_MAX else (1 << 63) - 1   # thunk_methods.cpp:83-91 (LLONG_MIN)
            b.store(ir.Constant(acc_t, init), a)
            extra[idx] = a
    b.branch(cond)

    b.position_at_end(cond)
    index = b.load(start_addr, name="index")
    endv = b.load(end_addr, name="end")
    b.cbranch(b.icmp_signed("<", index, endv), body, end)

    b.position_at_end(body)
    info.index = b.load(start_addr, name="index")
    info.index_addr = start_addr
    out_index_addr = [None] * n_out

    cmp_sym = {abi.OP_GE: ">=", abi.OP_GT: ">", abi.OP_LE: "<=", abi.OP_LT: "<", abi.OP_EQ: "==", abi.OP_NE: "!="}

    def widen(v):
        if is_float:
            return b.fpext(v, acc_t) if v.type != acc_t else v
        if v.type == i64:
            return v
        return b.zext(v, i64) if dt.kind == "u" else b.sext(v, i64)

    def perform(idx):
        """PerformOperation (compiler.cpp:202-269)."""
        nd = nodes[idx]
        if nd.op == abi.OP_INPUT:
            desc = ins[nd.input]
            if desc.kind == abi.IN_SCALAR_IMM:                  # getLLVMConstant (compiler.cpp:125-154)
                raw = int(desc.imm_bits).to_bytes(8, "little")[: dt.itemsize]
                val = np.frombuffer(raw, dtype=dt)[0]
                v = ir.Constant(col_t, float(val) if is_float else int(val))
            else:
                colp = b.load(in_addr[nd.input])
                v = b.load(b.gep(colp, [info.index], name="column[i]"))
        elif nd.op in (abi.OP_MUL, abi.OP_ADD, abi.OP_SUB):
            l, r = perform(nd.lhs), perform(nd.rhs)
            if is_float:
                v = {abi.OP_MUL: b.fmul, abi.OP_ADD: b.fadd, abi.OP_SUB: b.fsub}[nd.op](l, r)
            else:
                v = {abi.OP_MUL: b.mul, abi.OP_ADD: b.add, abi.OP_SUB: b.sub}[nd.op](l, r)   # thunk_as_number.cpp:45-47
        elif nd.op in cmp_sym:
            l, r = perform(nd.lhs), perform(nd.rhs)
            if is_float:
                v = b.fcmp_unordered("!=", l, r) if nd.op == abi.OP_NE else b.fcmp_ordered(cmp_sym[nd.op], l, r)
            elif dt.kind == "u":
                v = b.icmp_unsigned(cmp_sym[nd.op], l, r)
            else:
                v = b.icmp_signed(cmp_sym[nd.op], l, r)          # thunk_as_number.cpp:9-12 (ICmpSGE)
        elif nd.op == abi.OP_FILTER:
            l, r = perform(nd.lhs), perform(nd.rhs)
            branch = fn.append_basic_block("branch")            # thunk_as_mapping.cpp:9-21
            b.cbranch(r, branch, inc)
            b.position_at_end(branch)
            cnt = b.load(extra[idx], name="index_branch")
            b.store(b.add(cnt, ir.Constant(i64, 1), name="index++"), extra[idx])
            info.index, info.index_addr = cnt, extra[idx]
            v = l
        elif nd.op == abi.OP_SUM:
            l = perform(nd.lhs)
            cur = b.load(extra[idx], name="current_sum")        # thunk_methods.cpp:114-122
            upd = (b.fadd if is_float else b.add)(cur, widen(l), name="sum")
            b.store(upd, extra[idx])
            info.index, info.index_addr = b.load(thread_addr), thread_addr
            v = upd
        elif nd.op in (abi.OP_MAX, abi.OP_MIN):
            l = perform(nd.lhs)
            branch = fn.append_basic_block("branch")            # thunk_methods.cpp:70-81
            cur = b.load(extra[idx], name="current_max")
            lw = widen(l)
            sym = ">" if nd.op == abi.OP_MAX else "<"
            if is_float:
                better = b.fcmp_ordered(sym, lw, cur)
            elif dt.kind == "u":
                better = b.icmp_unsigned(sym, lw, cur)
            else:
                better = b.icmp_signed(sym, lw, cur)
            b.cbranch(better, branch, inc)
            b.position_at_end(branch)
            b.store(lw, extra[idx])
            info.index, info.index_addr = b.load(thread_addr), thread_addr
            v = b.load(extra[idx], name="current_max")
        else:
            raise NotImplementedError(f"no emitter for op {nd.op}")
        if nd.output >= 0:                                      # compiler.cpp:244-266
            k = nd.output
            t = _llvm_type(out_dtype[k])
            resp = b.load(out_addr[k])
            p = b.gep(resp, [info.index], name="result[i]")
            sv = v
            if sv.type != t:
                if isinstance(sv.type, ir.IntType) and isinstance(t, ir.IntType):
                    if sv.type.width == 1:
                        sv = b.zext(sv, t)     # deliberate departure: 0x01, not the signed cast's 0xFF
                    elif sv.type.width > t.width:
                        sv = b.trunc(sv, t)
                    else:
                        sv = b.sext(sv, t)
                else:
                    sv = b.fptrunc(sv, t)      # f64 accumulator -> f32 column
            b.store(sv, p)
            out_index_addr[k] = info.index_addr
            out_kind[k] = "reduce" if nd.op in (abi.OP_SUM, abi.OP_MAX, abi.OP_MIN) else (
                "filter" if info.index_addr is not start_addr else "index")
        return v

    root = max(i for i in range(len(nodes)))   # the root is the last node of the post-order list
    perform(root)
    b.branch(inc)

    b.position_at_end(inc)
    idxv = b.load(start_addr, name="index")
    b.store(b.add(idxv, ir.Constant(i64, 1), name="index++"), start_addr)
    b.branch(cond)

    b.position_at_end(end)
    sizes = b.load(sizes_addr, name="result_sizes[]")
    for k in range(n_out):
        cnt = b.load(out_index_addr[k], name="count") if out_index_addr[k] is not None else ir.Constant(i64, 1)
        b.store(cnt, b.gep(sizes, [ir.Constant(i64, k)]))
    b.ret(ir.Constant(i64, 0))

    text = str(module)
    mod = llvm.parse_assembly(text)
    mod.verify()
    target = llvm.Target.from_default_triple().create_target_machine(opt=2)
    engine = llvm.create_mcjit_compiler(mod, target)
    engine.finalize_object()
    _engines.append(engine)
    addr = engine.get_function_address("PipelineFunction0")
    VPP = C.POINTER(C.c_void_p)
    cfn = C.CFUNCTYPE(C.c_int64, VPP, VPP, C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.c_int64)(addr)
    cfn._ir = text
    return cfn, out_dtype, out_kind


def run(expr: abi.Expr, inputs: Sequence[Optional[np.ndarray]], size: int, thread_count: int = 8) -> List[np.ndarray]:
    """ScheduleFunction + ExecuteFunction + JITFunctionDECREF around the JIT-compiled loop: zero-filled outputs,
    `thread_count` range chunks, the memmove gather, the reduction finalize."""
    fn, out_dtype, out_kind = compile_pipeline(expr)
    root_op = expr.nodes[-1].op
    ins = [None if a is None else np.ascontiguousarray(a) for a in inputs]
    outs = []
    for k in range(expr.n_outputs):
        n_el = thread_count if out_kind[k] == "reduce" else max(size, 1)
        if out_kind[k] == "reduce":
            acc_dt = np.float64 if expr.dtype.kind == "f" else expr.dtype
            if root_op == abi.OP_SUM:
                o = np.zeros(n_el, dtype=out_dtype[k])                       # llvm_initialize_sum_data
            else:   # the identity a chunk without any hit leaves behind (intended max/min)
                ident = (-np.inf if root_op == abi.OP_MAX else np.inf) if expr.dtype.kind == "f" else (
                    np.iinfo(expr.dtype).min if root_op == abi.OP_MAX else np.iinfo(expr.dtype).max)   # (signed and unsigned alike)
                o = np.full(n_el, ident, dtype=out_dtype[k])
            del acc_dt
        else:
            o = np.zeros(n_el, dtype=out_dtype[k])                           # scheduler.cpp:114-118
        outs.append(o)
    in_arr = (C.c_void_p * max(1, len(ins)))(*[C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(None) for a in ins])
    out_arr = (C.c_void_p * max(1, len(outs)))(*[C.c_void_p(o.ctypes.data) for o in outs])
    starts, ends = [], []
    inc = size // thread_count                                               # scheduler.cpp:133-147
    for j in range(thread_count):
        s = j * inc
        e = size if j == thread_count - 1 else (j + 1) * inc
        sizes = (C.c_int64 * max(1, len(outs)))()
        rc = fn(out_arr, in_arr, s, e, sizes, j)              # outputs first: the call site, compiler.cpp:87
        assert rc == 0
        starts.append(s)
        ends.append([sizes[k] for k in range(len(outs))])
    res = []
    for k, o in enumerate(outs):
        if out_kind[k] == "reduce":
            if root_op == abi.OP_SUM:
                if expr.dtype.kind == "f":
                    res.append(np.array([o.astype(np.float64).sum()], dtype=o.dtype))   # finalize: j = 0..7 in order
                else:
                    with np.errstate(over="ignore"):
                        tot = o.dtype.type(0)
                        for j in range(thread_count):
                            tot = o.dtype.type(tot + o[j])
                    res.append(np.array([tot], dtype=o.dtype))
            else:
                res.append(np.array([o.max() if root_op == abi.OP_MAX else o.min()], dtype=o.dtype))
            continue
        # gather (compiler.cpp:30-49): chunk j's run [start_j, end_j) moves behind chunk j-1's
        current = ends[0][k]
        for j in range(1, thread_count):
            n_el = ends[j][k] - starts[j]
            if n_el > 0 and current != starts[j]:
                o[current:current + n_el] = o[starts[j]:starts[j] + n_el].copy()
            current += max(n_el, 0)
        res.append(o[:current])
    return res
