"""CPU-only checks of the C-ABI library: it loads, exports every symbol the header
declares, and its host-only entry points (plan lowering, dump, partition) behave.
No compute call is made — without a GPU the compute entry points must fail loudly."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from numpyllvm_b200 import abi
from numpyllvm_b200.abi import Expr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "nl_abi.h")).read()
    return sorted(set(re.findall(r"NL_API\s+[^;(]*?\b(nl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(abi.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in nl_abi.h but not exported"
    assert set(syms) == set(abi.EXPORTED_SYMBOLS), "abi.py prototypes out of sync with nl_abi.h"


def test_struct_layout_matches_header():
    # sizes are fixed by the C declaration; a drift here would corrupt plans silently
    assert C.sizeof(abi.ExprNode) == 20 and C.sizeof(abi.InputDesc) == 16 and C.sizeof(abi.OutputDesc) == 8
    assert C.sizeof(abi.Plan) == 8 * 4 + 16 * abi.NL_MAX_INPUTS + 8 * abi.NL_MAX_OUTPUTS + 4 * abi.NL_MAX_STEPS


def test_no_cpu_fallback_without_device():
    lib = abi.load_library()
    if lib.nl_device_count() > 0:
        pytest.skip("a device is initialised in this process")
    rc = lib.nl_init(0, None)
    if rc == 0:
        pytest.skip("GPU present")
    assert rc == abi.NL_ERR_NO_DEVICE and b"no CPU fallback" in lib.nl_last_error()
    e = Expr(np.float32)
    e.materialize(e.mul(e.array(), e.array()))
    plan = abi.compile_plan(e)
    sizes = (C.c_int64 * 1)()
    rc = lib.nl_launch_pipeline(C.byref(plan), abi.void_array([1]), abi.void_array([1, 1]), 0, 16,
                                C.cast(sizes, C.c_void_p), 0, 0, 0)
    assert rc == abi.NL_ERR_NO_DEVICE
    p = C.c_void_p()
    assert lib.nl_malloc(0, 64, C.byref(p)) == abi.NL_ERR_NO_DEVICE


def test_partition_reference_and_aligned():
    lib = abi.load_library()
    s, e = C.c_int64(), C.c_int64()
    got = []
    for j in range(8):
        assert lib.nl_partition(100, 8, 1, j, C.byref(s), C.byref(e)) == 0
        got.append((s.value, e.value))
    assert got == [(0, 12), (12, 24), (24, 36), (36, 48), (48, 60), (60, 72), (72, 84), (84, 100)]   # scheduler.cpp:133-147
    # aligned variant: interior boundaries multiples of 1024, full coverage, no overlap
    n, g = (1 << 30) + 12345, 8
    prev = 0
    for j in range(g):
        assert lib.nl_partition(n, g, 1024, j, C.byref(s), C.byref(e)) == 0
        assert s.value == prev and (s.value % 1024 == 0)
        prev = e.value
    assert prev == n
    assert lib.nl_partition(10, 0, 1, 0, C.byref(s), C.byref(e)) == abi.NL_ERR_INVALID


def steps_of(plan):
    return abi.dump_plan(plan).splitlines()


def test_lowering_filter_takes_no_temporary_for_a_reread_column():
    # Tests/filter.py: a[a*2 >= 50].  The column is needed before and after the filter; a plan that needs no
    # value temporary for the expression itself does not take one to park the column (the re-read is an
    # L1/L2 hit, and the super-tile compaction kernel re-reads its rows from L2 by design)
    e = Expr(np.int64)
    a = e.array()
    e.materialize(e.filter(a, e.ge(e.mul(e.ref(a), e.scalar(2)), e.scalar(50))))
    plan = abi.compile_plan(e)
    text = abi.dump_plan(plan)
    assert plan.flags == abi.PLAN_HAS_FILTER | abi.PLAN_HAS_COMPACT
    assert plan.n_temps == 0 and text.count("acc = in0") == 2 and "out0[compact] = acc" in text
    assert plan.out[0].index_space == abi.IDX_COMPACT
    # a[a > t]: the accumulator still holds the column when the filter has run — one load
    e = Expr(np.float64)
    a = e.array()
    e.materialize(e.filter(a, e.gt(e.ref(a), e.scalar(0.5))))
    assert abi.dump_plan(abi.compile_plan(e)).count("acc = in0") == 1
    # an expression that needs a temporary anyway may park a column in the other one
    e = Expr(np.int64)
    a, b = e.array(), e.array()
    e.materialize(e.filter(e.add(e.mul(e.ref(a), b), e.mul(e.ref(a), e.ref(a))), e.gt(a, e.scalar(3))))
    assert abi.compile_plan(e).n_temps >= 1


def test_lowering_six_way_product_uses_one_temp():
    e = Expr(np.float64)
    d, ee, f, a, b, c = (e.array() for _ in range(6))
    e.materialize(e.mul(e.mul(e.mul(d, ee), f), e.mul(e.mul(a, b), c)))
    plan = abi.compile_plan(e)
    assert plan.n_temps == 1 and plan.n_steps == 9 and plan.flags == 0


def test_lowering_filter_reduce_has_no_compaction():
    # (a[a >= t] * k).max(): the filter is a predicate on the reduction, nothing is compacted
    e = Expr(np.float32)
    a = e.array()
    e.materialize(e.max(e.mul(e.filter(a, e.ge(e.ref(a), e.scalar(0.5))), e.scalar(3.0))))
    plan = abi.compile_plan(e)
    assert plan.flags == abi.PLAN_HAS_FILTER | abi.PLAN_HAS_REDUCE
    assert plan.reduce_op == abi.OP_MAX and plan.out[0].index_space == abi.IDX_SHARD


def test_lowering_rejects_what_the_library_cannot_run():
    lib = abi.load_library()
    # array loaded after a filter (compiler.cpp:213 would index it with the compacted counter)
    e = Expr(np.int64)
    a, b = e.array(), e.array()
    e.materialize(e.mul(e.filter(a, e.ge(e.ref(a), e.scalar(1))), b))
    with pytest.raises(abi.NLError) as ei:
        abi.compile_plan(e)
    assert ei.value.code == abi.NL_ERR_SPLIT_FILTER      # legal once the host gives the filter its own pipeline
    # mixed column dtypes
    e = Expr(np.int64)
    e.materialize(e.mul(e.array(np.int64), e.array(np.int32)))
    with pytest.raises(abi.NLError):
        abi.compile_plan(e)
    # reduction below the root
    e = Expr(np.int64)
    e.materialize(e.mul(e.sum(e.array()), e.array()))
    with pytest.raises(abi.NLError):
        abi.compile_plan(e)
    # unmaterialised root
    e = Expr(np.int64)
    e.mul(e.array(), e.array())
    with pytest.raises(abi.NLError):
        abi.compile_plan(e)


def test_lowering_deep_tree_asks_for_a_split():
    # a balanced tree needs more value temporaries than the register file has
    e = Expr(np.float32)

    def tree(depth):
        if depth == 0:
            return e.mul(e.array(), e.scalar(2.0))   # non-leaf so both sides need evaluating
        return e.add(tree(depth - 1), tree(depth - 1))
    root = tree(2)          # 4 arrays + 4 scalars = 8 inputs, needs 2 temporaries: fits
    e.materialize(root)
    plan = abi.compile_plan(e)
    assert plan.n_temps == 2
    # depth 4 over two shared columns: more live partial results than temporaries
    e3 = Expr(np.float32)
    a, b = e3.array(), e3.array()

    def tree3(depth):
        if depth == 0:
            return e3.mul(e3.ref(a), e3.ref(b))
        return e3.add(tree3(depth - 1), tree3(depth - 1))
    e3.materialize(tree3(4))
    with pytest.raises(abi.NLError) as ei:
        abi.compile_plan(e3)
    assert ei.value.code == abi.NL_ERR_SPLIT
    e2 = Expr(np.float32)
    xs = [e2.array() for _ in range(9)]
    acc = xs[0]
    for x in xs[1:]:
        acc = e2.mul(acc, x)
    e2.materialize(acc)
    with pytest.raises(abi.NLError) as ei:
        abi.compile_plan(e2)
    assert ei.value.code == abi.NL_ERR_SPLIT     # 9 inputs > NL_MAX_INPUTS


def test_reversed_operands_keep_non_commutative_order():
    e = Expr(np.float64)
    x = e.array()
    e.materialize(e.sub(e.scalar(1.0), e.mul(x, e.scalar(2.0))))     # 1 - x*2
    text = abi.dump_plan(abi.compile_plan(e))
    assert "rsub" in text
    e = Expr(np.float64)
    x = e.array()
    e.materialize(e.lt(e.scalar(1.0), e.mul(x, e.scalar(2.0))))      # 1 < x*2  ==  x*2 > 1
    text = abi.dump_plan(abi.compile_plan(e))
    assert "pacc = gt(acc" in text


def test_benchmark_shapes_resolve_to_build_time_kernels():
    # kernel-template selection: the pipelines of BASELINE.json's configs and of the reference's
    # tests match a shape instantiated at build time; anything else runs the interpreter
    def kernel(e):
        return abi.plan_kernel(abi.compile_plan(e))
    e = Expr(np.float32); e.materialize(e.mul(e.array(), e.array()))
    assert kernel(e) == ("pipeline_kernel<f32,V=4,static:mul2>", True)                  # C2 / Tests/assignment.py
    e = Expr(np.float64); d, ee, f, a, b, c = (e.array() for _ in range(6))
    e.materialize(e.mul(e.mul(e.mul(d, ee), f), e.mul(e.mul(a, b), c)))
    assert kernel(e)[0].endswith("static:mul6>")                                        # C1
    for dt in (np.float64, np.int64):
        for op in ("max", "min", "sum"):
            e = Expr(dt); e.materialize(getattr(e, op)(e.array()))
            assert kernel(e)[0].endswith(f"static:reduce_{op}>")                        # C3
    e = Expr(np.float64); xa = e.array(); e.materialize(e.filter(xa, e.gt(e.ref(xa), e.scalar(0.5))))
    assert kernel(e)[0].endswith("static:filter_gt>")                                   # C4
    e = Expr(np.float32); xa = e.array()
    e.materialize(e.max(e.mul(e.filter(xa, e.ge(e.ref(xa), e.scalar(0.5))), e.scalar(3.0))))
    assert kernel(e)[0].endswith("static:filter_ge_mul_max>")                           # C5
    e = Expr(np.int64); xa = e.array()
    e.materialize(e.filter(xa, e.ge(e.mul(e.ref(xa), e.scalar(2)), e.scalar(50))))
    assert kernel(e)[0].endswith("static:filter_mul_ge>")                               # Tests/filter.py
    e = Expr(np.int64); xa = e.array(); e.materialize(e.sum(e.filter(xa, e.ge(e.ref(xa), e.scalar(50)))))
    assert kernel(e)[0].endswith("static:filter_ge_sum>")                               # Tests/max.py
    # second tier: shape FAMILIES (nl_family.cuh) — the step kinds are build-time, operators and inputs run-time
    e = Expr(np.int64); xa = e.array(); e.materialize(e.add(e.mul(xa, e.scalar(3)), e.scalar(1)))
    assert kernel(e) == ("pipeline_kernel<i64,V=2,family:ew2>", True)                     # a*3 + 1
    e = Expr(np.float64); xa, xb = e.array(), e.array(); e.materialize(e.sub(e.div(xa, xb), e.scalar(0.5)))
    assert kernel(e)[0].endswith("family:ew2>")                                          # a/b - 0.5: same family, other operators
    e = Expr(np.float32); a, b, c, d = (e.array() for _ in range(4)); e.materialize(e.add(e.mul(a, b), e.mul(c, d)))
    assert kernel(e) == ("pipeline_kernel<f32,V=4,family:ew1x1>", True)                   # a*b + c*d
    e = Expr(np.float64); xa = e.array()
    e.materialize(e.filter(e.add(e.mul(xa, e.scalar(2.0)), e.scalar(1.0)), e.gt(e.ref(xa), e.scalar(0.9))))
    assert kernel(e)[0].endswith("family:fc_v2>")                                        # (a*2+1)[a > t]: the accumulator still holds a
    e = Expr(np.float64); xa, xb, xc = e.array(), e.array(), e.array()
    e.materialize(e.filter(e.sub(xa, xb), e.ne(xc, e.scalar(0.0))))
    assert kernel(e)[0].endswith("family:fc_l_v1c>")                                     # (a-b)[c != 0]: a COLUMN operand is part of the shape
    e = Expr(np.int32); xa, xb = e.array(), e.array(); e.materialize(e.sum(e.mul(xa, xb)))
    assert kernel(e)[0].endswith("family:red1>")                                         # (a*b).sum()
    e = Expr(np.float32); xa = e.array()
    e.materialize(e.min(e.add(e.filter(xa, e.lt(e.ref(xa), e.scalar(0.5))), e.scalar(1.0))))
    assert kernel(e)[0].endswith("family:fr_v1>")                                        # (a[a < t] + 1).min()
    # third tier, the step interpreter: floor division (a long out-of-line sequence no family carries) ...
    e = Expr(np.int64); xa = e.array(); e.materialize(e.floordiv(xa, e.scalar(3)))
    name, is_static = kernel(e)
    assert not is_static and "dynamic" in name
    # ... shapes outside the lists ...
    e = Expr(np.float32); cols = [e.array() for _ in range(8)]
    acc = cols[0]
    for c in cols[1:]:
        acc = e.add(acc, c)
    e.materialize(acc)
    assert kernel(e) == ("pipeline_kernel<f32,V=4,compact=0,reduce=0,dynamic-light>", False)
    # ... and scalar rows (misaligned operands): the light variant when the plan has no value temporary, the general one otherwise
    e = Expr(np.float32); e.materialize(e.mul(e.array(), e.array()))
    assert abi.plan_kernel(abi.compile_plan(e), vec16=False) == ("pipeline_kernel<f32,V=1,compact=0,reduce=0,dynamic-light>", False)
    e = Expr(np.float32); a, b, c, d = (e.array() for _ in range(4)); e.materialize(e.add(e.mul(a, b), e.mul(c, d)))
    assert abi.plan_kernel(abi.compile_plan(e), vec16=False) == ("pipeline_kernel<f32,V=1,compact=0,reduce=0,dynamic>", False)
    # a column read twice is not parked in a temporary when that would be the plan's only one
    e = Expr(np.float32); x = e.array(); e.materialize(e.mul(e.add(e.ref(x), e.scalar(np.float32(1))), e.ref(x)))
    assert abi.compile_plan(e).n_temps == 0
