"""Second, independent pin of the CPU oracle: the reference's IR emitters restated on llvmlite
(oracle/llvm_ir.py — same function signature, basic blocks, allocas and per-operator instructions as
compiler.cpp:309-538 and the `llvm_generate_*` emitters), JIT-compiled on the host and driven with the
reference's 8-chunk schedule and memmove gather.  CPU only; the closest runnable stand-in for "the
reference's LLVM-JIT CPU path" (SURVEY §8c row 4).  The C restatement must agree with it bit for bit."""
import numpy as np
import pytest

import oracle
from numpyllvm_b200 import abi
from numpyllvm_b200.abi import Expr
from oracle import llvm_ir

pytestmark = pytest.mark.skipif(not llvm_ir.HAVE_LLVMLITE, reason="llvmlite is not installed")

I64 = np.int64


def both(e, ins, n):
    got = oracle.run(e, ins, n)
    ref = llvm_ir.run(e, ins, n)
    assert len(got) == len(ref)
    return got, ref


def arr(g, name):
    return np.array(g["inputs"][name], dtype=I64)


def test_emitted_ir_has_the_reference_skeleton():
    # compiler.cpp:362-366 block names, :340-347 argument order, thunk_as_mapping.cpp:10 "branch"
    e = Expr(I64)
    a = e.array()
    e.materialize(e.filter(a, e.ge(e.ref(a), e.scalar(50))))
    fn, dts, kinds = llvm_ir.compile_pipeline(e)
    text = fn._ir
    for block in ("entry:", "for.cond:", "for.body:", "for.inc:", "for.end:", "branch:"):
        assert block in text
    assert "i64 @\"PipelineFunction0\"(i8** " in text or "i64 @PipelineFunction0(" in text.replace('"', "")
    assert "icmp sge" in text and "icmp slt" in text
    assert kinds == ["filter"]


def test_multiplication_py(golden):
    g = golden["multiplication"]
    e = Expr(I64)
    d_, e_, f_ = e.array(), e.array(), e.array()
    e.materialize(e.mul(e.mul(d_, e_), f_))
    ins = [arr(g, "d"), arr(g, "e"), arr(g, "f")]
    (got,), (ref,) = both(e, ins, 100)
    assert ref.tolist() == g["res"]                      # the JIT-run IR reproduces the golden vector
    assert np.array_equal(got, ref)
    assert "mul i64" in llvm_ir.compile_pipeline(e)[0]._ir


def test_filter_py(golden):
    # Tests/filter.py: c = a[a * 2 >= 50] fused; cardinality from the counters, order stable
    g = golden["filter"]
    a = arr(g, "a")
    e = Expr(I64)
    xa = e.array()
    e.materialize(e.filter(xa, e.ge(e.mul(e.ref(xa), e.scalar(2)), e.scalar(50))))
    (got,), (ref,) = both(e, [a, None, None], a.size)
    assert np.array_equal(ref, a[a * 2 >= 50])
    assert np.array_equal(got, ref)


def test_max_py_filter_sum(golden):
    # Tests/max.py: a[a >= 50].sum(): the 8 partials and their finalize
    g = golden["max"]
    a = arr(g, "a")
    e = Expr(I64)
    xa = e.array()
    e.materialize(e.sum(e.filter(xa, e.ge(e.ref(xa), e.scalar(50)))))
    (got,), (ref,) = both(e, [a, None], a.size)
    assert int(ref[0]) == int(a[a >= 50].sum())
    assert got[0] == ref[0]


def test_assignment_py(golden):
    g = golden["assignment"]
    a, b = arr(g, "a"), arr(g, "b")
    e = Expr(I64)
    e.materialize(e.mul(e.array(), e.array()))
    (got,), (ref,) = both(e, [a, b], a.size)
    assert np.array_equal(ref, a * b) and np.array_equal(got, ref)


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 100, 1000, 4099])
@pytest.mark.parametrize("dtype", [np.int64, np.int32])
def test_random_integer_pipelines(n, dtype):
    rng = np.random.default_rng(n * 7 + np.dtype(dtype).itemsize)
    info = np.iinfo(dtype)
    a = rng.integers(info.min, info.max, size=n, dtype=dtype)      # full range: products wrap
    b = rng.integers(-1000, 1000, size=n, dtype=dtype)
    # held intermediate + root, like C1's fused shape
    e = Expr(dtype)
    xa, xb = e.array(), e.array()
    left = e.mul(xa, xb)
    root = e.sub(e.add(left, e.ref(xa)), e.scalar(3))
    e.materialize(left)
    e.materialize(root)
    got, ref = both(e, [a, b, None], n)
    for x, y in zip(got, ref):
        assert np.array_equal(x, y)
    # every compare, fused filter of a product
    for op in (e.ge, e.gt, e.le, e.lt, e.eq, e.ne):
        f = Expr(dtype)
        ya, yb = f.array(), f.array()
        cmpf = getattr(f, op.__name__)
        f.materialize(f.filter(f.mul(ya, f.scalar(2)), cmpf(yb, f.scalar(17))))
        (g1,), (r1,) = both(f, [a, b, None, None], n)
        assert np.array_equal(g1, r1), op.__name__
    # reductions of a filtered column
    for red in ("sum", "max", "min"):
        f = Expr(dtype)
        ya, yb = f.array(), f.array()
        f.materialize(getattr(f, red)(f.filter(ya, f.ge(yb, f.scalar(0)))))
        (g1,), (r1,) = both(f, [a, b, None], n)
        assert g1[0] == r1[0], red


@pytest.mark.parametrize("n", [1, 9, 1000, 4099])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_random_float_pipelines(n, dtype):
    # the float configs (C1-C5) are an extension of the reference's integer emitters: fmul / fcmp
    rng = np.random.default_rng(n)
    a = rng.random(n).astype(dtype) + 1
    b = rng.random(n).astype(dtype)
    e = Expr(dtype)
    xa, xb = e.array(), e.array()
    e.materialize(e.mul(e.mul(xa, xb), e.ref(xa)))
    (got,), (ref,) = both(e, [a, b], n)
    assert np.array_equal(got, ref)                     # no contraction, no reassociation on either side
    f = Expr(dtype)
    ya = f.array()
    f.materialize(f.filter(ya, f.gt(f.ref(ya), f.scalar(1.5))))   # C4's shape
    (g1,), (r1,) = both(f, [a, None], n)
    assert np.array_equal(g1, r1) and np.array_equal(r1, a[a > dtype(1.5)])
    f = Expr(dtype)
    ya = f.array()
    f.materialize(f.max(f.mul(f.filter(ya, f.ge(f.ref(ya), f.scalar(1.25))), f.scalar(3))))   # C5's shape
    (g1,), (r1,) = both(f, [a, None, None], n)
    assert g1[0] == r1[0]
    f = Expr(dtype)
    f.materialize(f.sum(f.array()))
    (g1,), (r1,) = both(f, [a], n)
    tol = 1e-12 if dtype == np.float64 else 1e-6
    assert abs(float(g1[0]) - float(r1[0])) <= tol * abs(float(r1[0]))


@pytest.mark.parametrize("dtype", [np.int8, np.int16, np.uint8, np.uint16, np.uint32, np.uint64], ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("n", [1, 9, 1000, 4099])
def test_narrow_and_unsigned_columns(n, dtype):
    # getLLVMType maps i8 / i16 (compiler.cpp:156-182) and leaves unsigned "todo": both restatements wrap every
    # operation to the column type, compare unsigned columns unsigned and keep the column type in reductions
    rng = np.random.default_rng(n + np.dtype(dtype).itemsize)
    info = np.iinfo(dtype)
    a = rng.integers(info.min, info.max, size=n, dtype=dtype, endpoint=True)
    b = rng.integers(info.min, info.max, size=n, dtype=dtype, endpoint=True)
    t = int(a[n // 2])
    e = Expr(dtype)
    xa, xb = e.array(), e.array()
    e.materialize(e.sub(e.add(e.mul(xa, xb), e.ref(xa)), e.scalar(3)))
    (got,), (ref,) = both(e, [a, b, None], n)
    assert np.array_equal(got, ref)
    with np.errstate(over="ignore"):
        assert np.array_equal(ref, (a * b + a - dtype(3)).astype(dtype))
    for name in ("ge", "gt", "le", "lt", "eq", "ne"):
        f = Expr(dtype)
        ya, yb = f.array(), f.array()
        f.materialize(f.filter(ya, getattr(f, name)(yb, f.scalar(t))))
        (g1,), (r1,) = both(f, [a, b, None], n)
        assert np.array_equal(g1, r1), name
        assert np.array_equal(r1, a[getattr(np, {"ge": "greater_equal", "gt": "greater", "le": "less_equal", "lt": "less",
                                                "eq": "equal", "ne": "not_equal"}[name])(b, dtype(t))]), name
    for red in ("sum", "max", "min"):
        f = Expr(dtype)
        f.materialize(getattr(f, red)(f.array()))
        (g1,), (r1,) = both(f, [a], n)
        assert g1[0] == r1[0], red
        with np.errstate(over="ignore"):
            want = a.sum(dtype=dtype) if red == "sum" else getattr(a, red)()
        assert r1[0] == want, red
