"""GPU parity against the SECOND oracle: the reference's IR emitters restated on llvmlite and JIT-run on the host
(oracle/llvm_ir.py).  The five configs' pipeline shapes, through the C-ABI, bit for bit (float sums: stated rtol)."""
import numpy as np
import pytest

import oracle
from numpyllvm_b200.abi import Expr
from numpyllvm_b200.device import Runtime, run_pipeline
from oracle import llvm_ir

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not llvm_ir.HAVE_LLVMLITE, reason="llvmlite is not installed")]

SIZES = [1, 1023, 4097, 100003, (1 << 20) + 5]


@pytest.fixture(scope="module")
def rt(nl):
    return Runtime.get()


def same(ref, got):
    assert len(ref) == len(got)
    for r, g in zip(ref, got):
        assert r.dtype == g.dtype and r.shape == g.shape
        assert np.array_equal(r.view(np.uint8), g.view(np.uint8))


@pytest.mark.parametrize("n", SIZES)
def test_c1_six_way_product(rt, n):
    cols = [oracle.fill(np.float64, 0, n, 10 + k, 2) for k in range(6)]
    e = Expr(np.float64)
    d_, e_, f_, a_, b_, c_ = (e.array() for _ in range(6))
    e.materialize(e.mul(e.mul(e.mul(d_, e_), f_), e.mul(e.mul(a_, b_), c_)))
    same(llvm_ir.run(e, cols, n), run_pipeline(rt, e, cols, n))


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("dtype", [np.float32, np.int64])
def test_c2_product(rt, n, dtype):
    a, b = oracle.fill(dtype, 0, n, 1, 2 if dtype == np.float32 else 1), oracle.fill(dtype, 0, n, 2, 2 if dtype == np.float32 else 1)
    e = Expr(dtype)
    e.materialize(e.mul(e.array(), e.array()))
    same(llvm_ir.run(e, [a, b], n), run_pipeline(rt, e, [a, b], n))


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("dtype", [np.float64, np.int64])
@pytest.mark.parametrize("red", ["max", "sum"])
def test_c3_reductions(rt, n, dtype, red):
    a = oracle.fill(dtype, 0, n, 3, 3 if dtype == np.float64 else 1)
    e = Expr(dtype)
    e.materialize(getattr(e, red)(e.array()))
    (ref,), (got,) = llvm_ir.run(e, [a], n), run_pipeline(rt, e, [a], n)
    if dtype == np.float64 and red == "sum":
        assert abs(float(got[0]) - float(ref[0])) <= 1e-12 * abs(float(ref[0]))   # north_star: rtol 1e-12 for f64 sums
    else:
        assert got[0] == ref[0]


@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("sel", [0.01, 0.5, 0.9])
def test_c4_filter(rt, n, sel):
    a = oracle.fill(np.float64, 0, n, 4, 3)
    e = Expr(np.float64)
    xa = e.array()
    e.materialize(e.filter(xa, e.gt(e.ref(xa), e.scalar(1.0 - sel))))
    same(llvm_ir.run(e, [a, None], n), run_pipeline(rt, e, [a, None], n))


@pytest.mark.parametrize("n", SIZES)
def test_c5_filter_mul_max(rt, n):
    a = oracle.fill(np.float32, 0, n, 5, 3)
    e = Expr(np.float32)
    xa = e.array()
    e.materialize(e.max(e.mul(e.filter(xa, e.ge(e.ref(xa), e.scalar(0.5))), e.scalar(3.0))))
    (ref,), (got,) = llvm_ir.run(e, [a, None, None], n), run_pipeline(rt, e, [a, None, None], n)
    assert got[0] == ref[0]


@pytest.mark.parametrize("n", SIZES)
def test_reference_tests_filter_sum_int64(rt, n):
    # Tests/max.py's real chain a[a >= t].sum() and Tests/filter.py's a[a*2 >= 50], int64 in [1,100)
    a = oracle.fill(np.int64, 0, n, 42, 0)
    e = Expr(np.int64)
    xa = e.array()
    e.materialize(e.sum(e.filter(xa, e.ge(e.ref(xa), e.scalar(50)))))
    (ref,), (got,) = llvm_ir.run(e, [a, None], n), run_pipeline(rt, e, [a, None], n)
    assert got[0] == ref[0]
    e = Expr(np.int64)
    xa = e.array()
    e.materialize(e.filter(xa, e.ge(e.mul(e.ref(xa), e.scalar(2)), e.scalar(50))))
    same(llvm_ir.run(e, [a, None, None], n), run_pipeline(rt, e, [a, None, None], n))


@pytest.mark.parametrize("n", [1023, 100003])
@pytest.mark.parametrize("dtype", [np.int8, np.int16, np.uint8, np.uint16, np.uint32, np.uint64, np.int32], ids=lambda d: np.dtype(d).name)
def test_narrow_and_unsigned_columns(rt, n, dtype):
    # the column types getLLVMType maps (i8, i16, i32) and the unsigned ones it left "todo"
    rng = np.random.default_rng(n)
    info = np.iinfo(dtype)
    a = rng.integers(info.min, info.max, size=n, dtype=dtype, endpoint=True)
    b = rng.integers(info.min, info.max, size=n, dtype=dtype, endpoint=True)
    e = Expr(dtype)
    xa, xb = e.array(), e.array()
    e.materialize(e.sub(e.add(e.mul(xa, xb), e.ref(xa)), e.scalar(3)))
    same(llvm_ir.run(e, [a, b, None], n), run_pipeline(rt, e, [a, b, None], n))
    f = Expr(dtype)
    ya, yb = f.array(), f.array()
    f.materialize(f.filter(ya, f.gt(yb, f.scalar(int(a[n // 2])))))
    same(llvm_ir.run(f, [a, b, None], n), run_pipeline(rt, f, [a, b, None], n))
    for red in ("sum", "max", "min"):
        g = Expr(dtype)
        g.materialize(getattr(g, red)(g.array()))
        (ref,), (got,) = llvm_ir.run(g, [a], n), run_pipeline(rt, g, [a], n)
        assert got[0] == ref[0], red
