"""Pins the CPU oracle against the reference's own tests (Tests/*.py) — CPU only.

Each case rebuilds the pipeline the reference's parser would hand to the
compiler for that script and checks the oracle's output against the committed
golden vector (tests/golden/reference_tests.json) AND the known answers
tabulated in SURVEY.md §4.
"""
import numpy as np
import pytest

import oracle
from numpyllvm_b200 import abi
from numpyllvm_b200.abi import Expr

I64 = np.int64


def arr(g, name):
    return np.array(g["inputs"][name], dtype=I64)


def test_golden_file_matches_generator(golden, tmp_path):
    # the committed fixture is what the committed script produces
    import importlib.util, json, os
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    a, b = mod.draw(2)
    assert a.tolist() == golden["assignment"]["inputs"]["a"]
    assert a[:10].tolist() == [52, 93, 15, 72, 61, 21, 83, 87, 75, 75] and a.max() == 95


def test_multiplication_res(golden):
    # Tests/multiplication.py:26  res = d*e*f   (held by Python -> materialised)
    g = golden["multiplication"]
    e = Expr(I64)
    d_, e_, f_ = e.array(), e.array(), e.array()
    e.materialize(e.mul(e.mul(d_, e_), f_))
    (res,) = oracle.run(e, [arr(g, "d"), arr(g, "e"), arr(g, "f")], 100)
    assert res.tolist() == g["res"]
    assert res[:5].tolist() == [91800, 35088, 101574, 8613, 80496] and res.sum() == 13095033


def test_multiplication_res2(golden):
    # Tests/multiplication.py:27  res2 = res.sort() * (a*b*c): the parser makes the sorted
    # result and (a*b*c) child pipelines (parser.cpp:99-145); the parent multiplies two
    # materialised inputs.  sort is the NumPy base function (thunk_methods.cpp:37-42).
    g = golden["multiplication"]
    e1 = Expr(I64)
    a_, b_, c_ = e1.array(), e1.array(), e1.array()
    e1.materialize(e1.mul(e1.mul(a_, b_), c_))
    (abc,) = oracle.run(e1, [arr(g, "a"), arr(g, "b"), arr(g, "c")], 100)
    sorted_res = np.sort(np.array(g["res"], dtype=I64))
    e2 = Expr(I64)
    e2.materialize(e2.mul(e2.array(), e2.array()))
    (res2,) = oracle.run(e2, [sorted_res, abc], 100)
    assert res2.tolist() == g["res2"]
    assert res2[:5].tolist() == [38095200, 195286608, 14908320, 43338240, 115948800]
    assert int(res2.sum()) == 1695327314761


def test_multiplication_fused_six(golden):
    # C1's fused shape (d*e*f)*(a*b*c) with two held intermediates -> 3 outputs in one loop
    g = golden["multiplication"]
    e = Expr(I64)
    d_, e_, f_, a_, b_, c_ = (e.array() for _ in range(6))
    left = e.mul(e.mul(d_, e_), f_)
    right = e.mul(e.mul(a_, b_), c_)
    root = e.mul(left, right)
    e.materialize(left)
    e.materialize(root)
    ins = [arr(g, k) for k in "defabc"]
    left_v, root_v = oracle.run(e, ins, 100)
    assert left_v.tolist() == g["res"]
    assert root_v.tolist() == g["unsorted_product"] and int(root_v.sum()) == 2066956057037


def test_assignment(golden):
    # Tests/assignment.py:14-20: c = a*b; a[0] = 1000 forces c first (thunk.cpp:81-88), so c
    # sees the original a; then NumPy setitem on the storage.
    g = golden["assignment"]
    a, b = arr(g, "a"), arr(g, "b")
    e = Expr(I64)
    e.materialize(e.mul(e.array(), e.array()))
    storage = a.copy()                         # PyThunk_FromArray copies (thunk.cpp:30)
    (c,) = oracle.run(e, [storage, b], 100)
    oracle.assign_fill(storage, g["assign"]["index"], 1, 1, g["assign"]["value"])
    assert storage[0] == 1000 and a[0] == 52
    assert c.tolist() == g["c"]
    assert c[:5].tolist() == [3120, 3813, 435, 1080, 2745] and c.sum() == 249755


def filter_expr():
    e = Expr(I64)
    a_ = e.array()
    mask = e.ge(e.mul(e.ref(a_), e.scalar(2)), e.scalar(50))
    e.materialize(e.filter(a_, mask))
    return e


def test_filter(golden):
    # Tests/filter.py:13  res = a[a*2 >= 50]
    g = golden["filter"]
    (res,) = oracle.run(filter_expr(), [arr(g, "a"), None, None], 100)
    assert res.tolist() == g["res"]
    assert len(res) == 73 and res.sum() == 4749
    assert res[:5].tolist() == [52, 93, 72, 61, 83] and res[-5:].tolist() == [80, 82, 53, 26, 89]


def test_filter_chunk_counts(golden):
    # the per-chunk runs the gather stitches together (compiler.cpp:36-46)
    g = golden["filter"]
    a = arr(g, "a")
    counts = []
    for j in range(8):
        s, en = oracle.partition(100, 8, j)
        (r,) = oracle.run(filter_expr(), [a[s:en], None, None], en - s, thread_count=1)
        counts.append(len(r))
    assert counts == g["chunk_counts"] == [9, 7, 10, 9, 7, 9, 9, 13]


def test_partition_matches_reference():
    # scheduler.cpp:133-147: increment = size/8, last chunk to size
    assert [oracle.partition(100, 8, j) for j in range(8)] == \
        [(0, 12), (12, 24), (24, 36), (36, 48), (48, 60), (60, 72), (72, 84), (84, 100)]
    assert oracle.partition(5, 8, 7) == (0, 5) and oracle.partition(5, 8, 0) == (0, 0)


def test_max_py_filter_sum(golden):
    # Tests/max.py:13-14: res = a[a >= 50].sum(); res2 = res * a  (res is a size-1 child
    # pipeline result folded into the parent as a constant, compiler.cpp:210-211)
    g = golden["max"]
    a = arr(g, "a")
    e = Expr(I64)
    a_ = e.array()
    e.materialize(e.sum(e.filter(a_, e.ge(e.ref(a_), e.scalar(50)))))
    (res,) = oracle.run(e, [a, None], 100)
    assert res.tolist() == [g["res"]] == [4117]
    e2 = Expr(I64)
    e2.materialize(e2.mul(e2.scalar(int(res[0])), e2.array()))
    (res2,) = oracle.run(e2, [None, a], 100)
    assert res2.tolist() == g["res2"]
    assert res2[:5].tolist() == [214084, 382881, 61755, 296424, 251137]


def test_max_py_chunk_partials(golden):
    g = golden["max"]
    a = arr(g, "a")
    partials = []
    for j in range(8):
        s, en = oracle.partition(100, 8, j)
        e = Expr(I64)
        a_ = e.array()
        e.materialize(e.sum(e.filter(a_, e.ge(e.ref(a_), e.scalar(50)))))
        (r,) = oracle.run(e, [a[s:en], None], en - s, thread_count=1)
        partials.append(int(r[0]))
    assert partials == g["chunk_partials"] == [686, 265, 605, 480, 468, 340, 490, 783]


def test_filtered_materialised_and_summed(golden):
    # res held by Python AND summed: one loop, a compacted output and a partial output
    g = golden["max"]
    a = arr(g, "a")
    e = Expr(I64)
    a_ = e.array()
    f = e.filter(a_, e.ge(e.ref(a_), e.scalar(50)))
    e.materialize(f)
    e.materialize(e.sum(f))
    filt, s = oracle.run(e, [a, None], 100)
    assert filt.tolist() == g["filtered"] and len(filt) == 57 and s[0] == 4117


@pytest.mark.parametrize("workers", [1, 4, 8])
def test_workers_do_not_change_results(golden, workers):
    g = golden["filter"]
    (res,) = oracle.run(filter_expr(), [arr(g, "a"), None, None], 100, n_workers=workers)
    assert res.tolist() == g["res"]


# ---- unpinned-by-the-reference semantics: pinned against NumPy -----------------
@pytest.mark.parametrize("dtype", [np.int32, np.int64, np.float32, np.float64])
@pytest.mark.parametrize("n", [0, 1, 7, 8, 100, 1000, 4099])
def test_reductions_vs_numpy(dtype, n):
    rng = np.random.default_rng(n + 7)
    x = (rng.integers(-1000, 1000, n) if np.issubdtype(dtype, np.integer) else rng.standard_normal(n)).astype(dtype)
    for op, ref in (("max", np.max), ("min", np.min), ("sum", np.sum)):
        e = Expr(dtype)
        e.materialize(getattr(e, op)(e.array()))
        (r,) = oracle.run(e, [x], n)
        if n == 0 and op != "sum":
            lowest = -np.inf if np.issubdtype(dtype, np.floating) else np.iinfo(dtype).min
            highest = np.inf if np.issubdtype(dtype, np.floating) else np.iinfo(dtype).max
            assert r[0] == (lowest if op == "max" else highest)
        elif op == "sum" and np.issubdtype(dtype, np.floating):
            exact = float(np.sum(x.astype(np.float64)))
            assert abs(float(r[0]) - exact) <= (1e-12 if dtype == np.float64 else 1e-6) * max(1.0, np.abs(x).sum())
        else:
            assert r[0] == ref(x)


def test_int_wraparound():
    x = oracle.fill(np.int64, 0, 1000, seed=42, kind=1)
    e = Expr(np.int64)
    a_ = e.array()
    e.materialize(e.mul(e.mul(a_, e.ref(a_)), e.scalar(3)))
    (r,) = oracle.run(e, [x, None], 1000)
    with np.errstate(over="ignore"):
        assert np.array_equal(r, x * x * 3)
    e = Expr(np.int64)
    e.materialize(e.sum(e.array()))
    (s,) = oracle.run(e, [x], 1000)
    with np.errstate(over="ignore"):
        assert s[0] == x.sum()


def test_nan_max_is_sticky():
    x = np.array([1.0, np.nan, 3.0, 2.0] * 5)
    e = Expr(np.float64)
    e.materialize(e.max(e.array()))
    (r,) = oracle.run(e, [x], len(x))
    assert np.isnan(r[0]) and np.isnan(np.max(x))


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32])
def test_compares_and_bool_ops_vs_numpy(dtype):
    rng = np.random.default_rng(3)
    x = rng.integers(0, 10, 1000).astype(dtype)
    y = rng.integers(0, 10, 1000).astype(dtype)
    for name, f in (("ge", np.greater_equal), ("gt", np.greater), ("le", np.less_equal),
                    ("lt", np.less), ("eq", np.equal), ("ne", np.not_equal)):
        e = Expr(dtype)
        e.materialize(getattr(e, name)(e.array(), e.array()))
        (r,) = oracle.run(e, [x, y], 1000)
        assert r.dtype == np.bool_ and np.array_equal(r, f(x, y))
    e = Expr(dtype)
    xa, ya = e.array(), e.array()
    m = e.and_(e.gt(xa, e.scalar(2)), e.not_(e.ge(ya, e.scalar(7))))
    e.materialize(e.filter(e.add(e.ref(xa), e.ref(ya)), m))
    (r,) = oracle.run(e, [x, y, None, None], 1000)
    assert np.array_equal(r, (x + y)[(x > 2) & ~(y >= 7)])


def test_where_store_in_place():
    x = np.arange(100, dtype=np.float32)
    target = x.copy()
    e = Expr(np.float32)
    a_ = e.array()
    e.materialize(e.where_store(e.scalar(-1.0), e.gt(a_, e.scalar(49.5))))
    oracle.run(e, [target, None, None], 100, inplace={0: target})
    expect = x.copy(); expect[x > 49.5] = -1.0
    assert np.array_equal(target, expect)


def test_fill_is_deterministic_and_chunkable():
    whole = oracle.fill(np.float64, 0, 1000, seed=42, kind=3)
    parts = np.concatenate([oracle.fill(np.float64, s, 250, seed=42, kind=3) for s in range(0, 1000, 250)])
    assert np.array_equal(whole, parts) and whole.min() >= 0.0 and whole.max() < 1.0
    ints = oracle.fill(np.int64, 0, 10000, seed=42, kind=0)
    assert ints.min() >= 1 and ints.max() <= 99
    f32 = oracle.fill(np.float32, 0, 1000, seed=42, kind=2)
    assert f32.min() >= 1.0 and f32.max() < 2.0


def test_fast_loops_match_generic_oracle():
    n = 10007
    a = oracle.fill(np.float32, 0, n, 1, 2); b = oracle.fill(np.float32, 0, n, 2, 2)
    e = Expr(np.float32); e.materialize(e.mul(e.array(), e.array()))
    (c,) = oracle.run(e, [a, b], n)
    assert np.array_equal(oracle.fast_mul2_f32(a, b, np.empty_like(a), n_workers=4), c)

    ins = [oracle.fill(np.float64, 0, n, s, 2) for s in range(6)]
    e = Expr(np.float64)
    d_, e_, f_, a_, b_, c_ = (e.array() for _ in range(6))
    e.materialize(e.mul(e.mul(e.mul(d_, e_), f_), e.mul(e.mul(a_, b_), c_)))
    (o,) = oracle.run(e, ins, n)
    # in6 order is {d,e,f,a,b,c}; product is ((a*b)*c) * ((d*e)*f) == commuted root (fmul commutes exactly)
    assert np.array_equal(oracle.fast_mul6_f64(ins, np.empty(n)), o)

    x = oracle.fill(np.float64, 0, n, 9, 3)
    e = Expr(np.float64); xa = e.array()
    e.materialize(e.filter(xa, e.gt(e.ref(xa), e.scalar(0.5))))
    (f,) = oracle.run(e, [x, None], n)
    assert np.array_equal(oracle.fast_filter_gt_f64(x, 0.5, np.empty(n), n_workers=8), f)

    e = Expr(np.float64); e.materialize(e.max(e.array()))
    assert oracle.fast_max_f64(x) == oracle.run(e, [x], n)[0][0] == x.max()
    xi = oracle.fill(np.int64, 0, n, 5, 1)
    e = Expr(np.int64); e.materialize(e.max(e.array()))
    assert oracle.fast_max_i64(xi) == oracle.run(e, [xi], n)[0][0] == xi.max()
    e = Expr(np.int64); e.materialize(e.sum(e.array()))
    assert oracle.fast_sum_i64(xi, n_workers=3) == oracle.run(e, [xi], n)[0][0]

    xf = oracle.fill(np.float32, 0, n, 11, 3)
    e = Expr(np.float32); xa = e.array()
    e.materialize(e.max(e.mul(e.filter(xa, e.ge(e.ref(xa), e.scalar(0.25))), e.scalar(3.0))))
    (m,) = oracle.run(e, [xf, None, None], n)
    assert oracle.fast_filter_mul_max_f32(xf, 0.25, 3.0) == m[0] == (xf[xf >= 0.25] * np.float32(3.0)).max()


# ---- division and negation: extensions on the operator slots, pinned by NumPy -----
def _div_columns(dtype, n=4001):
    rng = np.random.default_rng(12)
    if np.dtype(dtype).kind == "i":
        info = np.iinfo(dtype)
        a = rng.integers(-1000, 1000, n).astype(dtype)
        b = rng.integers(-9, 10, n).astype(dtype)          # contains zeros
        a[:4] = [info.min, info.min, info.max, 0]
        b[:4] = [-1, 1, -1, 0]
    else:
        a = ((rng.random(n) - 0.5) * 100).astype(dtype)
        b = ((rng.random(n) - 0.5) * 7).astype(dtype)
        a[:8] = [1.0, -1.0, 0.0, -0.0, 7.5, np.inf, np.nan, 1e30]
        b[:8] = [0.0, 0.0, 0.0, 3.0, -2.5, 2.0, 1.0, 1e-30]
        a[8:12] = [0.3, 0.9, 6.0, -6.0]
        b[8:12] = [0.1, 0.3, 0.1, 0.1]                      # quotients that round onto an integer
    return a, b


@pytest.mark.parametrize("dtype", [np.int32, np.int64, np.float32, np.float64], ids=lambda d: np.dtype(d).name)
def test_floor_division_is_numpy_floor_divide(dtype):
    a, b = _div_columns(dtype)
    e = Expr(dtype)
    e.materialize(e.floordiv(e.array(), e.array()))
    (got,) = oracle.run(e, [a, b], a.size)
    with np.errstate(all="ignore"):
        want = np.floor_divide(a, b)
    assert np.array_equal(got, want, equal_nan=np.dtype(dtype).kind == "f")
    # scalar on either side (the plan compiler reverses the operands)
    for k in (3, -4):
        e = Expr(dtype)
        e.materialize(e.floordiv(e.array(), e.scalar(np.dtype(dtype).type(k))))
        e2 = Expr(dtype)
        e2.materialize(e2.floordiv(e2.scalar(np.dtype(dtype).type(k)), e2.array()))
        with np.errstate(all="ignore"):
            assert np.array_equal(oracle.run(e, [a, None], a.size)[0], np.floor_divide(a, np.dtype(dtype).type(k)), equal_nan=True)
            assert np.array_equal(oracle.run(e2, [None, b], b.size)[0], np.floor_divide(np.dtype(dtype).type(k), b), equal_nan=True)


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=lambda d: np.dtype(d).name)
def test_true_division_is_ieee(dtype):
    a, b = _div_columns(dtype)
    e = Expr(dtype)
    e.materialize(e.div(e.array(), e.array()))
    (got,) = oracle.run(e, [a, b], a.size)
    with np.errstate(all="ignore"):
        want = a / b
    assert got.tobytes() == want.tobytes() or np.array_equal(got, want, equal_nan=True)
    with pytest.raises(Exception):
        ei = Expr(np.int64)
        ei.materialize(ei.div(ei.array(), ei.array()))
        abi.compile_plan(ei)                               # int / int would be a float: rejected


# ---- the narrow and the unsigned integer types (PARITY UNPINNED in the reference: pinned against NumPy) ----
@pytest.mark.parametrize("dtype", [np.int8, np.int16, np.uint8, np.uint16, np.uint32, np.uint64], ids=lambda d: np.dtype(d).name)
def test_oracle_new_integer_types_match_numpy(dtype):
    import oracle
    from numpyllvm_b200.abi import Expr
    n = 5000
    a, b = oracle.fill(dtype, 0, n, 1, 1), oracle.fill(dtype, 0, n, 2, 1)      # full range: every op wraps
    e = Expr(dtype); x, y = e.array(), e.array(); e.materialize(e.sub(e.add(e.mul(x, y), e.scalar(3)), e.ref(y)))
    with np.errstate(over="ignore"):
        want = (a * b + dtype(3) - b).astype(dtype)
    assert np.array_equal(oracle.run(e, [a, b, None], n)[0], want)
    e = Expr(dtype); x = e.array(); e.materialize(e.filter(x, e.gt(e.ref(x), e.scalar(50))))
    assert np.array_equal(oracle.run(e, [a, None], n)[0], a[a > dtype(50)])       # unsigned compares for the unsigned types
    e = Expr(dtype); x, y = e.array(), e.array(); e.materialize(e.floordiv(x, y))
    with np.errstate(divide="ignore", over="ignore"):
        want = np.where(b == 0, 0, a // np.where(b == 0, 1, b)).astype(dtype)
    assert np.array_equal(oracle.run(e, [a, b], n)[0], want)
    for name in ("max", "min", "sum"):
        e = Expr(dtype); x = e.array(); e.materialize(getattr(e, name)(x))
        want = getattr(a, name)() if name != "sum" else a.sum(dtype=dtype)        # reductions keep the column type (and wrap)
        assert oracle.run(e, [a], n)[0][0] == want
