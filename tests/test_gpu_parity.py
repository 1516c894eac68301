"""GPU parity: the CUDA pipeline path (through the C-ABI) against the CPU oracle.

Bar: bit-exact for integer ops, masks, filters (values, order, count), max/min and
float elementwise (the kernels are built with -fmad=false); float sums within
rtol 1e-12 (f64) / 1e-6 (f32) of the oracle.
"""
import ctypes as C

import numpy as np
import pytest

import oracle
from numpyllvm_b200 import abi
from numpyllvm_b200.abi import Expr
from numpyllvm_b200.device import Runtime, run_pipeline

pytestmark = pytest.mark.gpu

I64 = np.int64
SIZES = [0, 1, 3, 31, 1023, 1024, 1025, 4096, 4097, 100003, (1 << 20) + 5]
DTYPES = [np.int32, np.int64, np.float32, np.float64]
SUM_RTOL = {np.dtype(np.float32): 1e-6, np.dtype(np.float64): 1e-12}


@pytest.fixture(scope="module")
def rt(nl):
    return Runtime.get()


@pytest.fixture(autouse=True, params=["static", "family", "dynamic"])
def driver(request, nl):
    """Every case runs three times: with the build-time instantiated shapes (nl_static.cuh) where the
    plan matches one, with the shape families (nl_family.cuh) where it matches one of those, and with
    everything forced through the dynamic interpreter."""
    nl.nl_set_static_kernels({"static": 1, "family": 2, "dynamic": 0}[request.param])
    yield request.param
    nl.nl_set_static_kernels(1)


def column(dtype, n, seed, kind=None):
    if kind is None:
        kind = 2 if np.issubdtype(dtype, np.floating) else 0
    return oracle.fill(dtype, 0, n, seed, kind)


def both(rt, expr, inputs, n, **kw):
    ref = oracle.run(expr, inputs, n)
    got = run_pipeline(rt, expr, inputs, n, **kw)
    return ref, got


def assert_same(ref, got):
    assert len(ref) == len(got)
    for r, g in zip(ref, got):
        assert r.dtype == g.dtype and r.shape == g.shape, (r.dtype, g.dtype, r.shape, g.shape)
        assert np.array_equal(r.view(np.uint8), g.view(np.uint8)), "outputs differ bitwise"


# ---- the reference's own tests, on the GPU --------------------------------------
def test_ref_multiplication(rt, golden):
    g = golden["multiplication"]
    ins = [np.array(g["inputs"][k], dtype=I64) for k in "defabc"]
    e = Expr(I64)
    d_, e_, f_, a_, b_, c_ = (e.array() for _ in range(6))
    left = e.mul(e.mul(d_, e_), f_)
    root = e.mul(left, e.mul(e.mul(a_, b_), c_))
    e.materialize(left)
    e.materialize(root)
    left_v, root_v = run_pipeline(rt, e, ins, 100)
    assert left_v.tolist() == g["res"] and root_v.tolist() == g["unsorted_product"]
    e2 = Expr(I64)
    e2.materialize(e2.mul(e2.array(), e2.array()))
    abc = ins[3] * ins[4] * ins[5]
    (res2,) = run_pipeline(rt, e2, [np.sort(left_v), abc], 100)
    assert res2.tolist() == g["res2"]


def test_ref_assignment(rt, golden):
    g = golden["assignment"]
    a, b = np.array(g["inputs"]["a"], dtype=I64), np.array(g["inputs"]["b"], dtype=I64)
    da, db = rt.upload(a), rt.upload(b)
    dc, sizes = rt.empty(I64, 100), rt.empty(I64, 1)
    e = Expr(I64)
    e.materialize(e.mul(e.array(), e.array()))
    plan = abi.compile_plan(e)
    rt.launch(plan, [dc.ptr], [da.ptr, db.ptr], 0, 100, sizes.ptr)          # forced first (thunk.cpp:81-88)
    import ctypes as C
    abi.check(rt.lib.nl_assign_fill(0, 0, C.c_void_p(da.ptr), abi.NL_I64, 0, 1, 1, abi.scalar_bits(1000, I64)))
    assert rt.download(dc).tolist() == g["c"]
    assert rt.download(da)[0] == 1000 and rt.download(da)[1:].tolist() == a[1:].tolist()


def test_driver_selection(rt, driver):
    e = Expr(np.float32)
    e.materialize(e.mul(e.array(), e.array()))
    x = column(np.float32, 5000, 1)
    run_pipeline(rt, e, [x, x], 5000)
    name = rt.last_launch()[0]
    assert ("static:mul2" in name) == (driver == "static"), name


def test_ref_filter(rt, golden):
    g = golden["filter"]
    a = np.array(g["inputs"]["a"], dtype=I64)
    e = Expr(I64)
    a_ = e.array()
    e.materialize(e.filter(a_, e.ge(e.mul(e.ref(a_), e.scalar(2)), e.scalar(50))))
    (res,) = run_pipeline(rt, e, [a, None, None], 100)
    assert res.tolist() == g["res"] and len(res) == 73


def test_ref_max_py(rt, golden):
    g = golden["max"]
    a = np.array(g["inputs"]["a"], dtype=I64)
    e = Expr(I64)
    a_ = e.array()
    f = e.filter(a_, e.ge(e.ref(a_), e.scalar(50)))
    e.materialize(f)
    e.materialize(e.sum(f))
    filt, s = run_pipeline(rt, e, [a, None], 100)
    assert filt.tolist() == g["filtered"] and s.tolist() == [g["res"]]
    # res2 = res * a with res left on the device (NL_IN_SCALAR_DEV)
    e2 = Expr(I64)
    e2.materialize(e2.mul(e2.scalar_dev(), e2.array()))
    (res2,) = run_pipeline(rt, e2, [s, a], 100)
    assert res2.tolist() == g["res2"]


# ---- elementwise ------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", SIZES)
def test_mul_chain(rt, dtype, n):
    ins = [column(dtype, n, s) for s in range(6)]
    e = Expr(dtype)
    d_, e_, f_, a_, b_, c_ = (e.array() for _ in range(6))
    e.materialize(e.mul(e.mul(e.mul(d_, e_), f_), e.mul(e.mul(a_, b_), c_)))
    assert_same(*both(rt, e, ins, n))


@pytest.mark.parametrize("dtype", DTYPES)
def test_int_wrap_and_mixed_ops(rt, dtype):
    n = 70001
    kind = 1 if np.issubdtype(dtype, np.integer) else 2
    x, y = column(dtype, n, 11, kind), column(dtype, n, 12, kind)
    e = Expr(dtype)
    xa, ya = e.array(), e.array()
    t = e.sub(e.mul(xa, ya), e.add(e.ref(xa), e.scalar(3)))          # x*y - (x+3)
    u = e.sub(e.scalar(7), e.mul(e.ref(ya), e.ref(ya)))              # 7 - y*y
    held = e.mul(t, u)
    e.materialize(t)          # a held intermediate: two outputs from one pass
    e.materialize(held)
    assert_same(*both(rt, e, [x, y, None, None], n))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mis", [1, 3])
def test_misaligned_falls_back_to_scalar_rows(rt, dtype, mis):
    n = 50007
    x, y = column(dtype, n, 1), column(dtype, n, 2)
    e = Expr(dtype)
    e.materialize(e.mul(e.array(), e.array()))
    ref = oracle.run(e, [x, y], n)
    got = run_pipeline(rt, e, [x, y], n, misalign_elems=mis)
    assert_same(ref, got)
    name, grid, block, vec = rt.last_launch()
    if (mis * np.dtype(dtype).itemsize) % 16:
        assert vec == 1 and "dynamic" in name, name


@pytest.mark.parametrize("dtype", DTYPES)
def test_subrange_start_end(rt, dtype):
    # the [start,end) contract of the compiled loop (compiler.cpp:340-347)
    n, start = 40000, 4096
    x, y = column(dtype, n, 5), column(dtype, n, 6)
    e = Expr(dtype)
    e.materialize(e.add(e.array(), e.array()))
    (got,) = run_pipeline(rt, e, [x, y], n, start=start)
    assert np.array_equal(got, (x + y)[start:])


# ---- compares / masks ----------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("cmp", ["ge", "gt", "le", "lt", "eq", "ne"])
def test_materialised_masks(rt, dtype, cmp):
    n = 33333
    rng = np.random.default_rng(4)
    x, y = rng.integers(0, 6, n).astype(dtype), rng.integers(0, 6, n).astype(dtype)
    e = Expr(dtype)
    e.materialize(getattr(e, cmp)(e.array(), e.array()))
    ref, got = both(rt, e, [x, y], n)
    assert got[0].dtype == np.bool_ and set(np.unique(got[0].view(np.uint8))) <= {0, 1}
    assert_same(ref, got)


def test_float_compare_nan_is_unordered(rt):
    x = np.array([1.0, np.nan, 3.0, np.nan] * 300, dtype=np.float64)
    y = np.array([1.0, 1.0, np.nan, np.nan] * 300, dtype=np.float64)
    for cmp in ("ge", "lt", "eq", "ne"):
        e = Expr(np.float64)
        e.materialize(getattr(e, cmp)(e.array(), e.array()))
        assert_same(*both(rt, e, [x, y], len(x)))


@pytest.mark.parametrize("dtype", [np.int32, np.float64])
def test_bool_mask_input_and_bool_ops(rt, dtype):
    n = 77777
    rng = np.random.default_rng(8)
    x = rng.integers(0, 100, n).astype(dtype)
    m = rng.integers(0, 2, n).astype(np.bool_)
    e = Expr(dtype)
    xa = e.array()
    ma = e.array(np.bool_)
    mask = e.xor(e.or_(e.and_(ma, e.gt(e.ref(xa), e.scalar(10))), e.not_(e.ref(ma))), e.le(e.ref(xa), e.scalar(3)))
    e.materialize(e.filter(xa, mask))
    assert_same(*both(rt, e, [x, m, None, None], n))


# ---- filters ------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("sel", [0.0, 0.01, 0.5, 0.9, 1.0])
def test_filter_compaction(rt, dtype, n, sel):
    if np.issubdtype(dtype, np.floating):
        x = oracle.fill(dtype, 0, n, 21, 3)
        thr = 1.0 - sel if sel > 0 else 2.0
    else:
        x = oracle.fill(dtype, 0, n, 21, 0)
        thr = 100 - int(99 * sel) if sel > 0 else 1000
    e = Expr(dtype)
    xa = e.array()
    e.materialize(e.filter(xa, e.ge(e.ref(xa), e.scalar(thr))))
    ref, got = both(rt, e, [x, None], n)
    assert_same(ref, got)
    assert np.array_equal(got[0], x[x >= np.dtype(dtype).type(thr)])


def test_filter_with_two_compacted_outputs(rt):
    n = 250001
    x = oracle.fill(np.float32, 0, n, 3, 3)
    e = Expr(np.float32)
    xa = e.array()
    f = e.filter(xa, e.gt(e.ref(xa), e.scalar(0.3)))
    g2 = e.mul(f, e.scalar(2.5))
    e.materialize(f)
    e.materialize(g2)
    assert_same(*both(rt, e, [x, None, None], n))


def test_filter_keeps_unfiltered_intermediate_at_loop_index(rt):
    # a held intermediate evaluated before the branch is stored for every element
    n = 99999
    x = oracle.fill(np.int64, 0, n, 2, 0)
    e = Expr(np.int64)
    xa = e.array()
    dbl = e.mul(xa, e.scalar(2))
    e.materialize(dbl)
    e.materialize(e.filter(e.ref(xa), e.ge(dbl, e.scalar(50))))
    ref, got = both(rt, e, [x, None, None], n)
    assert_same(ref, got)
    assert len(got[0]) == n and len(got[1]) == int((x * 2 >= 50).sum())


# ---- reductions ----------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", SIZES)
@pytest.mark.parametrize("op", ["max", "min", "sum"])
def test_reductions(rt, dtype, n, op):
    kind = 1 if np.issubdtype(dtype, np.integer) else 3
    x = oracle.fill(dtype, 0, n, 31, kind)
    e = Expr(dtype)
    e.materialize(getattr(e, op)(e.array()))
    ref, got = both(rt, e, [x], n)
    if op == "sum" and np.issubdtype(dtype, np.floating):
        tol = SUM_RTOL[np.dtype(dtype)]          # stated tolerance for float sums
        assert abs(float(got[0][0]) - float(ref[0][0])) <= tol * max(1.0, abs(float(ref[0][0])))
    else:
        assert_same(ref, got)                    # ints wrap identically; max/min are exact


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_reduction_nan_propagates(rt, dtype):
    x = oracle.fill(dtype, 0, 300000, 1, 3)
    x[123457] = np.nan
    for op in ("max", "min"):
        e = Expr(dtype)
        e.materialize(getattr(e, op)(e.array()))
        (got,) = run_pipeline(rt, e, [x], len(x))
        assert np.isnan(got[0])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("op", ["max", "sum"])
def test_filter_reduce_fused(rt, dtype, op):
    # C5's shape: (a[a >= t] * k).max() — no compaction, the filter predicates the reduction
    n = 500009
    if np.issubdtype(dtype, np.floating):
        x, thr, k = oracle.fill(dtype, 0, n, 41, 3), 0.75, 3.0
    else:
        x, thr, k = oracle.fill(dtype, 0, n, 41, 0), 60, 3
    e = Expr(dtype)
    xa = e.array()
    e.materialize(getattr(e, op)(e.mul(e.filter(xa, e.ge(e.ref(xa), e.scalar(thr))), e.scalar(k))))
    ref, got = both(rt, e, [x, None, None], n)
    if op == "sum" and np.issubdtype(dtype, np.floating):
        tol = SUM_RTOL[np.dtype(dtype)]
        assert abs(float(got[0][0]) - float(ref[0][0])) <= tol * abs(float(ref[0][0]))
    else:
        assert_same(ref, got)


def test_reduction_is_deterministic(rt):
    x = oracle.fill(np.float64, 0, 3000001, 9, 3)
    e = Expr(np.float64)
    e.materialize(e.sum(e.array()))
    runs = [run_pipeline(rt, e, [x], len(x))[0][0] for _ in range(4)]
    assert len({float(r).hex() for r in runs}) == 1


def test_partials_and_finalize_match_the_eight_chunk_reference(rt):
    # the reference's structure: 8 chunk partials at result[thread_nr], folded in order
    import ctypes as C
    n = 100000
    x = oracle.fill(np.int64, 0, n, 17, 1)
    e = Expr(np.int64)
    e.materialize(e.sum(e.array()))
    plan = abi.compile_plan(e)
    dx, parts, out, sizes = rt.upload(x), rt.empty(I64, 8), rt.empty(I64, 1), rt.empty(I64, 1)
    for j in range(8):
        s, en = oracle.partition(n, 8, j)
        rt.launch(plan, [parts.ptr], [dx.ptr], s, en, sizes.ptr, thread_nr=j)
    abi.check(rt.lib.nl_finalize_partials(0, 0, abi.OP_SUM, abi.NL_I64, C.c_void_p(parts.ptr), 8, C.c_void_p(out.ptr)))
    got_parts = rt.download(parts)
    with np.errstate(over="ignore"):
        expect = [x[slice(*oracle.partition(n, 8, j))].sum() for j in range(8)]
    assert got_parts.tolist() == [int(v) for v in expect]
    assert rt.download(out)[0] == oracle.run(e, [x], n)[0][0]


# ---- in-place assignment -----------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
def test_where_store_in_place(rt, dtype):
    n = 123457
    x = column(dtype, n, 77, 3 if np.issubdtype(dtype, np.floating) else 0)
    thr = 0.5 if np.issubdtype(dtype, np.floating) else 50
    e = Expr(dtype)
    xa = e.array()
    e.materialize(e.where_store(e.scalar(-1), e.gt(xa, e.scalar(thr))))
    ref_target = x.copy()
    oracle.run(e, [ref_target, None, None], n, inplace={0: ref_target})
    (got,) = run_pipeline(rt, e, [x, None, None], n, inplace={0: 0})
    assert np.array_equal(got, ref_target)


@pytest.mark.parametrize("dtype", DTYPES + [np.bool_])
@pytest.mark.parametrize("lo,step,count", [(0, 1, 1), (5, 1, 100000), (3, 7, 1000), (99999, -3, 20000), (1, 1, 1023)])
def test_assign_fill_and_copy(rt, dtype, lo, step, count):
    import ctypes as C
    n = 200000
    base = (np.arange(n) % 2).astype(dtype) if dtype == np.bool_ else column(dtype, n, 3)
    src = (np.arange(count) % 2 == 0).astype(dtype) if dtype == np.bool_ else column(dtype, count, 4)
    value = True if dtype == np.bool_ else 42
    d = rt.upload(base)
    abi.check(rt.lib.nl_assign_fill(0, 0, C.c_void_p(d.ptr), abi.nl_dtype_of(dtype), lo, step, count, abi.scalar_bits(value, dtype)))
    ref = base.copy()
    oracle.assign_fill(ref, lo, step, count, value)
    assert np.array_equal(rt.download(d), ref)
    ds = rt.upload(src)
    abi.check(rt.lib.nl_assign_copy(0, 0, C.c_void_p(d.ptr), abi.nl_dtype_of(dtype), lo, step, count, C.c_void_p(ds.ptr)))
    oracle.assign_copy(ref, lo, step, count, src)
    assert np.array_equal(rt.download(d), ref)
    # NumPy setitem is what the reference delegated to (thunk_as_mapping.cpp:84)
    chk = base.copy()
    chk[lo:lo + step * count:step] = src if count > 1 else src[0]
    assert np.array_equal(ref, chk)


# ---- the synthetic generator is the oracle's, bit for bit -----------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_device_generator_matches_oracle(rt, dtype, kind):
    if np.issubdtype(dtype, np.floating) != (kind >= 2):
        pytest.skip("kind does not apply to this dtype")
    n, first = 100003, (1 << 33) + 17
    d = rt.empty(dtype, n)
    rt.fill(d, first, 42, kind)
    assert np.array_equal(rt.download(d), oracle.fill(dtype, first, n, 42, kind))


# ---- larger sizes ----------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.int64])
def test_large_filter_and_reduce(rt, dtype):
    n = (1 << 24) + 777
    kind = 3 if dtype == np.float64 else 0
    x = oracle.fill(dtype, 0, n, 5, kind)
    thr = 0.5 if dtype == np.float64 else 50
    e = Expr(dtype)
    xa = e.array()
    e.materialize(e.filter(xa, e.gt(e.ref(xa), e.scalar(thr))))
    ref = oracle.run(e, [x, None], n, n_workers=8)
    got = run_pipeline(rt, e, [x, None], n)
    assert_same(ref, got)
    e = Expr(dtype)
    e.materialize(e.max(e.array()))
    assert_same(oracle.run(e, [x], n, n_workers=8), run_pipeline(rt, e, [x], n))


# ---- division (extension on the operator slots) -----------------------------
def assert_same_up_to_nan_payload(ref, got):
    """Bit-exact, except that a NaN CREATED by an operation (0/0, inf - inf) carries the host's
    default payload on x86 (0xFFC00000) and the GPU's canonical one (0x7FFFFFFF): where the oracle has
    a NaN the kernel must have a NaN; every other element must match bitwise."""
    assert len(ref) == len(got)
    for r, g in zip(ref, got):
        assert r.dtype == g.dtype and r.shape == g.shape
        if r.dtype.kind != "f":
            assert np.array_equal(r, g)
            continue
        nan = np.isnan(r)
        assert np.array_equal(nan, np.isnan(g)), "NaN positions differ"
        u = np.uint32 if r.dtype.itemsize == 4 else np.uint64
        assert np.array_equal(r.view(u)[~nan], g.view(u)[~nan]), "outputs differ bitwise"


def _div_columns(dtype, n):
    rng = np.random.default_rng(12)
    if np.dtype(dtype).kind == "i":
        info = np.iinfo(dtype)
        a = rng.integers(-1000, 1000, n).astype(dtype)
        b = rng.integers(-9, 10, n).astype(dtype)
        a[:4] = [info.min, info.min, info.max, 0]
        b[:4] = [-1, 1, -1, 0]
    else:
        a = ((rng.random(n) - 0.5) * 100).astype(dtype)
        b = ((rng.random(n) - 0.5) * 7).astype(dtype)
        a[:12] = [1.0, -1.0, 0.0, -0.0, 7.5, np.inf, np.nan, 1e30, 0.3, 0.9, 6.0, -6.0]
        b[:12] = [0.0, 0.0, 0.0, 3.0, -2.5, 2.0, 1.0, 1e-30, 0.1, 0.3, 0.1, 0.1]
    return a, b


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [1000, 300_001])
def test_floor_division(rt, dtype, n):
    a, b = _div_columns(dtype, n)
    e = Expr(dtype)
    e.materialize(e.floordiv(e.array(), e.array()))
    assert_same_up_to_nan_payload(*both(rt, e, [a, b], n))
    k = np.dtype(dtype).type(-3)
    e = Expr(dtype)
    e.materialize(e.floordiv(e.array(), e.scalar(k)))
    assert_same_up_to_nan_payload(*both(rt, e, [a, None], n))
    e = Expr(dtype)
    e.materialize(e.add(e.floordiv(e.scalar(k), e.array()), e.scalar(np.dtype(dtype).type(1))))
    assert_same_up_to_nan_payload(*both(rt, e, [None, b, None], n))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_true_division(rt, dtype):
    n = 300_001
    a, b = _div_columns(dtype, n)
    e = Expr(dtype)
    e.materialize(e.div(e.array(), e.array()))
    assert_same_up_to_nan_payload(*both(rt, e, [a, b], n))
    e = Expr(dtype)
    xa = e.array()
    e.materialize(e.filter(e.div(e.scalar(np.dtype(dtype).type(1)), xa), e.gt(e.ref(xa), e.scalar(np.dtype(dtype).type(0)))))
    assert_same_up_to_nan_payload(*both(rt, e, [a, None, None], n))          # 1/a where a > 0: division before a compaction


@pytest.mark.parametrize("nbytes", [0, 8, 24, 4096 + 8, (1 << 22) + 16, (1 << 22) + 12, 1001])
@pytest.mark.parametrize("misalign", [0, 8])
def test_checksum_kernel_matches_its_host_definition(nl, nbytes, misalign):
    """nl_checksum (the device half of bench.py's full-size parity checks) against the NumPy restatement."""
    import bench
    rt = Runtime.get()
    rng = np.random.default_rng(nbytes + misalign)
    host = rng.integers(0, 256, nbytes + misalign, dtype=np.uint8)
    d = rt.upload(host) if host.size else rt.empty(np.uint8, 16)
    out = rt.empty(np.uint64, 3)
    abi.check(nl.nl_checksum(0, 0, C.c_void_p(d.ptr + misalign), nbytes, C.c_void_p(out.ptr)))
    got = tuple(int(v) for v in rt.download(out))
    assert got == bench.words_checksum(host[misalign:])[1:]


def test_cuda_graph_replays_a_chain_of_small_pipelines(nl, rt):
    """nl_graph_*: three dependent small pipelines captured once and replayed; the replay computes what the
    eager launches compute (launch-bound small pipelines)."""
    n = 50_000
    x = oracle.fill(np.float32, 0, n, 2, 2)
    dx = rt.upload(x)
    t1, t2, res = rt.empty(np.float32, n), rt.empty(np.float32, n), rt.empty(np.float32, 1)
    sizes = rt.empty(np.int64, 4)
    e1 = Expr(np.float32); e1.materialize(e1.add(e1.mul(e1.array(), e1.scalar(2.0)), e1.scalar(1.0)))
    e2 = Expr(np.float32); e2.materialize(e2.mul(e2.array(), e2.array()))
    e3 = Expr(np.float32); a3 = e3.array(); e3.materialize(e3.max(e3.filter(a3, e3.gt(e3.ref(a3), e3.scalar(3.0)))))
    p1, p2, p3 = (abi.compile_plan(e) for e in (e1, e2, e3))

    def enqueue():
        rt.launch(p1, [t1.ptr], [dx.ptr, None, None], 0, n, sizes.ptr)
        rt.launch(p2, [t2.ptr], [t1.ptr, dx.ptr], 0, n, sizes.ptr)
        rt.launch(p3, [res.ptr], [t2.ptr, None], 0, n, sizes.ptr)
    enqueue()
    want = rt.download(res)[0]
    assert want == ((x * 2 + 1) * x).max()
    abi.check(nl.nl_graph_begin(0))
    enqueue()
    g = C.c_void_p()
    abi.check(nl.nl_graph_end(0, C.byref(g)))
    for rep in range(5):
        x2 = oracle.fill(np.float32, 0, n, 3 + rep, 2)
        abi.check(nl.nl_memcpy_h2d(0, 0, C.c_void_p(dx.ptr), C.c_void_p(x2.ctypes.data), x2.nbytes))
        abi.check(nl.nl_graph_launch(g))
        assert rt.download(res)[0] == ((x2 * 2 + 1) * x2).max()
    abi.check(nl.nl_graph_destroy(g))


# ---- the narrow and the unsigned integer types: interpreter kernels vs the oracle ----------------
NEW_DTYPES = [np.int8, np.int16, np.uint8, np.uint16, np.uint32, np.uint64]


@pytest.mark.parametrize("dtype", NEW_DTYPES, ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("n", [0, 1, 1023, 4097, 300_001])
def test_new_integer_types(rt, driver, dtype, n):
    if driver != "dynamic":
        pytest.skip("these types have interpreter kernels only")
    a, b = oracle.fill(dtype, 0, n, 11, 1), oracle.fill(dtype, 0, n, 12, 1)      # full range: everything wraps
    small = oracle.fill(dtype, 0, n, 13, 0)                                      # [1, 100)
    cases = []
    e = Expr(dtype); x, y = e.array(), e.array(); e.materialize(e.sub(e.add(e.mul(x, y), e.scalar(3)), e.ref(y)))
    cases.append((e, [a, b, None]))
    e = Expr(dtype); x = e.array(); e.materialize(e.filter(x, e.gt(e.ref(x), e.scalar(50))))
    cases.append((e, [small, None]))
    e = Expr(dtype); x, y = e.array(), e.array(); e.materialize(e.filter(y, e.le(x, e.ref(y))))
    cases.append((e, [a, b]))
    e = Expr(dtype); x, y = e.array(), e.array(); e.materialize(e.floordiv(x, y))
    cases.append((e, [a, b]))
    e = Expr(dtype); x = e.array(); e.materialize(e.ge(e.mul(x, e.scalar(2)), e.scalar(77)))
    cases.append((e, [a, None, None]))
    for name in ("max", "min", "sum"):
        e = Expr(dtype); x = e.array(); e.materialize(getattr(e, name)(x))
        cases.append((e, [a]))
        e = Expr(dtype); x = e.array(); e.materialize(getattr(e, name)(e.filter(x, e.lt(e.ref(x), e.scalar(40)))))
        cases.append((e, [small, None]))
    for e, ins in cases:
        want = oracle.run(e, ins, n)
        got = run_pipeline(rt, e, ins, n)
        for w, g in zip(want, got):
            assert w.dtype == g.dtype and w.shape == g.shape and w.tobytes() == g.tobytes()


@pytest.mark.parametrize("dtype", NEW_DTYPES, ids=lambda d: np.dtype(d).name)
def test_new_integer_types_large_filter_runs_the_super_tile_kernel(rt, nl, driver, dtype):
    if driver != "dynamic":
        pytest.skip("these types have interpreter kernels only")
    n = 256 * 4 * 4 * 4 * 148 + 12345           # above one super-tile (4 tiles of 4 rows x 4 elements) per SM
    a = oracle.fill(dtype, 0, n, 21, 0)
    e = Expr(dtype); x = e.array(); e.materialize(e.filter(x, e.ge(e.ref(x), e.scalar(30))))
    (got,) = run_pipeline(rt, e, [a, None], n)
    assert "compact_super_kernel" in rt.last_launch()[0]
    assert got.tobytes() == a[a >= dtype(30)].tobytes()


ALL_VALUE_DTYPES = DTYPES + NEW_DTYPES


@pytest.mark.parametrize("src", ALL_VALUE_DTYPES + [np.bool_], ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("dst", ALL_VALUE_DTYPES, ids=lambda d: np.dtype(d).name)
def test_cast_matches_numpy_astype(nl, rt, src, dst):
    n = 100_003
    if src == np.bool_:
        x = (oracle.fill(np.int64, 0, n, 3, 0) > 50)
    elif np.dtype(src).kind == "f":
        x = (oracle.fill(src, 0, n, 3, 3) * 200 - (0 if np.dtype(dst).kind == "u" else 100)).astype(src)   # in range of every target
    else:
        x = oracle.fill(src, 0, n, 3, 0 if np.dtype(dst).kind == "f" else 1)
        if np.dtype(dst).kind == "f" and np.dtype(src).itemsize == 8:
            x = oracle.fill(src, 0, n, 3, 1)                                      # full range: rounding of 64-bit ints to float
    d_in = rt.upload(x)
    d_out = rt.empty(dst, n)
    abi.check(nl.nl_cast(0, 0, C.c_void_p(d_out.ptr), abi.nl_dtype_of(dst), C.c_void_p(d_in.ptr), abi.nl_dtype_of(x.dtype), n))
    with np.errstate(invalid="ignore", over="ignore"):
        want = x.astype(dst)
    assert rt.download(d_out).tobytes() == want.tobytes()


@pytest.mark.parametrize("dtype", NEW_DTYPES, ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("n", [1000, 300_001])
def test_sort_new_integer_types(rt, dtype, n):
    x = oracle.fill(dtype, 0, n, 31, 1)
    d = rt.upload(x)
    rt.sort(d)
    assert rt.download(d).tobytes() == np.sort(x).tobytes()


def test_large_blocks_are_recycled_by_size_class(rt, nl):
    """nl_malloc / nl_free: a block above 1 MiB comes back for the next request of its size class (sizes that differ by
    a few elements — filter outputs, sort buckets — must not make the pool map fresh memory at every evaluate), a
    live block is never handed out twice, small blocks and nl_pool_trim() bypass / empty the cache."""
    def alloc(nbytes):
        p = C.c_void_p()
        abi.check(nl.nl_malloc(0, nbytes, C.byref(p)))
        return p.value
    abi.check(nl.nl_pool_trim(0))
    a = alloc((3 << 20) + 8)
    abi.check(nl.nl_free(0, C.c_void_p(a)))
    b = alloc((3 << 20) + 4096)                 # same class (granule 2 MiB at this size): the parked block
    assert b == a
    c = alloc((3 << 20) + 8)                    # the class's block is live: a different one
    assert c != b
    abi.check(nl.nl_free(0, C.c_void_p(b)))
    abi.check(nl.nl_free(0, C.c_void_p(c)))
    assert {alloc(3 << 20), alloc(3 << 20)} == {b, c}
    abi.check(nl.nl_free(0, C.c_void_p(b)))
    abi.check(nl.nl_free(0, C.c_void_p(c)))
    x = rt.upload(np.arange(1 << 19, dtype=np.int64))          # 4 MiB: takes a parked block; its content must be its own
    assert np.array_equal(rt.download(x), np.arange(1 << 19, dtype=np.int64))
    x.free()
    abi.check(nl.nl_pool_trim(0))               # empties the cache (and the pool): everything still works afterwards
    d = alloc((3 << 20) + 8)
    abi.check(nl.nl_free(0, C.c_void_p(d)))
    s = alloc(1000)                             # small blocks go straight to the stream-ordered pool
    abi.check(nl.nl_free(0, C.c_void_p(s)))


def test_compaction_arena_stays_consistent_across_kernels_and_graph_capture(nl, rt):
    """The super-tile compaction kernel leaves ticket and descriptors clean for the next launch; the other compaction
    kernels do not, and a launch that is only CAPTURED into a graph cleans nothing.  Interleave all three and replay."""
    n_big, n_small = 3_000_001, 5_000
    x = oracle.fill(np.float64, 0, n_big, 9, 3)
    dx, out, sizes = rt.upload(x), rt.empty(np.float64, n_big), rt.empty(np.int64, 4)

    def plan_for(t):
        e = Expr(np.float64)
        a = e.array()
        e.materialize(e.filter(a, e.gt(e.ref(a), e.scalar(t))))
        return abi.compile_plan(e)

    def run(t, n):
        rt.launch(plan_for(t), [out.ptr], [dx.ptr, None], 0, n, sizes.ptr)
        k = int(rt.download(sizes, 1)[0])
        assert np.array_equal(rt.download(out, k), x[:n][x[:n] > t])

    run(0.5, n_big)                      # super-tile kernel
    assert "compact_super" in rt.last_launch()[0]
    run(0.5, n_small)                    # plain kernel: leaves its descriptors behind
    assert "compact_super" not in rt.last_launch()[0]
    abi.check(nl.nl_graph_begin(0))      # captured, not run: must not count as a clean-up
    rt.launch(plan_for(0.25), [out.ptr], [dx.ptr, None], 0, n_big, sizes.ptr)
    g = C.c_void_p()
    abi.check(nl.nl_graph_end(0, C.byref(g)))
    run(0.9, n_big)                      # super-tile kernel right after: the arena was still dirty
    run(0.1, n_big)
    for _ in range(3):
        abi.check(nl.nl_graph_launch(g))
        k = int(rt.download(sizes, 1)[0])
        assert np.array_equal(rt.download(out, k), x[x > 0.25])
        run(0.7, n_small)                # dirty again between the replays
    abi.check(nl.nl_graph_destroy(g))
    for a in (dx, out, sizes):
        a.free()
