"""GPU parity of the staged compaction pipeline (nl_compact_pipe.cuh: TMA producer warp, consumer
warps, scan warps) against the oracle, and of the plain compaction kernel it replaces for large
inputs.  Bit-exact: values, order and count."""
import numpy as np
import pytest

import oracle
from numpyllvm_b200 import abi
from numpyllvm_b200.abi import Expr
from numpyllvm_b200.device import Runtime, run_pipeline

pytestmark = pytest.mark.gpu

# (nl_set_static_kernels mode, nl_tune "compact_kernel", nl_tune "pipe_dynamic"): every compaction kernel under
# exact build-time shapes (1), shape families (2) and the step interpreter (0)
MODES = {"pipe-static": (1, 2, 1), "pipe-dynamic": (0, 2, 1), "super-static": (1, 0, 0), "super-family": (2, 0, 0),
         "super-dynamic": (0, 0, 0), "plain-static": (1, 1, 0), "plain-family": (2, 1, 0), "plain-dynamic": (0, 1, 0)}


@pytest.fixture(scope="module")
def rt(nl):
    return Runtime.get()


@pytest.fixture(params=list(MODES))
def mode(request, nl):
    tier, which, dyn_pipe = MODES[request.param]
    nl.nl_set_static_kernels(tier)
    nl.nl_tune(b"compact_kernel", which)
    nl.nl_tune(b"pipe_dynamic", dyn_pipe)
    yield request.param
    nl.nl_set_static_kernels(1)
    nl.nl_tune(b"reset", 0)


def tile_elems(dtype):
    return 512 * 4 * (16 // np.dtype(dtype).itemsize)


def check(rt, mode, e, inputs, n, expect_pipe=True, expect_super=True):
    ref = oracle.run(e, inputs, n, n_workers=8)
    got = run_pipeline(rt, e, inputs, n)
    name = rt.last_launch()[0]
    if mode.startswith("super"):
        # below its own size threshold the super-tile kernel hands over to the staged pipeline / plain kernel
        assert ("compact_super_kernel" in name) == expect_super, name
    else:
        assert ("compact_pipe_kernel" in name) == (expect_pipe and mode.startswith("pipe")), name
    assert len(ref) == len(got)
    for r, g in zip(ref, got):
        assert r.shape == g.shape and np.array_equal(r.view(np.uint8), g.view(np.uint8))
    return got


@pytest.mark.parametrize("dtype", [np.int32, np.int64, np.float32, np.float64])
@pytest.mark.parametrize("sel", [0.0, 0.003, 0.5, 1.0])
def test_filter_all_dtypes(rt, mode, dtype, sel):
    n = tile_elems(dtype) * 1300 + 12345          # above the pipeline threshold, ragged last tile
    if np.issubdtype(dtype, np.floating):
        x, thr = oracle.fill(dtype, 0, n, 5, 3), (1.0 - sel if sel > 0 else 2.0)
    else:
        x, thr = oracle.fill(dtype, 0, n, 5, 0), (100 - int(99 * sel) if sel > 0 else 1000)
    e = Expr(dtype)
    xa = e.array()
    e.materialize(e.filter(xa, e.ge(e.ref(xa), e.scalar(thr))))
    (got,) = check(rt, mode, e, [x, None], n)
    assert np.array_equal(got, x[x >= np.dtype(dtype).type(thr)])


@pytest.mark.parametrize("extra", [0, 1, -1, 2048, 4097])
def test_threshold_and_tile_boundaries(rt, mode, extra):
    n = tile_elems(np.float64) * 296 + extra      # exactly at / just below the switch-over size
    x = oracle.fill(np.float64, 0, n, 9, 3)
    e = Expr(np.float64)
    xa = e.array()
    e.materialize(e.filter(xa, e.gt(e.ref(xa), e.scalar(0.25))))
    # (the interpreter's super-tile kernel has 4-row tiles: the same size is exactly ITS switch-over point)
    check(rt, mode, e, [x, None], n, expect_pipe=(extra >= 0), expect_super=(mode == "super-dynamic" and extra >= 0))


@pytest.mark.parametrize("extra", [0, 1, -1, 4096 * 4 - 1, 4096 * 4 + 1])
def test_super_tile_threshold_and_boundaries(rt, mode, extra):
    n = 256 * 8 * 2 * 4 * 148 + extra             # one super-tile (4 x 4096 f64) per SM: the switch-over size
    x = oracle.fill(np.float64, 0, n, 9, 3)
    e = Expr(np.float64)
    xa = e.array()
    e.materialize(e.filter(xa, e.gt(e.ref(xa), e.scalar(0.25))))
    check(rt, mode, e, [x, None], n, expect_super=(extra >= 0 or mode == "super-dynamic"))


def test_reference_filter_shape_and_two_outputs(rt, mode):
    # Tests/filter.py's a[a*2 >= 50] (re-reads a temporary after the filter) and a pipeline that
    # stores two compacted outputs
    n = tile_elems(np.int64) * 1300 + 7
    x = oracle.fill(np.int64, 0, n, 42, 0)
    e = Expr(np.int64)
    xa = e.array()
    e.materialize(e.filter(xa, e.ge(e.mul(e.ref(xa), e.scalar(2)), e.scalar(50))))
    check(rt, mode, e, [x, None, None], n)
    e = Expr(np.int64)
    xa = e.array()
    f = e.filter(xa, e.gt(e.ref(xa), e.scalar(30)))
    e.materialize(f)
    e.materialize(e.mul(f, e.scalar(3)))
    check(rt, mode, e, [x, None, None], n)


def test_ineligible_pipelines_stay_on_the_plain_kernel(rt, mode):
    n = tile_elems(np.int64) * 1300
    x = oracle.fill(np.int64, 0, n, 1, 0)
    y = oracle.fill(np.int64, 0, n, 2, 0)
    # an intermediate stored BEFORE the filter (loop index, every element)
    e = Expr(np.int64)
    xa = e.array()
    dbl = e.mul(xa, e.scalar(2))
    e.materialize(dbl)
    e.materialize(e.filter(e.ref(xa), e.ge(dbl, e.scalar(50))))
    check(rt, mode, e, [x, None, None], n, expect_pipe=False, expect_super=False)
    # nine columns cannot be staged (the plan compiler caps a pipeline at eight inputs anyway); five
    # staged columns leave fewer than four ring slots even with the narrow tiles
    e = Expr(np.int64)
    cols = [e.array() for _ in range(5)]
    acc = cols[0]
    for c in cols[1:]:
        acc = e.add(acc, c)
    e.materialize(e.filter(acc, e.gt(e.ref(cols[0]), e.scalar(40))))
    check(rt, mode, e, [x, y, x, y, x, None], n, expect_pipe=False)


@pytest.mark.parametrize("dtype", [np.int64, np.float32, np.float64], ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("sel", [0.02, 0.5, 0.97])
def test_two_and_three_staged_columns_use_the_narrow_tiles(rt, mode, dtype, sel):
    """b[a > t] and (a*b + c)[a > t]: several columns share a ring slot (4-row instantiation)."""
    n = tile_elems(dtype) * 1330 + 17
    kind = 0 if np.dtype(dtype).kind == "i" else 3
    a = oracle.fill(dtype, 0, n, kind, 1)
    b = oracle.fill(dtype, 0, n, kind, 2)
    c = oracle.fill(dtype, 0, n, kind, 3)
    thr = np.dtype(dtype).type(np.sort(a)[int((1 - sel) * (n - 1))])
    e = Expr(dtype)
    xa, xb = e.array(), e.array()
    e.materialize(e.filter(xb, e.gt(xa, e.scalar(thr))))
    # b[a > t] has a build-time shape whose plain kernel is the faster one: no pipeline in static mode
    check(rt, mode, e, [a, b, None], n, expect_pipe=(mode == "pipe-dynamic"))
    e = Expr(dtype)
    xa, xb, xc = e.array(), e.array(), e.array()
    e.materialize(e.filter(e.add(e.mul(e.ref(xa), xb), xc), e.gt(xa, e.scalar(thr))))
    check(rt, mode, e, [a, b, c, None], n)


def test_repeated_launches_reuse_the_descriptors(rt, mode):
    # the descriptor array and the ticket are reset per launch; back-to-back launches on one
    # stream with different selectivities must not see each other's state
    n = tile_elems(np.float32) * 1300 + 99
    x = oracle.fill(np.float32, 0, n, 3, 3)
    for thr in (0.9, 0.1, 0.5, 0.9):
        e = Expr(np.float32)
        xa = e.array()
        e.materialize(e.filter(xa, e.gt(e.ref(xa), e.scalar(thr))))
        (got,) = check(rt, mode, e, [x, None], n)
        assert len(got) == int((x > np.float32(thr)).sum())


@pytest.mark.parametrize("dtype", [np.int32, np.int64, np.float32, np.float64], ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("ops", [("gt", "lt"), ("ge", "le"), ("ge", "lt"), ("gt", "le")], ids=lambda o: "_".join(o))
def test_range_filter_shapes(rt, mode, dtype, ops):
    """a[(a lo s1) & (a hi s2)] has build-time shapes (predicate temporaries before the filter)."""
    n = tile_elems(dtype) * 1300 + 5
    kind = 0 if np.dtype(dtype).kind == "i" else 3
    x = oracle.fill(dtype, 0, n, kind, 6)
    lo, hi = (np.dtype(dtype).type(v) for v in ((20, 70) if kind == 0 else (0.2, 0.7)))
    e = Expr(dtype)
    xa = e.array()
    m = e.and_(getattr(e, ops[0])(e.ref(xa), e.scalar(lo)), getattr(e, ops[1])(e.ref(xa), e.scalar(hi)))
    e.materialize(e.filter(xa, m))
    check(rt, mode, e, [x, None, None], n)
    if "static" in mode:
        assert "filter_range_" in rt.last_launch()[0]
