"""Device radix sort (nl_sort) — the sort() full breaker — against NumPy's sort, which is what the
reference runs (PyArray_Sort on a copy, thunk_methods.cpp:37-62).  Bit-exact for ints; floats are
compared bit-for-bit too on data without mixed-sign zeros (NumPy leaves the relative order of
-0.0 / +0.0 unspecified; the radix order puts -0.0 first)."""
import ctypes as C

import numpy as np
import pytest

import oracle
from numpyllvm_b200 import abi
from numpyllvm_b200.device import Runtime

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rt():
    return Runtime.get()


def dev_sort(rt, x):
    d = rt.upload(x)
    rt.sort(d)
    out = rt.download(d)
    d.free()
    return out


@pytest.mark.parametrize("dtype", [np.int32, np.int64, np.float32, np.float64], ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("n", [0, 1, 2, 100, 2047, 2048, 2049, 100_003, 3_000_001])
@pytest.mark.parametrize("kind", ["small", "full"])
def test_sort_matches_numpy(rt, dtype, n, kind):
    if np.dtype(dtype).kind == "i":
        x = oracle.fill(dtype, 0, n, 0 if kind == "small" else 1, 77)     # [1,100) like the reference's tests / full range
    else:
        x = oracle.fill(dtype, 0, n, 2 if kind == "small" else 3, 77)
        if kind == "full" and n:
            x = ((x - 0.5) * np.dtype(dtype).type(1e6)).astype(dtype)     # both signs, wide exponent range
            x[x == 0] = 1                                                 # no mixed-sign zeros
    got = dev_sort(rt, x)
    assert got.tobytes() == np.sort(x).tobytes()


def test_sort_is_the_reference_test_vector(rt, golden):
    """Tests/multiplication.py: res2 = sort(d*e*f) * (a*b*c) on the seed-42 vectors (committed golden result)."""
    m = golden["multiplication"]
    a, b, c, d, e, f = (np.array(m["inputs"][k], dtype=np.int64) for k in "abcdef")
    assert np.array_equal(dev_sort(rt, d * e * f) * (a * b * c), np.array(m["res2"], dtype=np.int64))


def test_special_floats(rt):
    x = np.array([3.0, -np.inf, np.nan, 1.5, np.inf, -2.0, np.nan, 0.0, -1e-310, 1e-310, 7.0] * 300, dtype=np.float64)
    got = dev_sort(rt, x)
    want = np.sort(x)
    assert np.array_equal(got, want, equal_nan=True)
    assert np.isnan(got[-600:]).all() and not np.isnan(got[:-600]).any()
    neg_nan = np.array([np.nan, 1.0, -np.nan, -1.0, np.inf], dtype=np.float32)
    neg_nan[2] = np.frombuffer(np.uint32(0xFFC00000).tobytes(), dtype=np.float32)[0]   # NaN with the sign bit set
    got = dev_sort(rt, np.tile(neg_nan, 1000))
    assert np.isnan(got[-2000:]).all() and got[0] == -1.0 and got[2999] == np.inf


def test_sort_is_stable_for_equal_keys_and_handles_duplicates(rt):
    x = np.repeat(np.arange(50, dtype=np.int64)[::-1], 4001)
    assert np.array_equal(dev_sort(rt, x), np.sort(x))
    x = np.full(10_000, 7, dtype=np.int32)          # every digit position is trivial: zero passes
    assert np.array_equal(dev_sort(rt, x), x)


@pytest.mark.parametrize("dtype", [np.int32, np.int64, np.float32, np.float64, np.uint32, np.uint64], ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("algo", [0, 1])
def test_both_radix_forms(rt, dtype, algo):
    """nl_tune("sort_algo"): one sweep per digit position (default) and count + scatter give NumPy's order."""
    rng = np.random.default_rng(5)
    n = 1_234_567
    if np.dtype(dtype).kind == "f":
        x = ((rng.random(n) - 0.5) * 1e9).astype(dtype)
        x[x == 0] = 1
    else:
        info = np.iinfo(dtype)
        x = rng.integers(info.min, info.max, size=n, dtype=dtype, endpoint=True)
    abi.check(rt.lib.nl_tune(b"sort_algo", algo))
    try:
        got = dev_sort(rt, x)
    finally:
        abi.check(rt.lib.nl_tune(b"sort_algo", 0))
    assert got.tobytes() == np.sort(x).tobytes()


@pytest.mark.parametrize("dtype", [np.int64, np.uint32], ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("n", [8191, 8192, 8193, 100_003, 1_000_000])
def test_one_sweep_chains_its_launches(rt, dtype, n):
    """Above 2^28 keys the one-sweep form runs in several launches per position, each continuing where the previous
    one stopped (the last tile leaves the running digit offsets behind); a 2^13-key portion exercises that chain —
    portions smaller than a tile included."""
    rng = np.random.default_rng(n)
    info = np.iinfo(dtype)
    x = rng.integers(info.min, info.max, size=n, dtype=dtype, endpoint=True)
    x[: n // 3] &= 0xFF                      # skewed digits: long runs of one digit in the upper positions
    for log2 in (13, 16):
        abi.check(rt.lib.nl_tune(b"sort_portion_log2", log2))
        try:
            got = dev_sort(rt, x)
        finally:
            abi.check(rt.lib.nl_tune(b"sort_portion_log2", 0))
        assert got.tobytes() == np.sort(x).tobytes()


def order_keys(x):
    """nl_sort's order-preserving bijection, widened to 64 bits (integers only here)."""
    dt = x.dtype
    u = x.view(np.dtype(f"u{dt.itemsize}")).astype(np.uint64)
    if dt.kind == "i":
        u ^= np.uint64(1 << (8 * dt.itemsize - 1))
    return u


@pytest.mark.parametrize("dtype", [np.int8, np.uint16, np.int32, np.int64, np.uint64], ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("n", [1, 1000, 8192, 300_001])
@pytest.mark.parametrize("m", [1, 7, 15])
def test_split_groups_stably_at_the_splitters(rt, dtype, n, m):
    """nl_sort_sample / nl_sort_split: the device steps of the sort across GPUs."""
    rng = np.random.default_rng(n + m)
    info = np.iinfo(dtype)
    x = rng.integers(info.min, info.max, size=n, dtype=dtype, endpoint=True)
    d, out = rt.upload(x), rt.empty(dtype, n)
    k = 64
    samples = (C.c_uint64 * k)()
    abi.check(rt.lib.nl_sort_sample(d.dev, 0, abi.nl_dtype_of(dtype), C.c_void_p(d.ptr), n, k, samples))
    rt.sync()
    assert np.array_equal(np.array(samples[:], dtype=np.uint64), order_keys(x[(np.arange(k) * n) // k]))
    spl = np.sort(order_keys(rng.integers(info.min, info.max, size=m, dtype=dtype, endpoint=True)))
    if m > 1:
        spl[1] = spl[0]                     # an empty run in the middle
    counts = (C.c_int64 * (m + 1))()
    abi.check(rt.lib.nl_sort_split(d.dev, 0, abi.nl_dtype_of(dtype), C.c_void_p(d.ptr), C.c_void_p(out.ptr), n,
                                   spl.ctypes.data_as(C.POINTER(C.c_uint64)), m, counts))
    rt.sync()
    run = np.searchsorted(spl, order_keys(x), side="left")      # number of splitters below the key
    want = x[np.argsort(run, kind="stable")]
    assert list(counts) == np.bincount(run, minlength=m + 1).tolist()
    assert np.array_equal(rt.download(out), want)
    d.free(); out.free()


@pytest.mark.parametrize("dtype", [np.int32, np.uint64], ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("n", [1000, 300_001])
def test_split_in_two_steps_scatters_the_runs_to_their_destinations(rt, dtype, n):
    """nl_sort_split_count + nl_sort_split_scatter: every run goes to an address of its own (across GPUs these are
    the buckets on the destination devices; here separate buffers, one of them missing because its run is empty)."""
    rng = np.random.default_rng(n)
    info = np.iinfo(dtype)
    x = rng.integers(info.min, info.max, size=n, dtype=dtype, endpoint=True)
    m = 5
    spl = np.sort(order_keys(rng.integers(info.min, info.max, size=m, dtype=dtype, endpoint=True)))
    spl[0] = 0                                   # run 0 is empty for unsigned keys, almost empty for signed ones
    spl[3] = spl[2]                              # an empty run in the middle
    d = rt.upload(x)
    counts = (C.c_int64 * (m + 1))()
    p_spl = spl.ctypes.data_as(C.POINTER(C.c_uint64))
    abi.check(rt.lib.nl_sort_split_count(d.dev, 0, abi.nl_dtype_of(dtype), C.c_void_p(d.ptr), n, p_spl, m, counts))
    rt.sync()
    run = np.searchsorted(spl, order_keys(x), side="left")
    assert list(counts) == np.bincount(run, minlength=m + 1).tolist()
    bufs = [rt.empty(dtype, c + 3) if c else None for c in counts]       # every run 3 elements into its own buffer
    es = np.dtype(dtype).itemsize
    dst = (C.c_void_p * (m + 1))(*[C.c_void_p(b.ptr + 3 * es) if b else C.c_void_p(None) for b in bufs])
    abi.check(rt.lib.nl_sort_split_scatter(d.dev, 0, abi.nl_dtype_of(dtype), C.c_void_p(d.ptr), n, p_spl, m, dst))
    rt.sync()
    for j, b in enumerate(bufs):
        if b is None:
            assert not (run == j).any()
            continue
        assert np.array_equal(rt.download(b)[3:], x[run == j])           # stable within the run
        b.free()
    d.free()
