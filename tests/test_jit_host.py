"""CPU-only checks of the `jit` extension's host logic: the Python surface of the reference
(jitpackage.cpp, thunk.cpp, thunk_as_*.cpp, thunk_methods.cpp), the parser's pipeline split
(parser.cpp:81-173) and the plan optimiser.  Nothing here touches a GPU."""
import numpy as np
import pytest

import numpyllvm_b200

jit = numpyllvm_b200.install_jit_alias()


def vec(n=100, seed=42, k=1):
    np.random.seed(seed)
    out = [np.random.randint(1, 100, size=(n,)) for _ in range(k)]
    return out if k > 1 else out[0]


def has_gpu():
    from numpyllvm_b200 import abi
    lib = abi.load_library()
    return lib.nl_device_count() > 0 or lib.nl_init(0, None) == 0


def test_module_surface_matches_the_reference():
    import jit as alias
    assert alias is jit                                   # `import jit` is the drop-in name
    assert jit.__doc__ == "This module provides THUNKS."  # jitpackage.cpp:5
    t = jit.thunk(vec())
    assert type(t).__name__ == "thunk" and type(t) is jit.thunk_type
    for m in ("evaluate", "isevaluated", "asnumpyarray", "sort", "max", "sum"):    # thunk_methods.cpp:186-194
        assert callable(getattr(t, m))
    with pytest.raises(TypeError):
        hash(t)                                           # tp_hash = PyObject_HashNotImplemented (thunk.cpp:156)
    assert t.isevaluated() is True
    assert (t * t).isevaluated() is False


def test_thunk_rejects_non_arrays_like_the_reference():
    with pytest.raises(TypeError):
        jit.thunk(object())
    with pytest.raises(TypeError):
        jit.thunk(np.array(["a", "b"]))
    with pytest.raises(TypeError):
        jit.thunk(np.arange(4, dtype=np.float16))          # f16 is the one getLLVMType entry (compiler.cpp:156-182) not built
    with pytest.raises(TypeError):
        jit.thunk(np.arange(4, dtype=np.complex64))
    for dt in (np.int8, np.int16, np.uint8, np.uint16, np.uint32, np.uint64):
        assert jit.thunk(np.arange(4, dtype=dt)).isevaluated()


def test_lazy_operators_build_thunks_without_computing():
    a, b = (jit.thunk(v) for v in vec(k=2))
    for expr in (a * b, a * 2, 2 * a, a + b, a - 1, a >= 50, a > b, a <= 3, a < b, a == b, a != 7,
                 a[a >= 50], a.max(), a.min(), a.sum(), a.sort(), (a >= 3) & (a < 9), ~(a >= 3)):
        assert type(expr) is jit.thunk_type and not expr.isevaluated()


def test_error_behaviour():
    a = jit.thunk(vec())
    with pytest.raises(IndexError):
        a[3]                                              # "FIXME: other types of array access"
    with pytest.raises(IndexError):
        a[jit.thunk(np.arange(3.0))]                      # indices must be integer or boolean
    with pytest.raises(IndexError):
        a[jit.thunk(np.arange(3, dtype=np.int32))]        # integer gather takes int64 indices
    with pytest.raises(RuntimeError):
        a[jit.thunk(np.arange(3))]                        # the gather is eager: without a GPU it fails loudly
    with pytest.raises(TypeError):
        a * jit.thunk(np.arange(7))                       # "Incompatible cardinality." (the check Q9 broke)
    with pytest.raises(OverflowError):
        jit.thunk(np.arange(100, dtype=np.int8)) * 1000   # a Python int that does not fit the column type (NumPy 2)
    with pytest.raises(TypeError):
        (a >= 3) * a
    with pytest.raises(TypeError):
        (a >= 3).sum()
    with pytest.raises(TypeError):
        a & a


def plan_lines(t):
    return jit.plan(t).splitlines()


def test_parser_fuses_a_vectorizable_chain_into_one_pipeline():
    a, b, c, d, e, f = (jit.thunk(v) for v in vec(k=6))
    res2 = (d * e * f) * (a * b * c)
    text = jit.plan(res2)
    assert text.count("pipeline Pipe") == 1 and "static:mul6" in text
    assert "inputs=6 outputs=1" in text


def test_parser_materialises_held_intermediates():
    # parser.cpp:63-79: an intermediate that Python still references is stored as well
    a, b, c = (jit.thunk(v) for v in vec(k=3))
    held = a * b
    root = held * c
    text = jit.plan(root)
    assert text.count("pipeline Pipe") == 1 and "outputs=2" in text
    del held
    assert "outputs=1" in jit.plan(root)


def test_parser_splits_at_breakers_like_the_reference():
    # Tests/multiplication.py:26-27: sort is a full breaker, so res, sort(res), (a*b*c) and the
    # final product are four pipelines (parser.cpp:99-145)
    a, b, c, d, e, f = (jit.thunk(v) for v in vec(k=6))
    res = d * e * f
    res2 = res.sort() * (a * b * c)
    lines = [l for l in plan_lines(res2) if l.startswith("pipeline")]
    assert len(lines) == 4
    assert sum("base function sort" in l for l in lines) == 1
    assert sum("static:mul3" in l for l in lines) == 2
    # Tests/max.py:13-14: the sum is a parallel breaker; its parent consumes the size-1 result
    s = (a[a >= 50]).sum()
    r2 = s * a
    lines = [l for l in plan_lines(r2) if l.startswith("pipeline")]
    assert len(lines) == 2 and "static:filter_ge_sum" in lines[0]


def test_filter_pipeline_shape():
    a = jit.thunk(vec())
    text = jit.plan(a[a * 2 >= 50])
    assert "static:filter_mul_ge" in text and "out0: i64 compact" in text
    assert "temps=0" in text and text.count("acc = in0") == 2   # no temporary just to park the column: it is re-read (L1/L2 hit)


def test_scalars_are_baked_as_immediates_of_the_column_type():
    x = jit.thunk(np.arange(10, dtype=np.float32))
    text = jit.plan(x * 2)                        # Python int with a float32 column
    assert "dtype=f32" in text and "in1: f32 imm" in text
    y = jit.thunk(np.arange(10, dtype=np.int32))
    assert "in1: i32 imm" in jit.plan(y >= 3)


def test_evaluation_fails_loudly_without_a_gpu():
    if has_gpu():
        pytest.skip("GPU present")
    a, b = (jit.thunk(v) for v in vec(k=2))
    c = a * b
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        c.evaluate()
    with pytest.raises(RuntimeError):
        str(c)
    with pytest.raises(RuntimeError):
        a[0] = 1000
    assert not c.isevaluated()


def test_children_do_not_dangle():
    # the reference keeps raw, un-INCREF'd child pointers (thunk.cpp:68); dropping a dependent
    # must not leave a stale entry behind
    a, b = (jit.thunk(v) for v in vec(k=2))
    for _ in range(100):
        c = a * b
        del c
    d = a * b
    assert "static:mul2" in jit.plan(d)


def test_a_thunk_reached_twice_through_breakers_gets_one_pipeline():
    # every occurrence used to get a child pipeline of its own, and the second one ran
    # on operator records the first had already released
    a, b = (jit.thunk(v) for v in vec(k=2))
    m = a.max()
    y = (a - m) * (b - m)
    lines = [l for l in plan_lines(y) if l.startswith("pipeline")]
    assert len(lines) == 2 and "static:reduce_max" in lines[0]
    x = a * b
    z = x.sort() * x
    lines = [l for l in plan_lines(z) if l.startswith("pipeline")]
    assert len(lines) == 3                                 # x, sort(x), sort(x) * x
    assert sum("base function sort" in l for l in lines) == 1
    assert sum("static:mul2" in l for l in lines) == 1
    s = (a * b).sum()
    d = a - s
    lines = [l for l in plan_lines(d * d) if l.startswith("pipeline")]
    assert len(lines) == 2


def test_materialisation_counts_references_from_outside_the_dag_only():
    # the reference tests ob_refcnt > 1 (parser.cpp:66); a thunk used twice INSIDE one expression is
    # not "held by somebody else"
    a = jit.thunk(vec())
    d = a - 1
    r = d * d
    assert "outputs=2" in jit.plan(r)          # `d` is still a Python variable
    del d
    assert "outputs=1" in jit.plan(r)          # both references are the expression's own
    for _ in range(3):                         # parsing takes and drops its own references
        assert "outputs=1" in jit.plan(r)


def test_a_column_meeting_a_filter_result_is_split_not_rejected():
    # a[m] * b[m] and chained lazy filters are legal NumPy; the plan compiler asks
    # for the filters to get pipelines of their own (NL_ERR_SPLIT_FILTER) instead of failing
    a, b = (jit.thunk(v) for v in vec(k=2))
    m = a > 5
    assert "will be split" in jit.plan(a[m] * b[m])
    f = a[a > 5]
    assert "will be split" in jit.plan(f[f < 10])


def test_numpy_promotion_builds_lazy_conversions():
    # two columns of different types meet in NumPy's promoted type; the narrower one goes
    # through a lazy astype() (a pass of its own), scalars are converted on the host
    a = jit.thunk(vec())                                   # int64
    f = jit.thunk(np.arange(100.0))
    assert jit.plan(a * f).count("base function astype") == 1
    assert "astype" in jit.plan(a * 2.5)                   # int column, Python float: float64 like NumPy
    assert "astype" not in jit.plan(f * 2) and "in1: f64 imm" in jit.plan(f * 2)
    assert jit.plan(a / a).count("base function astype") == 2      # int / int is a float64: both operands are converted
    assert "astype" not in jit.plan(f / 2)
    u = jit.thunk(np.arange(100, dtype=np.uint8))
    i = jit.thunk(np.arange(100, dtype=np.int8))
    assert jit.plan(u * i).count("base function astype") == 2      # uint8 * int8 -> int16
    assert "dtype=u8" in jit.plan(u * 3) and "dynamic" in jit.plan(u * 3)      # the new types run on the interpreter
    assert "dtype=u8" in jit.plan(u[u > 7]) and "dtype=i8" in jit.plan((-i).max())
    assert a.astype(np.int64) is a                          # a no-op conversion is the thunk itself
    assert not a.astype(np.float32).isevaluated()
    with pytest.raises(TypeError):
        a.astype(np.bool_)
