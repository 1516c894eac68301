#!/usr/bin/env python
"""The three compaction kernels (plain, staged TMA pipeline, super-tile) plus a fused filter->reduce and the
checksum / cast kernels, once each on small inputs, results checked against NumPy — the program to run under
`compute-sanitizer --tool memcheck|racecheck`.

  python tools/sanitize_compact.py            # plain run, must exit 0 first
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from numpyllvm_b200 import abi  # noqa: E402
from numpyllvm_b200.abi import Expr  # noqa: E402
from numpyllvm_b200.device import Runtime, run_pipeline  # noqa: E402

rt = Runtime.get(1)
lib = rt.lib
rng = np.random.default_rng(1)


def filt(dtype, n, thr):
    x = rng.random(n).astype(dtype)
    e = Expr(dtype)
    xa = e.array()
    e.materialize(e.filter(xa, e.gt(e.ref(xa), e.scalar(thr))))
    (got,) = run_pipeline(rt, e, [x, None], n)
    assert np.array_equal(got, x[x > dtype(thr)]), rt.last_launch()[0]
    return rt.last_launch()[0]


# super-tile kernel: at least one super-tile (4 x 4096 f64) per SM
n_super = 256 * 8 * 2 * 4 * 148 + 777
for tier, label in ((1, "static"), (2, "family"), (0, "interpreter")):
    abi.check(lib.nl_set_static_kernels(tier))
    abi.check(lib.nl_tune(b"compact_kernel", 0))
    print(label, "super :", filt(np.float64, n_super, 0.5))
    abi.check(lib.nl_tune(b"compact_kernel", 1))
    print(label, "plain :", filt(np.float64, 300_001, 0.3))
abi.check(lib.nl_set_static_kernels(1))
abi.check(lib.nl_tune(b"compact_kernel", 2))
print("staged pipeline:", filt(np.float64, 512 * 8 * 2 * 2 * 148 + 4097, 0.9))
abi.check(lib.nl_tune(b"reset", 0))

# value operators after the filter on the family kernel, filter -> reduce, checksum, cast
x = rng.random(n_super)
e = Expr(np.float64)
xa = e.array()
e.materialize(e.filter(e.add(e.mul(xa, e.scalar(2.0)), e.scalar(1.0)), e.gt(e.ref(xa), e.scalar(0.9))))
(got,) = run_pipeline(rt, e, [x, None, None, None], n_super)
assert np.array_equal(got, x[x > 0.9] * 2.0 + 1.0)
print("family fc_v2   :", rt.last_launch()[0])
e = Expr(np.float64)
xa = e.array()
e.materialize(e.max(e.filter(xa, e.le(e.ref(xa), e.scalar(0.25)))))
(got,) = run_pipeline(rt, e, [x, None], n_super)
assert got[0] == x[x <= 0.25].max()
d = rt.upload(x)
out = rt.empty(np.uint64, 3)
abi.check(lib.nl_checksum(0, 0, C.c_void_p(d.ptr), x.nbytes, C.c_void_p(out.ptr)))
w = x.view(np.uint64)
assert int(rt.download(out)[0]) == int(np.add.reduce(w))
y = rt.empty(np.float32, x.size)
abi.check(lib.nl_cast(0, 0, C.c_void_p(y.ptr), abi.NL_F32, C.c_void_p(d.ptr), abi.NL_F64, x.size))
assert np.array_equal(rt.download(y), x.astype(np.float32))
print("sanitize_compact ok")
