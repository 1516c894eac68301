#!/usr/bin/env python
"""Time nl_sort (device radix sort) on synthetic columns:  python tools/sort_bench.py [log2n=28]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from numpyllvm_b200 import abi  # noqa: E402
from numpyllvm_b200.device import Runtime  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
algo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = 1 << log2n
rt = Runtime.get(1)
abi.check(rt.lib.nl_tune(b"sort_algo", algo))
print(f"sort_algo={algo}")
for dtype, kind, label in ((np.int64, 1, "i64 full range"), (np.int64, 0, "i64 in [1,100)"), (np.float64, 3, "f64 [0,1)"),
                           (np.int32, 1, "i32 full range"), (np.float32, 3, "f32 [0,1)")):
    x = rt.empty(dtype, n)
    tmp = rt.empty(dtype, n)
    best = 1e9
    for rep in range(3):
        rt.fill(x, 0, 5 + rep, kind)
        rt.sync()
        before = rt.lib.nl_launch_count()
        e0, e1 = rt.event(), rt.event()
        rt.record(e0)
        abi.check(rt.lib.nl_sort(0, 0, abi.nl_dtype_of(dtype), C.c_void_p(x.ptr), C.c_void_p(tmp.ptr), n))
        rt.record(e1)
        rt.sync()
        best = min(best, rt.elapsed_ms(e0, e1))
        launches = rt.lib.nl_launch_count() - before
    per_pass = 3 if algo else max(1, -(-n // (1 << 28)))       # one launch per 2^28-key portion and position
    passes = (launches - 1 - (0 if algo else 1)) // per_pass
    es = np.dtype(dtype).itemsize
    traffic = n * es * (1 + (3 if algo else 2) * passes + (2 if passes % 2 else 0))
    head = rt.download(x, 4096)
    assert np.all(head[:-1] <= head[1:])
    print(f"{label:16s} n=2^{log2n}: {best:8.3f} ms  {n / best / 1e6:7.2f} Gkeys/s  {passes} passes  {traffic / best / 1e6:7.0f} GB/s of pass traffic")
    x.free(); tmp.free()
