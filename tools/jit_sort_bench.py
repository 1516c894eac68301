#!/usr/bin/env python
"""sort() through the jit API over all visible GPUs (one process, NL_DEVICES), device-resident input:
   python tools/jit_sort_bench.py [log2 keys per GPU = 26] [dtype = int64]
Prints keys/s of `a.sort(); evaluate` (the copy, the cut at the splitters, the exchange over NVLink, the per-GPU
radix sort and the re-cut into the regular partition all inside) and checks the result against NumPy at 2^22 keys."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from numpyllvm_b200 import jit  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
dtype = np.dtype(sys.argv[2] if len(sys.argv) > 2 else "int64")
rng = np.random.default_rng(1)


def draw(n):
    if dtype.kind == "f":
        return ((rng.random(n) - 0.5) * 1e9).astype(dtype)
    info = np.iinfo(dtype)
    return rng.integers(info.min, info.max, size=n, dtype=dtype, endpoint=True)


x = draw(1 << 22)
t = jit.thunk(x)
G = jit.device_count()
got = t.sort().asnumpyarray()
assert got.tobytes() == np.sort(x).tobytes(), "jit sort differs from numpy.sort"
n = G << log2n
x = draw(n)
t = jit.thunk(x)
t.evaluate()
jit.synchronize()
best = 1e30
for rep in range(4):
    t0 = time.perf_counter()
    s = t.sort()
    s.evaluate()
    jit.synchronize()
    dt = time.perf_counter() - t0
    best = min(best, dt)
    if rep == 0:
        head = s.asnumpyarray()
        assert np.all(head[:-1] <= head[1:]) and head.size == n, "not sorted"
        del head
    del s
print(f"jit sort {dtype.name} over {G} GPU(s): {n} keys ({n >> 20} Mi) in {best * 1e3:.2f} ms = {n / best / 1e9:.2f} Gkeys/s "
      f"({n / best / 1e9 / G:.2f} per GPU)")
