#!/usr/bin/env python
"""Time a[idx] (integer-array indexing, nl_gather) through the jit API:  python tools/gather_bench.py [log2n=27]
With several GPUs visible the column and the index are sharded and remote elements come over NVLink."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpyllvm_b200  # noqa: E402

jit = numpyllvm_b200.install_jit_alias()
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 27
n = 1 << log2n
rng = np.random.default_rng(1)
a = jit.thunk(rng.random(n))
for label, idx_n in (("random", rng.integers(0, n, n)), ("sequential", np.arange(n, dtype=np.int64))):
    idx = jit.thunk(idx_n)
    a[idx]                                   # warm-up
    best = 1e9
    for _ in range(5):
        t0 = time.perf_counter()
        g = a[idx]                           # eager: returns after the devices have finished
        best = min(best, time.perf_counter() - t0)
    print(f"{label:10s} n=2^{log2n} on {jit.device_count()} GPU(s): {best * 1e3:8.3f} ms  {n / best / 1e9:6.2f} Gelem/s  "
          f"{24 * n / best / 1e9:7.0f} GB/s (8 B index + 8 B element + 8 B result)")
    assert np.array_equal(g.asnumpyarray()[:1000], a.asnumpyarray()[idx_n[:1000]])
