#!/usr/bin/env python
"""sort() through the jit API over all visible GPUs (one process, NL_DEVICES), device-resident input:
   python tools/jit_sort_bench.py [log2 keys per GPU = 26] [dtype = int64] [--json]
Times `s = a.sort(); s.evaluate()` — the cut at the splitters, the exchange over NVLink, the per-GPU radix sort and the
re-cut into the regular partition all inside — checks a 2^22-key sort bit for bit against numpy.sort and the full-size
result by its properties (no inversion, same length, same wrapping sum / xor of the bit patterns as the input).
bench.py runs this in a process of its own for its "sort xN" row (its ranks own one GPU each; sort() across GPUs is
the one-process mode of the jit drop-in)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from numpyllvm_b200 import jit  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
as_json = "--json" in sys.argv
log2n = int(args[0]) if len(args) > 0 else 26
dtype = np.dtype(args[1] if len(args) > 1 else "int64")
rng = np.random.default_rng(1)


def draw(n):
    if dtype.kind == "f":
        return ((rng.random(n) - 0.5) * 1e9).astype(dtype)
    info = np.iinfo(dtype)
    return rng.integers(info.min, info.max, size=n, dtype=dtype, endpoint=True)


def bits(a):
    return a.view(np.dtype(f"u{a.dtype.itemsize}"))


x = draw(1 << 22)
t = jit.thunk(x)
G = jit.device_count()
small_ok = t.sort().asnumpyarray().tobytes() == np.sort(x).tobytes()
n = G << log2n
x = draw(n)
t = jit.thunk(x)
t.evaluate()
jit.synchronize()
best = 1e30
props = {}
for rep in range(4):
    l0 = jit.launch_count()
    t0 = time.perf_counter()
    s = t.sort()
    s.evaluate()
    jit.synchronize()
    dt = time.perf_counter() - t0
    best = min(best, dt)
    launches = jit.launch_count() - l0
    if rep == 0:
        y = s.asnumpyarray()
        with np.errstate(over="ignore"):
            props = {"length": int(y.size) == n, "inversions": int(np.count_nonzero(y[1:] < y[:-1])),
                     "sum_unchanged": bool(bits(y).sum(dtype=np.uint64) == bits(x).sum(dtype=np.uint64)),
                     "xor_unchanged": bool(np.bitwise_xor.reduce(bits(y)) == np.bitwise_xor.reduce(bits(x)))}
        del y
    del s
good = small_ok and props["length"] and props["inversions"] == 0 and props["sum_unchanged"] and props["xor_unchanged"]
row = {"id": f"sort x{G}", "config": f"sort() through the jit API over {G} GPU(s) in one process: sample sort over NVLink, {dtype.name} full range",
       "dtype": dtype.name, "elements_per_gpu": 1 << log2n, "n_gpus": G, "ms_per_step": round(best * 1e3, 3),
       "Gkeys/s": round(n / best / 1e9, 2), "Gkeys/s_per_gpu": round(n / best / 1e9 / G, 2), "kernel_launches": int(launches),
       "timing": "host wall clock around evaluate() + synchronize(), best of 4",
       "properties_at_full_size": props,
       "parity": ("bit-exact: " if good else "MISMATCH: ") + "2^22 keys identical to numpy.sort; full size: zero inversions, same length, unchanged sum/xor of the bit patterns"}
if as_json:
    print(json.dumps(row))
else:
    print(f"jit sort {dtype.name} over {G} GPU(s): {n} keys ({n >> 20} Mi) in {best * 1e3:.2f} ms = {n / best / 1e9:.2f} Gkeys/s "
          f"({n / best / 1e9 / G:.2f} per GPU), {launches} launches, {row['parity']}")
