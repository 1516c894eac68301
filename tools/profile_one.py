#!/usr/bin/env python
"""Run ONE pipeline config a few times (for ncu / quick timing).

  python tools/profile_one.py c2 --log2n 27 --reps 3
Configs: c1 c2 c3_f64_max c3_f64_sum c3_i64_max c3_i64_sum c4_1 c4_10 c4_50 c4_90 c4b_1 c4b_50 c4b_90 c5
Prints the CUDA-event time per launch and the implied GB/s (never quote a number
taken under a profiler).
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from numpyllvm_b200 import abi  # noqa: E402
from numpyllvm_b200.abi import Expr  # noqa: E402
from numpyllvm_b200.device import Runtime  # noqa: E402


def build(name, rt, n):
    """returns (plan, out_ptrs, in_ptrs, bytes_per_launch, keepalive)"""
    sizes = rt.empty(np.int64, 4)
    if name == "c2":
        a, b, c = (rt.empty(np.float32, n) for _ in range(3))
        rt.fill(a, 0, 42, 2); rt.fill(b, 0, 43, 2)
        e = Expr(np.float32); e.materialize(e.mul(e.array(), e.array()))
        return abi.compile_plan(e), [c.ptr], [a.ptr, b.ptr], 12 * n, (a, b, c, sizes)
    if name == "c1":
        cols = [rt.empty(np.float64, n) for _ in range(7)]
        for j in range(6):
            rt.fill(cols[j], 0, 100 + j, 2)
        e = Expr(np.float64)
        d_, e_, f_, a_, b_, c_ = (e.array() for _ in range(6))
        e.materialize(e.mul(e.mul(e.mul(d_, e_), f_), e.mul(e.mul(a_, b_), c_)))
        return abi.compile_plan(e), [cols[6].ptr], [x.ptr for x in cols[:6]], 56 * n, (cols, sizes)
    if name.startswith("c3_"):
        _, dt, op = name.split("_")
        dtype = np.float64 if dt == "f64" else np.int64
        x = rt.empty(dtype, n); rt.fill(x, 0, 7, 3 if dt == "f64" else 1)
        res = rt.empty(dtype, 1)
        e = Expr(dtype); e.materialize(e.unop(abi.OP_MAX if op == "max" else abi.OP_SUM, e.array()))
        return abi.compile_plan(e), [res.ptr], [x.ptr], 8 * n, (x, res, sizes)
    if name.startswith("c4_"):
        sel = int(name.split("_")[1]) / 100.0
        x = rt.empty(np.float64, n); rt.fill(x, 0, 9, 3)
        y = rt.empty(np.float64, n)
        e = Expr(np.float64); xa = e.array()
        e.materialize(e.filter(xa, e.gt(e.ref(xa), e.scalar(1.0 - sel))))
        return abi.compile_plan(e), [y.ptr], [x.ptr, None], int(8 * n * (1 + sel)), (x, y, sizes)
    if name.startswith("c4d_"):          # (a*2+1)[a > t]: one staged column, no build-time shape
        sel = int(name.split("_")[1]) / 100.0
        a = rt.empty(np.float64, n); rt.fill(a, 0, 9, 3)
        y = rt.empty(np.float64, n)
        e = Expr(np.float64); xa = e.array()
        e.materialize(e.filter(e.add(e.mul(xa, e.scalar(2.0)), e.scalar(1.0)), e.gt(e.ref(xa), e.scalar(1.0 - sel))))
        plan = abi.compile_plan(e)
        return plan, [y.ptr], [a.ptr if plan.in_[i].kind == abi.IN_ARRAY else None for i in range(plan.n_inputs)], int(8 * n * (1 + sel)), (a, y, sizes)
    if name.startswith("c4r_"):          # a[(a > lo) & (a < hi)]: the range filter, selectivity = hi - lo
        sel = int(name.split("_")[1]) / 100.0
        a = rt.empty(np.float64, n); rt.fill(a, 0, 9, 3)
        y = rt.empty(np.float64, n)
        e = Expr(np.float64); xa = e.array()
        e.materialize(e.filter(xa, e.and_(e.gt(e.ref(xa), e.scalar(0.25)), e.lt(e.ref(xa), e.scalar(0.25 + sel)))))
        return abi.compile_plan(e), [y.ptr], [a.ptr, None, None], int(8 * n * (1 + sel)), (a, y, sizes)
    if name.startswith("c4b_"):          # b[a > t]: two staged columns
        sel = int(name.split("_")[1]) / 100.0
        a = rt.empty(np.float64, n); rt.fill(a, 0, 9, 3)
        b = rt.empty(np.float64, n); rt.fill(b, 0, 10, 2)
        y = rt.empty(np.float64, n)
        e = Expr(np.float64); xa, xb = e.array(), e.array()
        e.materialize(e.filter(xb, e.gt(xa, e.scalar(1.0 - sel))))
        return abi.compile_plan(e), [y.ptr], [a.ptr, b.ptr, None], int(8 * n * (2 + sel)), (a, b, y, sizes)
    if name == "c5":
        x = rt.empty(np.float32, n); rt.fill(x, 0, 11, 3)
        res = rt.empty(np.float32, 1)
        e = Expr(np.float32); xa = e.array()
        e.materialize(e.max(e.mul(e.filter(xa, e.ge(e.ref(xa), e.scalar(0.5))), e.scalar(3.0))))
        return abi.compile_plan(e), [res.ptr], [x.ptr, None, None], 4 * n, (x, res, sizes)
    raise SystemExit(f"unknown config {name}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--log2n", type=int, default=27)
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--trend", action="store_true", help="time every launch and print the trend over the repetitions")
    ap.add_argument("--tier", type=int, default=1, help="nl_set_static_kernels(): 1 all tiers, 2 families + interpreter, 0 interpreter")
    ap.add_argument("--tune", action="append", default=[], help="nl_tune key=value (repeatable), e.g. compact_kernel=1")
    args = ap.parse_args()
    rt = Runtime.get(1)
    rt.lib.nl_set_static_kernels(args.tier)
    for kv in args.tune:
        key, val = kv.split("=")
        abi.check(rt.lib.nl_tune(key.encode(), int(val)))
    n = args.n or (1 << args.log2n)
    plan, outs, ins, nbytes, keep = build(args.config, rt, n)
    sizes = keep[-1]
    for _ in range(args.warmup):
        rt.launch(plan, outs, ins, 0, n, sizes.ptr)
    rt.sync()
    if args.trend:
        evs = [rt.event() for _ in range(args.reps + 1)]
        rt.record(evs[0])
        for i in range(args.reps):
            rt.launch(plan, outs, ins, 0, n, sizes.ptr)
            rt.record(evs[i + 1])
        rt.sync()
        t = [rt.elapsed_ms(evs[i], evs[i + 1]) for i in range(args.reps)]
        picks = sorted(set(i for i in (0, 1, 2, 5, 10, 20, 40, 80, 120, 160, 199, 299, 399) if i < args.reps))
        print("  per-launch ms:", " ".join(f"[{i}]{t[i]:.4f}" for i in picks), f" min {min(t):.4f} max {max(t):.4f}")
    e0, e1 = rt.event(), rt.event()
    rt.record(e0)
    for _ in range(args.reps):
        rt.launch(plan, outs, ins, 0, n, sizes.ptr)
    rt.record(e1)
    rt.sync()
    ms = rt.elapsed_ms(e0, e1) / args.reps
    name, grid, block, vec = rt.last_launch()
    print(f"{args.config} n={n} {name} grid={grid}: {ms:.4f} ms/launch, {nbytes / ms / 1e6:.1f} GB/s algorithmic")


if __name__ == "__main__":
    main()
