#!/usr/bin/env python
"""Throughput of expressions that have NO build-time shape (they run on the step interpreter),
beside the same expression forced through the interpreter when a shape exists.
  python tools/interp_bench.py [log2n=28]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from numpyllvm_b200 import abi  # noqa: E402
from numpyllvm_b200.abi import Expr  # noqa: E402
from numpyllvm_b200.device import Runtime  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
n = 1 << log2n
rt = Runtime.get(1)
sizes = rt.empty(np.int64, 4)


def cases(dt):
    def axpy(e):
        a, b = e.array(), e.array()
        return e.add(e.mul(a, e.scalar(dt(3))), b), 2
    def fma4(e):
        a, b, c, d = (e.array() for _ in range(4))
        return e.add(e.mul(a, b), e.mul(c, d)), 4
    def poly(e):
        x = e.array()
        return e.add(e.mul(e.add(e.mul(e.ref(x), e.scalar(dt(2))), e.scalar(dt(1))), e.ref(x)), e.scalar(dt(5))), 1
    def sum3(e):
        a, b, c = (e.array() for _ in range(3))
        return e.sub(e.add(a, b), c), 3
    def mul2(e):
        a, b = e.array(), e.array()
        return e.mul(a, b), 2
    def div2(e):
        a, b = e.array(), e.array()
        return e.div(a, b), 2
    return [("a*3+b", axpy), ("a*b+c*d", fma4), ("(x*2+1)*x+5", poly), ("a+b-c", sum3), ("a*b", mul2), ("a/b", div2)]


for dt in (np.float32, np.float64):
    es = np.dtype(dt).itemsize
    for name, build in cases(dt):
        e = Expr(dt)
        root, ncols = build(e)
        e.materialize(root)
        plan = abi.compile_plan(e)
        cols = [rt.empty(dt, n) for _ in range(ncols)]
        for j, c in enumerate(cols):
            rt.fill(c, 0, 40 + j, 2)
        out = rt.empty(dt, n)
        ins = []
        k = 0
        for i in range(plan.n_inputs):
            if plan.in_[i].kind == abi.IN_ARRAY:
                ins.append(cols[k].ptr); k += 1
            else:
                ins.append(None)
        row = f"{np.dtype(dt).name} {name:14s}"
        for flags, label in ((1, "default"), (0, "interpreter")):
            rt.lib.nl_set_static_kernels(flags)
            for _ in range(2):
                rt.launch(plan, [out.ptr], ins, 0, n, sizes.ptr)
            rt.sync()
            e0, e1 = rt.event(), rt.event()
            rt.record(e0)
            for _ in range(10):
                rt.launch(plan, [out.ptr], ins, 0, n, sizes.ptr)
            rt.record(e1)
            rt.sync()
            ms = rt.elapsed_ms(e0, e1) / 10
            kname = rt.last_launch()[0]
            row += f" | {label}: {(ncols + 1) * es * n / ms / 1e6:7.0f} GB/s ({'static' if 'static' in kname else 'dyn'})"
        print(row)
        for c in cols:
            c.free()
        out.free()
rt.lib.nl_set_static_kernels(1)
