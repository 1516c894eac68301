import sys, numpy as np
sys.path.insert(0, "/root/repo")
from numpyllvm_b200 import abi
from numpyllvm_b200.abi import Expr
from numpyllvm_b200.device import Runtime
rt = Runtime.get(1)
n = 1 << 30
x = rt.empty(np.float64, n); rt.fill(x, 0, 9, 3)
y = rt.empty(np.float64, n)
sizes = rt.empty(np.int64, 4)
for rnd in range(2):
    for sel in (0.01, 0.10, 0.50, 0.90):
        e = Expr(np.float64); xa = e.array()
        e.materialize(e.filter(xa, e.gt(e.ref(xa), e.scalar(1.0 - sel))))
        plan = abi.compile_plan(e)
        for _ in range(25): rt.launch(plan, [y.ptr], [x.ptr, None], 0, n, sizes.ptr)
        rt.sync()
        e0, e1 = rt.event(), rt.event(); rt.record(e0)
        for _ in range(20): rt.launch(plan, [y.ptr], [x.ptr, None], 0, n, sizes.ptr)
        rt.record(e1); rt.sync()
        ms = rt.elapsed_ms(e0, e1) / 20
        print(rnd, sel, rt.last_launch()[:2], round(ms, 4), "ms", round(8 * n * (1 + sel) / ms / 1e6), "GB/s", flush=True)
