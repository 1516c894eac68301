#!/usr/bin/env python
"""bench.py — effective HBM GB/s (and elements/s) of numpyllvm's fused thunk pipelines on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Headline workload (BASELINE.json configs[1]): Tests/assignment.py at scale —
`c = a * b` pending on 2^30-element f32 columns, `a[0] = 1000` forces it
(thunk_as_mapping.cpp:80-86), one step = the forced fused pipeline (12 B/elem)
plus the in-place update.  Weak scaling: every GPU owns a 2^30-element shard
and there is no data-path collective.  The other BASELINE.json configs are
reported in `pipelines` (N = 1: C1, C3, C4, C5; N > 1: C3/C4/C5 with their NCCL
finish).  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "effective HBM GB/s per fused pipeline"
FALLBACK_PEAK_GBS = 6650.0     # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy_)"
    except Exception:
        return FALLBACK_PEAK_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0: float, t1: float):
        rows = [(t, l) for t, l in self.samples if t0 <= t <= t1 + 0.06]
        note = None
        if not rows and self.samples:   # region shorter than the sampling period: nearest samples
            rows = sorted(self.samples, key=lambda s: abs(s[0] - (t0 + t1) / 2))[:3]
            note = "timed region shorter than the 50 ms sampling period; nearest samples"
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for _, l in rows:
            parts = [p.strip() for p in l.split(",")]
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except Exception:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": sorted(reasons), "samples": len(sm)}
        if note:
            out["note"] = note
        return out


# ------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    from numpyllvm_b200 import abi
    from numpyllvm_b200.abi import Expr
    from numpyllvm_b200.device import Runtime
    from numpyllvm_b200.dist import Ranks

    # the JSON line is the only thing this arm may print on stdout: keep NCCL's banner off it
    os.environ["NCCL_DEBUG"] = os.environ.get("NL_NCCL_DEBUG", "WARN")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")    # the version banner must not precede the JSON line on stdout
    ranks = Ranks(args.gpus)
    lib = abi.load_library()
    dev_ids = (C.c_int * 1)(ranks.local_rank)
    abi.check(lib.nl_init(1, dev_ids))          # one rank drives one GPU
    rt = Runtime.get(1)
    ranks.init_library_comm(lib)                # NCCL communicator for the data-path collectives
    peak, peak_src = measured_peak()
    G = ranks.world
    n = 1 << args.log2n           # elements per GPU (weak scaling)
    K, W = args.steps, max(args.warmup, 3)

    def timed(step_fn, k, w, per_launch=None):
        """w warm-up + k timed steps, barrier + device sync on both sides, CUDA events on the
        launching stream, max over ranks.  Returns (ms total, [per-launch ms])."""
        for _ in range(w):
            step_fn(None)
        rt.sync()
        ranks.barrier()
        e0, e1 = rt.event(), rt.event()
        pairs = []
        rt.record(e0)
        for _ in range(k):
            if per_launch:
                a_, b_ = rt.event(), rt.event()
                pairs.append((a_, b_))
                step_fn((a_, b_))
            else:
                step_fn(None)
        rt.record(e1)
        rt.sync()
        ms = rt.elapsed_ms(e0, e1)
        launch_ms = [rt.elapsed_ms(a_, b_) for a_, b_ in pairs]
        for ev in [e0, e1] + [x for p in pairs for x in p]:
            lib.nl_event_destroy(ev)
        ranks.barrier()
        return ranks.max(ms), launch_ms

    # ---------------- headline: C2 assignment (f32, 2^30 per GPU) ----------------
    f32 = np.float32
    a, b, c = rt.empty(f32, n), rt.empty(f32, n), rt.empty(f32, n)
    rt.fill(a, ranks.rank * n, 42, 2)
    rt.fill(b, ranks.rank * n, 43, 2)
    sizes = rt.empty(np.int64, 4)
    e = Expr(f32)
    e.materialize(e.mul(e.array(), e.array()))
    plan = abi.compile_plan(e)
    thousand = abi.scalar_bits(1000, f32)

    def c2_step(evs):
        # c = a*b is pending; a[0] = 1000 forces it, then updates a in place (Tests/assignment.py:14-16)
        if evs:
            rt.record(evs[0])
        rt.launch(plan, [c.ptr], [a.ptr, b.ptr], 0, n, sizes.ptr)
        if evs:
            rt.record(evs[1])
        if ranks.rank == 0:
            abi.check(lib.nl_assign_fill(0, 0, C.c_void_p(a.ptr), abi.NL_F32, 0, 1, 1, thousand))

    sampler = ClockSampler(ranks.local_rank)
    time.sleep(0.12)
    launches0 = lib.nl_launch_count()
    for _ in range(W):
        c2_step(None)
    rt.sync()
    launches_warm = lib.nl_launch_count()
    t0 = time.time()
    ms_total, launch_ms = timed(c2_step, K, 0, per_launch=True)
    t1 = time.time()
    gpu_launches = lib.nl_launch_count() - launches_warm
    clocks = sampler.summary(t0, t1)
    sampler.stop()                              # nvidia-smi polling perturbs latency-sensitive kernels: headline region only
    alg_bytes = 12 * n + 4                     # 2 f32 reads + 1 f32 write per element + the 4-byte update
    value = G * alg_bytes * K / (ms_total * 1e-3) / 1e9
    kern_ms = statistics.mean(launch_ms)
    achieved = 12 * n / (kern_ms * 1e-3) / 1e9
    kname, kgrid, kblock, kvec = rt.last_launch()

    # spot-check parity on sampled windows against the oracle (outside the timed region)
    parity = "unchecked"
    if ranks.rank == 0 and not args.no_check:
        import oracle
        ok = True
        for w0 in (0, n // 3 + 5, n - 4096):
            wa = oracle.fill(f32, w0, 4096, 42, 2)
            wb = oracle.fill(f32, w0, 4096, 43, 2)
            if w0 == 0 and K + W > 1:
                wa[0] = 1000.0            # every step after the first sees the updated a[0]
            got = rt.download(c, 4096, offset_elems=w0)
            ok &= bool(np.array_equal(got, wa * wb))
        ok &= float(rt.download(a, 1)[0]) == 1000.0
        parity = "bit-exact on sampled windows" if ok else "MISMATCH"

    # ---------------- e2e: same step with HOST buffers, copies inside the timed region ------------
    e2e = e2e_streamed = None
    if not args.no_e2e:
        e2e_streamed = run_e2e_cabi(args, rt, lib, ranks, plan, a, b, c, sizes, n, thousand)
        a.free(); b.free(); c.free()
        e2e = run_e2e_jit(args, ranks, n)

    result = {
        "metric": METRIC, "value": round(value, 2), "unit": "GB/s", "n_gpus": G, "steps": K, "warmup": W,
        "ms_per_step": round(ms_total / K, 5), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Tests/assignment.py at scale: c = a*b forced by a[0] = 1000, then the in-place update",
                   "elements_per_gpu": n, "algorithmic_bytes_per_step_per_gpu": alg_bytes,
                   "l2": "inputs (12 GiB per GPU) are larger than L2; no flush needed",
                   "parallelism": f"{G} contiguous shards, one rank per GPU, no data-path collective"},
        "elements_per_s": round(G * n * K / (ms_total * 1e-3), 1),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": load_traffic("c2"),
                     "kernel": kname, "grid": kgrid, "block": kblock, "peak_source": peak_src,
                     "frac_of_nominal_8000": round(achieved / 8000.0, 4),
                     "avg_launch_ms": round(kern_ms, 5)},
        "gpu_launches": int(gpu_launches),
        "clocks": clocks,
        "parity": parity,
    }
    if e2e:
        result["e2e"] = e2e
        result["e2e_streamed_cabi"] = e2e_streamed
    a.free(); b.free(); c.free()

    # ---------------- the other BASELINE.json configs ----------------
    if not args.no_extra:
        try:
            result["pipelines"] = run_extra(args, rt, lib, ranks, peak, timed)
        except Exception as exc:                    # the extra configs must never cost the headline line
            result["pipelines_error"] = f"{type(exc).__name__}: {exc}"

    # ---------------- CPU baseline beside it (rank 0, N = 1 only) ----------------
    if ranks.rank == 0 and G == 1 and not args.no_cpu:
        result["cpu_baseline"] = cpu_baseline(args)
    sampler.stop()
    if ranks.rank == 0:
        print(json.dumps(result), flush=True)
    ranks.finish()


def load_traffic(key):
    """dram bytes read+written per launch from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh).get(key)
    except Exception:
        return None


def run_e2e_cabi(args, rt, lib, ranks, plan, a, b, c, sizes, n, thousand):
    """Extra data point: the same step through the C-ABI with HOST (pinned) buffers: H2D of both inputs, the
    fused pipeline, the in-place update and D2H of the result, chunked so the upload, the
    kernel and the download overlap on the three streams of the device."""
    from numpyllvm_b200 import abi
    nbytes = 4 * n
    hp = [C.c_void_p() for _ in range(3)]
    for p in hp:
        abi.check(lib.nl_host_alloc(nbytes, C.byref(p)))
    ha = np.ctypeslib.as_array(C.cast(hp[0], C.POINTER(C.c_float)), shape=(n,))
    hb = np.ctypeslib.as_array(C.cast(hp[1], C.POINTER(C.c_float)), shape=(n,))
    hc = np.ctypeslib.as_array(C.cast(hp[2], C.POINTER(C.c_float)), shape=(n,))
    # host inputs: the same synthetic columns (fetched once from the device copy, untimed)
    rt.fill(a, ranks.rank * n, 42, 2)
    abi.check(lib.nl_memcpy_d2h(0, 0, hp[0], C.c_void_p(a.ptr), nbytes))
    abi.check(lib.nl_memcpy_d2h(0, 0, hp[1], C.c_void_p(b.ptr), nbytes))
    rt.sync()
    chunk = min(n, 1 << 24)          # 64 MiB per column per chunk
    nchunks = (n + chunk - 1) // chunk

    def step():
        # the uploads overwrite a/b: they may not start before the previous step's kernels are done
        abi.check(lib.nl_stream_wait(0, 1, 0, 0))
        for i in range(nchunks):
            s, e_ = i * chunk, min(n, (i + 1) * chunk)
            off, nb = 4 * s, 4 * (e_ - s)
            abi.check(lib.nl_memcpy_h2d(0, 1, C.c_void_p(a.ptr + off), C.c_void_p(hp[0].value + off), nb))
            abi.check(lib.nl_memcpy_h2d(0, 1, C.c_void_p(b.ptr + off), C.c_void_p(hp[1].value + off), nb))
            abi.check(lib.nl_stream_wait(0, 0, 0, 1))
            rt.launch(plan, [c.ptr], [a.ptr, b.ptr], s, e_, sizes.ptr)
            abi.check(lib.nl_stream_wait(0, 2, 0, 0))
            abi.check(lib.nl_memcpy_d2h(0, 2, C.c_void_p(hp[2].value + off), C.c_void_p(c.ptr + off), nb))
        if ranks.rank == 0:
            abi.check(lib.nl_assign_fill(0, 0, C.c_void_p(a.ptr), abi.NL_F32, 0, 1, 1, thousand))
        abi.check(lib.nl_stream_wait(0, 0, 0, 2))

    k = max(1, min(args.steps, args.e2e_steps))
    step(); rt.sync()
    ranks.barrier()
    e0, e1 = rt.event(), rt.event()
    t0 = time.perf_counter()
    rt.record(e0)
    for _ in range(k):
        step()
    rt.record(e1)
    rt.sync()
    wall = time.perf_counter() - t0
    ms = ranks.max(max(rt.elapsed_ms(e0, e1), wall * 1e3))
    ok = bool(np.array_equal(hc[:4096], ha[:4096] * hb[:4096])) and bool(np.array_equal(hc[-4096:], ha[-4096:] * hb[-4096:]))
    for p in hp:
        lib.nl_host_free(p)
    G = ranks.world
    return {"value": round(G * (12 * n + 4) * k / (ms * 1e-3) / 1e9, 2), "unit": "GB/s",
            "h2d_bytes_per_step": 2 * nbytes, "d2h_bytes_per_step": nbytes, "steps": k,
            "ms_per_step": round(ms / k, 3), "api": "C-ABI nl_memcpy_h2d / nl_launch_pipeline / nl_memcpy_d2h, pinned host buffers, 64 MiB chunks on 3 streams",
            "parity": "bit-exact on sampled windows" if ok else "MISMATCH"}


def run_e2e_jit(args, ranks, n):
    """The headline end-to-end number: the step exactly as a user of the reference writes it
    (Tests/assignment.py), through the `jit` extension, with HOST buffers.  Every step uploads
    both columns from pinned host memory (jit.thunk copies on wrap), runs the forced pipeline and
    the in-place update, and reads the result back into host memory (asnumpyarray)."""
    import numpyllvm_b200
    jit = numpyllvm_b200.install_jit_alias()
    import oracle
    n_e = n if ranks.world == 1 else min(n, 1 << 28)      # bound the pinned host memory of 8 ranks
    ha, hb = jit.pinned_empty(n_e, np.float32), jit.pinned_empty(n_e, np.float32)
    blk = 1 << 22
    for s0 in range(0, n_e, blk):                          # synthetic columns, generated on the host
        m = min(blk, n_e - s0)
        ha[s0:s0 + m] = oracle.fill(np.float32, ranks.rank * n + s0, m, 42, 2)
        hb[s0:s0 + m] = oracle.fill(np.float32, ranks.rank * n + s0, m, 43, 2)

    def step():
        a = jit.thunk(ha)
        b = jit.thunk(hb)
        c = a * b
        a[0] = 1000
        c.evaluate()
        return c.asnumpyarray()

    out = step()                                           # warm-up: pins the result pool, loads kernels
    ok = bool(np.array_equal(out[:4096], ha[:4096] * hb[:4096])) and bool(np.array_equal(out[-4096:], ha[-4096:] * hb[-4096:]))
    del out
    k = max(1, min(args.steps, args.e2e_steps))
    ranks.barrier()
    t0 = time.perf_counter()
    for _ in range(k):
        out = step()
        del out
    dt = ranks.max(time.perf_counter() - t0)
    ranks.barrier()
    G = ranks.world
    return {"value": round(G * (12 * n_e + 4) * k / dt / 1e9, 2), "unit": "GB/s",
            "h2d_bytes_per_step": 8 * n_e, "d2h_bytes_per_step": 4 * n_e, "steps": k,
            "ms_per_step": round(1e3 * dt / k, 3), "elements_per_gpu": n_e,
            "api": "jit.thunk(pinned ndarray) x2 -> c = a*b -> a[0] = 1000 -> c.evaluate() -> c.asnumpyarray(); wall clock",
            "bound": "PCIe: 12 bytes cross the host link per element for 12 algorithmic bytes",
            "parity": "bit-exact on sampled windows" if ok else "MISMATCH"}


def _reduce_device(rt, abi, Expr, op, dtype, ptr, n, sizes):
    """op over a device column through the reduction pipeline (used for full-size property checks)."""
    res = rt.empty(dtype, 1)
    e = Expr(dtype)
    e.materialize(e.unop(op, e.array()))
    rt.launch(abi.compile_plan(e), [res.ptr], [ptr], 0, n, sizes.ptr)
    v = rt.download(res)[0]
    res.free()
    return v


def filter_properties(rt, abi, Expr, x, y, n, cnt, thr, sizes):
    """Size-independent checks of a[a > t] at full size: every survivor passes the predicate and the
    survivors' sum equals the fused a[a > t].sum() pipeline's (same multiset)."""
    dt = np.float64
    ok_min = bool(cnt == 0 or _reduce_device(rt, abi, Expr, abi.OP_MIN, dt, y.ptr, cnt, sizes) > thr)
    s_out = _reduce_device(rt, abi, Expr, abi.OP_SUM, dt, y.ptr, cnt, sizes) if cnt else 0.0
    res = rt.empty(dt, 1)
    e = Expr(dt)
    xa = e.array()
    e.materialize(e.sum(e.filter(xa, e.gt(e.ref(xa), e.scalar(thr)))))
    rt.launch(abi.compile_plan(e), [res.ptr], [x.ptr, None], 0, n, sizes.ptr)
    s_fused = rt.download(res)[0]
    res.free()
    return {"all_survivors_pass": ok_min, "sum_equals_fused_filter_sum": bool(abs(s_out - s_fused) <= 1e-9 * max(1.0, abs(s_fused)))}


def sort_properties(rt, abi, Expr, lib, x, n, dt, sum_before, sizes):
    """Full-size checks of the device sort: no inversion anywhere (count of y[i+1] < y[i] is zero) and,
    for integers, the wrapping sum is unchanged (same multiset)."""
    es = np.dtype(dt).itemsize
    scratch = rt.empty(dt, n)
    e = Expr(dt)
    nxt, cur = e.array(), e.array()
    e.materialize(e.filter(nxt, e.lt(e.ref(nxt), cur)))
    rt.launch(abi.compile_plan(e), [scratch.ptr], [x.ptr + es, x.ptr], 0, n - 1, sizes.ptr)
    inversions = int(rt.download(sizes, 1)[0])
    scratch.free()
    out = {"inversions": inversions}
    if np.dtype(dt).kind == "i":
        out["sum_unchanged"] = bool(_reduce_device(rt, abi, Expr, abi.OP_SUM, dt, x.ptr, n, sizes) == sum_before)
    return out


def run_extra(args, rt, lib, ranks, peak, timed):
    """C1, C3, C4, C5 of BASELINE.json: GB/s, elements/s and fraction of the HBM roofline."""
    from numpyllvm_b200 import abi
    from numpyllvm_b200.abi import Expr
    G, rank = ranks.world, ranks.rank
    out = []
    # the configs run 0.1-3 ms per step: a long warm-up keeps the allocation pause before each config
    # (clocks ramp down, the pool maps fresh pages) out of the 20 timed steps
    k, w = max(3, min(args.steps, 20)), 25
    sizes = rt.empty(np.int64, 4)
    lib.nl_pool_trim(0)                         # columns of the earlier arms go back to the driver: fresh, large chunks for these
    fused = rt.peer_ready()                     # NVLink mailboxes exchanged: reductions finish inside their kernel

    def report(name, dtype, n_per_gpu, bytes_per_gpu, ms, extra=None, k=k):
        gbs = G * bytes_per_gpu * k / (ms * 1e-3) / 1e9
        row = {"config": name, "dtype": dtype, "elements_per_gpu": n_per_gpu, "ms_per_step": round(ms / k, 5),
               "GB/s": round(gbs, 2), "elements_per_s": round(G * n_per_gpu * k / (ms * 1e-3), 1),
               "frac_of_measured_peak": round(gbs / (G * peak), 4), "frac_of_nominal_8000": round(gbs / (G * 8000.0), 4)}
        if extra:
            row.update(extra)
        out.append(row)

    # C1: (d*e*f)*(a*b*c), f64, 1e7 elements, 1 GPU; columns rotate through 3 buffer sets (1.7 GB)
    # so consecutive steps do not re-read L2-resident data
    if G == 1:
        n1 = 10_000_000
        sets = []
        for s in range(3):
            cols = [rt.empty(np.float64, n1) for _ in range(7)]
            for j in range(6):
                rt.fill(cols[j], 0, 100 + 10 * s + j, 2)
            sets.append(cols)
        e = Expr(np.float64)
        d_, e_, f_, a_, b_, c_ = (e.array() for _ in range(6))
        e.materialize(e.mul(e.mul(e.mul(d_, e_), f_), e.mul(e.mul(a_, b_), c_)))
        plan = abi.compile_plan(e)
        it = [0]

        def c1(_):
            cols = sets[it[0] % 3]; it[0] += 1
            rt.launch(plan, [cols[6].ptr], [x.ptr for x in cols[:6]], 0, n1, sizes.ptr)
        ms, _ = timed(c1, 30, 6)
        report("C1 Tests/multiplication.py (d*e*f)*(a*b*c)", "f64", n1, 56 * n1, ms,
               {"l2": "3 rotating column sets (1.68 GB) > 126 MB L2"}, k=30)
        for cols in sets:
            for x in cols:
                x.free()

    # C3: max over 2^31 f64 / int64 (sharded: 2^31 / G per GPU), NCCL allreduce of the partial
    n3 = (1 << args.log2n_reduce) // G
    for dt, nm, nl_dt in ((np.float64, "f64", abi.NL_F64), (np.int64, "i64", abi.NL_I64)):
        x = rt.empty(dt, n3)
        rt.fill(x, rank * n3, 7, 3 if dt == np.float64 else 1)
        res = rt.empty(dt, 1)
        for op, opname in ((abi.OP_MAX, "max"), (abi.OP_SUM, "sum")):
            e = Expr(dt)
            e.materialize(e.unop(op, e.array()))
            plan = abi.compile_plan(e)
            ptrs = abi.void_array([res.ptr])

            def c3(_):
                rt.launch(plan, [res.ptr], [x.ptr], 0, n3, sizes.ptr)
                abi.check(lib.nl_finish_reduce(op, nl_dt, ptrs, 0))

            def c3_fused(_):
                rt.launch_allreduce(plan, [res.ptr], [x.ptr], 0, n3, sizes.ptr)
            if G > 1 and fused:
                # the finish runs inside the reduction kernel over NVLink mailboxes; the NCCL finish is timed beside it
                ms, _ = timed(c3_fused, k, w)
                v_fused = rt.download(res)[0]
                ms_nccl, _ = timed(c3, k, w)
                same = bool(np.array_equal(v_fused, rt.download(res)[0])) if opname == "max" or nm == "i64" else None
                report(f"C3 Tests/max.py a.{opname}() + NVLink mailbox finish (in-kernel)", nm, n3, 8 * n3, ms,
                       {"ms_per_step_nccl_finish": round(ms_nccl / k, 5), "same_as_nccl": same})
            else:
                ms, _ = timed(c3, k, w)
                report(f"C3 Tests/max.py a.{opname}()" + (" + NCCL allreduce" if G > 1 else ""), nm, n3, 8 * n3, ms)
        x.free(); res.free()
        lib.nl_pool_trim(0)

    # C4: a[a > t] on 2^30 f64 at 1/10/50/90 % selectivity, counts scanned across shards
    n4 = (1 << args.log2n) // G
    x = rt.empty(np.float64, n4)
    rt.fill(x, rank * n4, 9, 3)
    y = rt.empty(np.float64, n4)
    offs = rt.empty(np.int64, G + 1)
    for sel in (0.01, 0.10, 0.50, 0.90):
        e = Expr(np.float64)
        xa = e.array()
        e.materialize(e.filter(xa, e.gt(e.ref(xa), e.scalar(1.0 - sel))))
        plan = abi.compile_plan(e)
        cptr, optr = abi.void_array([sizes.ptr]), abi.void_array([offs.ptr])

        def c4(_):
            rt.launch(plan, [y.ptr], [x.ptr, None], 0, n4, sizes.ptr)
            if G > 1:
                abi.check(lib.nl_finish_filter(cptr, optr, 0))
        def c4_fused(_):
            rt.launch_allgather(plan, [y.ptr], [x.ptr, None], 0, n4, sizes.ptr, offs.ptr)
        if G > 1 and fused:
            # the shard counts are exchanged and scanned inside the compaction kernel (NVLink mailboxes)
            ms, _ = timed(c4_fused, k, w)
            o_fused = rt.download(offs)
            ms_nccl, _ = timed(c4, k, w)
            cnt = int(rt.download(sizes, 1)[0])
            report(f"C4 Tests/filter.py a[a > t] sel={int(sel * 100)}% + NVLink mailbox count scan (in-kernel)", "f64", n4,
                   8 * n4 + 8 * cnt, ms, {"kept": cnt, "selectivity": round(cnt / n4, 5),
                                          "ms_per_step_nccl_finish": round(ms_nccl / k, 5),
                                          "same_as_nccl": bool(np.array_equal(o_fused, rt.download(offs)))})
        else:
            ms, _ = timed(c4, k, w)
            cnt = int(rt.download(sizes, 1)[0])
            extra = {"kept": cnt, "selectivity": round(cnt / n4, 5)}
            if G == 1:
                extra["properties_at_full_size"] = filter_properties(rt, abi, Expr, x, y, n4, cnt, 1.0 - sel, sizes)
            report(f"C4 Tests/filter.py a[a > t] sel={int(sel * 100)}%" + (" + count scan" if G > 1 else ""), "f64", n4,
                   8 * n4 + 8 * cnt, ms, extra)
    x.free(); y.free(); offs.free()
    lib.nl_pool_trim(0)

    # C5: (a[a >= t] * k).max() over 4e9 f32 on 8 GPUs = 5e8 per GPU, single pass, no compaction
    n5 = 500_000_000
    x = rt.empty(np.float32, n5)
    rt.fill(x, rank * n5, 11, 3)
    res = rt.empty(np.float32, 1)
    e = Expr(np.float32)
    xa = e.array()
    e.materialize(e.max(e.mul(e.filter(xa, e.ge(e.ref(xa), e.scalar(0.5))), e.scalar(3.0))))
    plan = abi.compile_plan(e)
    ptrs = abi.void_array([res.ptr])

    def c5(_):
        rt.launch(plan, [res.ptr], [x.ptr, None, None], 0, n5, sizes.ptr)
        abi.check(lib.nl_finish_reduce(abi.OP_MAX, abi.NL_F32, ptrs, 0))

    def c5_fused(_):
        rt.launch_allreduce(plan, [res.ptr], [x.ptr, None, None], 0, n5, sizes.ptr)
    if G > 1 and fused:
        ms, _ = timed(c5_fused, k, w)
        v_fused = float(rt.download(res)[0])
        ms_nccl, _ = timed(c5, k, w)
        report("C5 (a[a >= t] * k).max() fused + NVLink mailbox finish (in-kernel)", "f32", n5, 4 * n5, ms,
               {"result": v_fused, "ms_per_step_nccl_finish": round(ms_nccl / k, 5),
                "same_as_nccl": v_fused == float(rt.download(res)[0])})
    else:
        ms, _ = timed(c5, k, w)
        report("C5 (a[a >= t] * k).max() fused" + (" + NCCL allreduce" if G > 1 else ""), "f32", n5, 4 * n5, ms,
               {"result": float(rt.download(res)[0])})
    x.free(); res.free()

    # next row: the sort() full breaker as a device radix sort, one GPU's shard
    if G == 1:
        n6 = 1 << 28
        for dt, kind, nm in ((np.int64, 1, "i64 full range"), (np.float64, 3, "f64 in [0,1)"), (np.int64, 0, "i64 in [1,100) like Tests/")):
            x = rt.empty(dt, n6)
            tmp = rt.empty(dt, n6)
            best, passes, sum_before = None, 0, None
            for rep in range(3):
                rt.fill(x, 0, 5 + rep, kind)
                if np.dtype(dt).kind == "i":
                    sum_before = _reduce_device(rt, abi, Expr, abi.OP_SUM, dt, x.ptr, n6, sizes)
                rt.sync()
                l0 = lib.nl_launch_count()
                e0, e1 = rt.event(), rt.event()
                rt.record(e0)
                abi.check(lib.nl_sort(0, 0, abi.nl_dtype_of(dt), C.c_void_p(x.ptr), C.c_void_p(tmp.ptr), n6))
                rt.record(e1)
                rt.sync()
                ms1 = rt.elapsed_ms(e0, e1)
                best = ms1 if best is None else min(best, ms1)
                passes = (lib.nl_launch_count() - l0 - 1) // 3
            head = rt.download(x, 1 << 16)
            traffic = n6 * 8 * (1 + 3 * passes + (2 if passes % 2 else 0))     # histogram read + (count read, scatter read+write) per pass
            out.append({"config": f"sort() device radix sort, {nm}", "dtype": np.dtype(dt).name, "elements_per_gpu": n6,
                        "ms_per_step": round(best, 4), "Gkeys/s": round(n6 / best / 1e6, 2), "radix_passes": int(passes),
                        "GB/s": round(traffic / best / 1e6, 1), "frac_of_measured_peak": round(traffic / best / 1e6 / peak, 4),
                        "sorted": bool(np.all(head[:-1] <= head[1:])),
                        "properties_at_full_size": sort_properties(rt, abi, Expr, lib, x, n6, dt, sum_before, sizes)})
            x.free(); tmp.free()
    return out


# ------------------------------------------------------------------------------ CPU arms
def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def cpu_mul2(n, reps, threads):
    """The oracle port of the headline pipeline (c = a*b, f32) on `threads` host threads, chunked
    like the reference (scheduler.cpp:133-147 with thread_count = threads)."""
    import oracle
    a = oracle.fill(np.float32, 0, n, 42, 2)
    b = oracle.fill(np.float32, 0, n, 43, 2)
    c = np.empty_like(a)
    chunks = max(8, threads)
    oracle.fast_mul2_f32(a, b, c, chunks, threads)      # warm-up, page-faults the output
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        oracle.fast_mul2_f32(a, b, c, chunks, threads)
        a[0] = 1000.0                                    # the in-place update of the step
        times.append(time.perf_counter() - t0)
    return times


def cpu_baseline(args):
    threads = host_threads()
    n = 1 << min(args.log2n, 27)
    times = cpu_mul2(n, 8, threads)
    best = statistics.median(times)
    return {"value": round(12 * n / best / 1e9, 2), "unit": "GB/s", "cores": threads, "kind": "port",
            "sample": f"c = a*b on 2^{n.bit_length() - 1} f32 elements x 8 repetitions (median), gcc -O3 compiled loop, "
                      f"{max(8, threads)} chunks on {threads} pthreads",
            "elements_per_s": round(n / best, 1)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference
    cannot be built here (CPython 2 + LLVM 3.9), so this is the oracle port with all host
    threads; rank 0 alone runs it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    n = 1 << min(args.log2n, 27)
    K, W = args.steps, max(args.warmup, 1)
    K = min(K, 50)
    times = cpu_mul2(n, K + W, threads)[W:]
    total = sum(times)
    value = 12 * n * len(times) / total / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": "GB/s", "n_gpus": args.gpus,
            "steps": len(times), "warmup": W, "ms_per_step": round(1e3 * total / len(times), 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Tests/assignment.py at scale: c = a*b forced by a[0] = 1000, then the in-place update",
                       "elements_per_step": n, "note": "bounded sample of the 2^30-element workload"},
            "cpu_baseline": {"value": round(value, 2), "unit": "GB/s", "cores": threads, "kind": "port",
                             "sample": f"2^{n.bit_length() - 1} f32 elements per step, {max(8, threads)} chunks on {threads} pthreads"},
            "e2e": {"value": round(value, 2), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "elements_per_s": round(n * len(times) / total, 1)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=30, help="log2 of elements per GPU for C2/C4")
    ap.add_argument("--log2n-reduce", dest="log2n_reduce", type=int, default=31, help="log2 of total elements for C3")
    ap.add_argument("--e2e-steps", dest="e2e_steps", type=int, default=3)
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
