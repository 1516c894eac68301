#!/usr/bin/env python
"""bench.py — effective HBM GB/s (and elements/s) of numpyllvm's fused thunk pipelines on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

Headline workload (BASELINE.json configs[1]): Tests/assignment.py at scale —
`c = a * b` pending on 2^30-element f32 columns, `a[0] = 1000` forces it
(thunk_as_mapping.cpp:80-86), one step = the forced fused pipeline (12 B/elem)
plus the in-place update.  Weak scaling: every GPU owns a 2^30-element shard
and there is no data-path collective.  `e2e` is the same step through the `jit`
extension with HOST buffers (upload, kernel and download overlapped chunk by
chunk).  The other BASELINE.json configs are reported in `pipelines` (C1, C3, C4,
C5, sort), each with a full-size parity verdict against the oracle and, at
N > 1, its efficiency against the same config run on one GPU in the same
invocation.  One JSON line is printed by rank 0; its last key is a compact
summary of every pipeline row.
"""
from __future__ import annotations

import os

# NCCL's own log stays on (the driver reads the ranks of the communicator from it); set before anything that might
# load or initialise NCCL is imported
os.environ.setdefault("NCCL_DEBUG", "INFO")
os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")

import argparse
import ctypes as C
import json
import statistics
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "effective HBM GB/s per fused pipeline"
FALLBACK_PEAK_GBS = 6650.0     # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
M64 = (1 << 64) - 1


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy_)"
    except Exception:
        return FALLBACK_PEAK_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def headline_config(n, G):
    """`config` of the JSON line — the SAME dict in both arms (the reference arm times the same step)."""
    return {"workload": "Tests/assignment.py at scale: c = a*b forced by a[0] = 1000, then the in-place update",
            "elements_per_gpu": n, "algorithmic_bytes_per_step_per_gpu": 12 * n + 4,
            "l2": "inputs (12 GiB per GPU) are larger than L2; no flush needed",
            "parallelism": f"{G} contiguous shards, one rank per GPU, no data-path collective"}


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
            self.proc = None

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


# ------------------------------------------------------------------------------ host side of the parity checks
def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def words_checksum(buf: np.ndarray):
    """(n_words, sum, xor, position-weighted sum) of an array's bytes read as little-endian u64 words, a
    trailing partial word zero-padded — the host twin of nl_checksum() (nl_abi.h)."""
    raw = np.ascontiguousarray(buf).view(np.uint8).reshape(-1)
    pad = (-raw.size) % 8
    if pad:
        raw = np.concatenate([raw, np.zeros(pad, np.uint8)])
    w = raw.view(np.uint64)
    if w.size == 0:
        return 0, 0, 0, 0
    k = np.arange(w.size, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return int(w.size), int(np.add.reduce(w)), int(np.bitwise_xor.reduce(w)), int(np.add.reduce(w * (2 * k + 1)))


def combine_checksums(parts):
    """Fold per-chunk (n_words, sum, xor, weighted) in order: a chunk starting at word b contributes
    weighted + 2 b sum (see nl_checksum.cu)."""
    base = s = x = p = 0
    for nw, cs, cx, cp in parts:
        s = (s + cs) & M64
        x ^= cx
        p = (p + cp + 2 * base * cs) & M64
        base += nw
    return base, s, x, p


def oracle_stream(n, work, threads, chunk=1 << 22):
    """work(first, count) over [0, n) in chunks on `threads` host threads (ctypes releases the GIL);
    the partial results in chunk order."""
    spans = [(s, min(chunk, n - s)) for s in range(0, n, chunk)]
    if threads <= 1:
        return [work(s, m) for s, m in spans]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        return list(ex.map(lambda sm: work(*sm), spans))


def device_checksum(rt, lib, ptr, nbytes, scratch):
    from numpyllvm_b200 import abi
    abi.check(lib.nl_checksum(0, 0, C.c_void_p(ptr), nbytes, C.c_void_p(scratch.ptr)))
    v = rt.download(scratch, 3)
    return int(v[0]), int(v[1]), int(v[2])


def verdict(ok, what):
    return ("bit-exact vs oracle: " if ok else "MISMATCH vs oracle: ") + what


# ------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    from numpyllvm_b200 import abi
    from numpyllvm_b200.abi import Expr
    from numpyllvm_b200.device import Runtime
    from numpyllvm_b200.dist import Ranks

    # NCCL's own log stays on (the driver reads the ranks of the communicator from it).  NCCL prints to
    # the process's stdout, and the JSON line is the only thing this arm may print there: file descriptor 1
    # is pointed at stderr for everything native, the line itself goes out through a private copy.
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ranks = Ranks(args.gpus)
    lib = abi.load_library()
    dev_ids = (C.c_int * 1)(ranks.local_rank)
    abi.check(lib.nl_init(1, dev_ids))          # one rank drives one GPU
    rt = Runtime.get(1)
    ranks.init_library_comm(lib)                # NCCL communicator for the data-path collectives
    if ranks.world > 1:
        # this repo's own note beside NCCL's INFO lines: what the library's communicator looks like from this rank
        print(f"[numpyllvm_b200] rank {ranks.rank}: libnlkernels communicator ready, nranks {lib.nl_comm_world_size()}, "
              f"cudaDev {ranks.local_rank}, NVLink mailboxes {'mapped (fused in-kernel finish)' if lib.nl_comm_peer_ready() else 'not available (NCCL finish)'}",
              file=sys.stderr, flush=True)
    peak, peak_src = measured_peak()
    G = ranks.world
    n = 1 << args.log2n           # elements per GPU (weak scaling)
    K, W = args.steps, max(args.warmup, 3)
    threads = max(1, host_threads() // G)       # host threads of this rank for the oracle side of the checks

    def timed(step_fn, k, w, per_launch=None, solo=False):
        """w warm-up + k timed steps, barrier + device sync on both sides, CUDA events on the
        launching stream, max over ranks (solo: rank 0 alone, no barrier).  Returns (ms total, [per-launch ms])."""
        for _ in range(w):
            step_fn(None)
        rt.sync()
        if not solo:
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
        if solo:
            return ms, launch_ms
        ranks.barrier()
        return ranks.max(ms), launch_ms

    # ---------------- headline: C2 assignment (f32, 2^30 per GPU) ----------------
    f32 = np.float32
    a, b, c = rt.empty(f32, n), rt.empty(f32, n), rt.empty(f32, n)
    rt.fill(a, ranks.rank * n, 42, 2)
    rt.fill(b, ranks.rank * n, 43, 2)
    sizes = rt.empty(np.int64, 4)
    csum = rt.empty(np.uint64, 4)
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

    # full-size parity (outside the timed region): 64-bit checksums of the whole result column against
    # the oracle's compiled loop run chunk-streamed on regenerated inputs — every rank checks its shard
    parity = "unchecked"
    if not args.no_check:
        import oracle
        got = device_checksum(rt, lib, c.ptr, 4 * n, csum)
        first = ranks.rank * n

        def c2_chunk(s, m):
            wa = oracle.fill(f32, first + s, m, 42, 2)
            wb = oracle.fill(f32, first + s, m, 43, 2)
            if s == 0 and ranks.rank == 0 and K + W > 1:
                wa[0] = 1000.0            # every step after the first sees the updated a[0]
            return words_checksum(oracle.fast_mul2_f32(wa, wb, np.empty_like(wa), 8, 1))
        want = combine_checksums(oracle_stream(n, c2_chunk, threads))
        ok = got == want[1:] and want[0] == n // 2
        ok &= (ranks.rank != 0) or float(rt.download(a, 1)[0]) == 1000.0
        ok = ranks.max(0.0 if ok else 1.0) == 0.0
        parity = verdict(ok, f"sum/xor/position-weighted 64-bit checksums of all {n} elements of c on every rank")

    # ---------------- e2e: same step with HOST buffers, copies inside the timed region ------------
    e2e = e2e_streamed = None
    if not args.no_e2e:
        if args.cabi_e2e:
            e2e_streamed = run_e2e_cabi(args, rt, lib, ranks, plan, a, b, c, sizes, n, thousand)
        a.free(); b.free(); c.free()
        e2e = run_e2e_jit(args, ranks, n)

    result = {
        "metric": METRIC, "value": round(value, 2), "unit": "GB/s", "n_gpus": G, "steps": K, "warmup": W,
        "ms_per_step": round(ms_total / K, 5), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": headline_config(n, G),
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
    if e2e_streamed:
        result["e2e_streamed_cabi"] = e2e_streamed
    a.free(); b.free(); c.free(); csum.free()

    # ---------------- CPU baseline beside it (rank 0, N = 1 only) ----------------
    if ranks.rank == 0 and G == 1 and not args.no_cpu:
        result["cpu_baseline"] = cpu_baseline(args)

    # ---------------- the other BASELINE.json configs ----------------
    if not args.no_extra:
        try:
            rows = run_extra(args, rt, lib, ranks, peak, timed, threads)
            result["pipelines"] = rows
            # last key of the line: one short entry per row, so that a tail of the output still shows every config
            result["pipelines_summary"] = {r["id"]: [r.get("GB/s", r.get("Gkeys/s")), r.get("frac_of_measured_peak"),
                                                     r.get("efficiency_vs_n1"), "ok" if str(r.get("parity", "")).startswith("bit-exact") or
                                                     str(r.get("parity", "")).startswith("within") else r.get("parity", "unchecked")]
                                           for r in rows}
            result["pipelines_summary"]["_columns"] = ["GB/s (Gkeys/s for sort)", "frac_of_measured_peak", "efficiency_vs_n1", "parity"]
        except Exception as exc:                    # the extra configs must never cost the headline line
            result["pipelines_error"] = f"{type(exc).__name__}: {exc}"
    sampler.stop()
    if ranks.rank == 0:
        print(json.dumps(result), file=json_out, flush=True)
    ranks.finish()
    try:
        C.CDLL(None).fflush(None)               # NCCL logs through C stdio, which is fully buffered on a pipe
    except Exception:
        pass


def load_traffic(key):
    """dram bytes read+written per launch from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            return json.load(fh).get(key)
    except Exception:
        return None


def run_e2e_cabi(args, rt, lib, ranks, plan, a, b, c, sizes, n, thousand):
    """Extra data point (--cabi-e2e): the same step through the bare C-ABI with HOST (pinned) buffers: H2D of
    both inputs, the fused pipeline, the in-place update and D2H of the result, chunked so the upload,
    the kernel and the download overlap on the three streams of the device."""
    from numpyllvm_b200 import abi
    nbytes = 4 * n
    hp = [C.c_void_p() for _ in range(3)]
    for p in hp:
        abi.check(lib.nl_host_alloc(nbytes, C.byref(p)))
    ha = np.ctypeslib.as_array(C.cast(hp[0], C.POINTER(C.c_float)), shape=(n,))
    hb = np.ctypeslib.as_array(C.cast(hp[1], C.POINTER(C.c_float)), shape=(n,))
    hc = np.ctypeslib.as_array(C.cast(hp[2], C.POINTER(C.c_float)), shape=(n,))
    rt.fill(a, ranks.rank * n, 42, 2)
    abi.check(lib.nl_memcpy_d2h(0, 0, hp[0], C.c_void_p(a.ptr), nbytes))
    abi.check(lib.nl_memcpy_d2h(0, 0, hp[1], C.c_void_p(b.ptr), nbytes))
    rt.sync()
    chunk = min(n, 1 << 24)          # 64 MiB per column per chunk
    nchunks = (n + chunk - 1) // chunk

    def step():
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
    t0 = time.perf_counter()
    for _ in range(k):
        step()
    rt.sync()
    ms = ranks.max((time.perf_counter() - t0) * 1e3)
    ok = bool(np.array_equal(hc[:4096], ha[:4096] * hb[:4096])) and bool(np.array_equal(hc[-4096:], ha[-4096:] * hb[-4096:]))
    for p in hp:
        lib.nl_host_free(p)
    G = ranks.world
    return {"value": round(G * (12 * n + 4) * k / (ms * 1e-3) / 1e9, 2), "unit": "GB/s",
            "h2d_bytes_per_step": 2 * nbytes, "d2h_bytes_per_step": nbytes, "steps": k,
            "ms_per_step": round(ms / k, 3), "api": "C-ABI nl_memcpy_h2d / nl_launch_pipeline / nl_memcpy_d2h, pinned host buffers, 64 MiB chunks on 3 streams",
            "parity": "bit-exact on sampled windows" if ok else "MISMATCH"}


def mem_available_bytes():
    try:
        with open("/proc/meminfo") as fh:
            for line in fh:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) * 1024
    except Exception:
        pass
    return None


def run_e2e_jit(args, ranks, n):
    """The headline end-to-end number: the step exactly as a user of the reference writes it
    (Tests/assignment.py), through the `jit` extension, with HOST buffers.  Every step hands both
    columns to jit.thunk() from page-locked host memory, forces the pipeline with the in-place update
    and reads the result back into host memory (asnumpyarray).  jit.thunk() returns while PCIe is
    busy; the kernel follows the upload chunk by chunk and each chunk of the result leaves as soon as
    it is computed, so the three legs overlap (storage.cpp, scheduler.cpp)."""
    import numpyllvm_b200
    jit = numpyllvm_b200.install_jit_alias()
    import oracle
    G = ranks.world
    # the same workload at every N: 2^30 elements per GPU (12 GiB of page-locked host memory per rank)
    n_e, note = n, None
    avail = mem_available_bytes()
    while avail is not None and n_e > (1 << 20) and 12 * n_e * G * 1.3 > avail:
        n_e //= 2
        note = f"host memory ({avail >> 30} GiB available) bounds the page-locked buffers: {n_e} elements per GPU"
    ha, hb = jit.pinned_empty(n_e, np.float32), jit.pinned_empty(n_e, np.float32)
    threads = max(1, host_threads() // G)
    first = ranks.rank * n

    def gen(s, m):                                         # synthetic columns, generated on the host
        ha[s:s + m] = oracle.fill(np.float32, first + s, m, 42, 2)
        hb[s:s + m] = oracle.fill(np.float32, first + s, m, 43, 2)
        return 0
    oracle_stream(n_e, gen, threads)

    def step():
        a = jit.thunk(ha)
        b = jit.thunk(hb)
        c = a * b
        a[0] = 1000
        c.evaluate()
        return c.asnumpyarray()

    out = step()                                           # warm-up: pins the result pool, loads kernels
    # the whole host result against the oracle's loop on the host columns (the upload is part of what is checked)
    def chk(s, m):
        return bool(np.array_equal(out[s:s + m], oracle.fast_mul2_f32(ha[s:s + m], hb[s:s + m], np.empty(m, np.float32), 8, 1)))
    ok = all(oracle_stream(n_e, chk, threads))
    del out
    out = step(); del out                                  # second warm-up: every pool block exists now
    k = max(1, min(args.steps, args.e2e_steps))
    jit.synchronize()
    ranks.barrier()
    t0 = time.perf_counter()
    for _ in range(k):
        out = step()
        del out
    jit.synchronize()
    mine = time.perf_counter() - t0
    dt = ranks.max(mine)
    # every rank's own time: on a box whose GPUs sit behind two sockets the ranks far from the host buffers are the slow ones
    by_rank = [round(1e3 * ranks.max(mine if ranks.rank == r else 0.0) / k, 1) for r in range(G)] if G > 1 else None
    ranks.barrier()
    ok = ranks.max(0.0 if ok else 1.0) == 0.0
    res = {"value": round(G * (12 * n_e + 4) * k / dt / 1e9, 2), "unit": "GB/s",
           "h2d_bytes_per_step": 8 * n_e, "d2h_bytes_per_step": 4 * n_e, "steps": k,
           "ms_per_step": round(1e3 * dt / k, 3), "elements_per_gpu": n_e,
           "api": "jit.thunk(pinned ndarray) x2 -> c = a*b -> a[0] = 1000 -> c.evaluate() -> c.asnumpyarray(); wall clock, max over ranks",
           "overlap": "upload (copy stream), kernel (compute stream) and download (copy-back stream) pipelined in 32 MiB chunks",
           "bound": "PCIe: 8 bytes up and 4 bytes down per element for 12 algorithmic bytes",
           "parity": verdict(ok, f"all {n_e} elements of the host result on every rank")}
    if by_rank:
        res["ms_per_step_by_rank"] = by_rank
    if note:
        res["note"] = note
    return res


def _reduce_device(rt, abi, Expr, op, dtype, ptr, n, sizes):
    """op over a device column through the reduction pipeline (used for full-size property checks)."""
    res = rt.empty(dtype, 1)
    e = Expr(dtype)
    e.materialize(e.unop(op, e.array()))
    rt.launch(abi.compile_plan(e), [res.ptr], [ptr], 0, n, sizes.ptr)
    v = rt.download(res)[0]
    res.free()
    return v


def _sum_of_squares(rt, abi, Expr, dtype, ptr, n, sizes):
    res = rt.empty(dtype, 1)
    e = Expr(dtype)
    a = e.array()
    e.materialize(e.sum(e.mul(a, e.ref(a))))
    rt.launch(abi.compile_plan(e), [res.ptr], [ptr], 0, n, sizes.ptr)
    v = rt.download(res)[0]
    res.free()
    return v


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


def run_extra(args, rt, lib, ranks, peak, timed, threads):
    """C1, C3, C4, C5 of BASELINE.json (+ the device sort): GB/s, elements/s, fraction of the HBM roofline, a
    full-size parity verdict against the oracle and, at N > 1, the efficiency against one GPU."""
    from numpyllvm_b200 import abi
    from numpyllvm_b200.abi import Expr
    import oracle
    G, rank = ranks.world, ranks.rank
    out = []
    check = not args.no_check
    # the configs run 0.1-3 ms per step: a long warm-up keeps the allocation pause before each config
    # (clocks ramp down, the pool maps fresh pages) out of the 20 timed steps
    k, w = max(3, min(args.steps, 20)), 25
    sizes = rt.empty(np.int64, 4)
    csum = rt.empty(np.uint64, 4)
    lib.nl_pool_trim(0)                         # columns of the earlier arms go back to the driver: fresh, large chunks for these
    fused = rt.peer_ready()                     # NVLink mailboxes exchanged: reductions finish inside their kernel

    def all_ok(ok):
        return ranks.max(0.0 if ok else 1.0) == 0.0

    def report(rid, name, dtype, n_per_gpu, bytes_per_gpu, ms, extra=None, k=k, solo_ms=None, solo_bytes=None):
        gbs = G * bytes_per_gpu * k / (ms * 1e-3) / 1e9
        row = {"id": rid, "config": name, "dtype": dtype, "elements_per_gpu": n_per_gpu, "ms_per_step": round(ms / k, 5),
               "GB/s": round(gbs, 2), "elements_per_s": round(G * n_per_gpu * k / (ms * 1e-3), 1),
               "frac_of_measured_peak": round(gbs / (G * peak), 4), "frac_of_nominal_8000": round(gbs / (G * 8000.0), 4)}
        if solo_ms is not None:
            # the same config on ONE GPU (rank 0 alone, same invocation): strong scaling unless noted
            solo_gbs = solo_bytes * k / (solo_ms * 1e-3) / 1e9
            row["GB/s_on_1_gpu"] = round(solo_gbs, 2)
            row["efficiency_vs_n1"] = round(gbs / (G * solo_gbs), 4)
        if extra:
            row.update(extra)
        out.append(row)

    def solo(fn_factory):
        """rank 0 runs the N = 1 form of a config alone while the others wait; returns ms over k steps or None"""
        ms = None
        if G > 1:
            if rank == 0:
                ms = fn_factory()
            ranks.barrier()
            ms = ranks.max(ms if ms is not None else 0.0)
        return ms

    # C1: (d*e*f)*(a*b*c), f64, 1e7 elements, 1 GPU; columns rotate through 3 buffer sets (1.7 GB)
    # so consecutive steps do not re-read L2-resident data
    if G == 1:
        n1 = args.n1
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
        extra = {"l2": "3 rotating column sets (1.68 GB) > 126 MB L2"}
        wants = []
        if check:
            ok = True
            for s in range(3):
                got = device_checksum(rt, lib, sets[s][6].ptr, 8 * n1, csum)

                def c1_chunk(s0, m, s=s):
                    cols = [oracle.fill(np.float64, s0, m, 100 + 10 * s + j, 2) for j in range(6)]
                    # nlo_fast_mul6_f64: ((in0*in1)*in2) * ((in3*in4)*in5), the association of the plan above
                    return words_checksum(oracle.fast_mul6_f64(cols, np.empty(m), 8, 1))
                wants.append(combine_checksums(oracle_stream(n1, c1_chunk, threads))[1:])
                ok &= got == wants[-1]
            extra["parity"] = verdict(ok, f"64-bit checksums of all {n1} elements of all 3 result columns")
        report("C1", "C1 Tests/multiplication.py (d*e*f)*(a*b*c)", "f64", n1, 56 * n1, ms, extra, k=30)
        # the same launches replayed from a CUDA graph (one graph = the three column sets in turn)
        try:
            abi.check(lib.nl_graph_begin(0))
            for _ in range(3):
                c1(None)
            graph = C.c_void_p()
            abi.check(lib.nl_graph_end(0, C.byref(graph)))
            for s in range(3):                                  # the replays must produce the columns again
                abi.check(lib.nl_memset(0, 0, C.c_void_p(sets[s][6].ptr), 0, 8 * n1))
            ms, _ = timed(lambda _: abi.check(lib.nl_graph_launch(graph)), 10, 3)
            gextra = {"launches_per_step": 3}
            if check and wants:
                okg = all(device_checksum(rt, lib, sets[s][6].ptr, 8 * n1, csum) == wants[s] for s in range(3))
                gextra["parity"] = verdict(okg, f"the replays rebuilt all 3 zeroed result columns: 64-bit checksums of all {n1} elements each")
            report("C1 graph", "C1 replayed from a CUDA graph (nl_graph_*: 3 launches per replay)", "f64", n1, 56 * n1 * 3, ms, gextra, k=10)
            lib.nl_graph_destroy(graph)
        except Exception as exc:
            out.append({"id": "C1 graph", "config": "C1 replayed from a CUDA graph", "error": f"{type(exc).__name__}: {exc}"})
        for cols in sets:
            for x in cols:
                x.free()
        # ... and through the `jit` API with device-resident thunks: what one small pipeline costs end to end on the
        # host side (parse, lower, plan, allocate, launch); evaluation is stream-ordered, so launches overlap kernels
        try:
            import numpyllvm_b200
            jit = numpyllvm_b200.install_jit_alias()
            hosts = [oracle.fill(np.float64, 0, n1, 100 + j, 2) for j in range(6)]
            ts = [jit.thunk(h) for h in hosts]
            jit.synchronize()

            def jit_step():
                r = (ts[0] * ts[1] * ts[2]) * (ts[3] * ts[4] * ts[5])
                r.evaluate()
                return r
            keep = [jit_step() for _ in range(3)]
            jit.synchronize()
            t0 = time.perf_counter()
            r = jit_step()
            jit.synchronize()
            one = time.perf_counter() - t0                       # latency of ONE pipeline, launch to completion
            t0 = time.perf_counter()
            for _ in range(30):
                keep[0] = jit_step()
            t_host = time.perf_counter() - t0                    # host time to enqueue 30 pipelines
            jit.synchronize()
            t_all = time.perf_counter() - t0
            okj = bool(np.array_equal(r.asnumpyarray()[:4096], ((hosts[0] * hosts[1]) * hosts[2] * ((hosts[3] * hosts[4]) * hosts[5]))[:4096]))
            out.append({"id": "C1 jit", "config": "C1 through the jit API, device-resident thunks: r = (d*e*f)*(a*b*c); r.evaluate()", "dtype": "f64",
                        "elements_per_gpu": n1, "latency_ms_one_pipeline": round(one * 1e3, 4), "ms_per_step": round(t_all / 30 * 1e3, 5),
                        "host_enqueue_ms_per_step": round(t_host / 30 * 1e3, 5), "GB/s": round(56 * n1 * 30 / t_all / 1e9, 2),
                        "frac_of_measured_peak": round(56 * n1 * 30 / t_all / 1e9 / peak, 4),
                        "parity": verdict(okj, "first 4096 elements vs NumPy (same association)")})
            del ts, keep, r
            jit.synchronize()
        except Exception as exc:
            out.append({"id": "C1 jit", "config": "C1 through the jit API", "error": f"{type(exc).__name__}: {exc}"})

    # C3: max / sum over 2^31 f64 / int64 (sharded: 2^31 / G per GPU), partials exchanged over NVLink
    n3_total = 1 << args.log2n_reduce
    n3 = n3_total // G
    for dt, nm, nl_dt in ((np.float64, "f64", abi.NL_F64), (np.int64, "i64", abi.NL_I64)):
        kind = 3 if dt == np.float64 else 1
        want = {}
        if check and rank == 0:
            # the oracle's answer for the WHOLE column (every rank's shard), chunk-streamed on rank 0's host threads
            def c3_chunk(s0, m):
                xs = oracle.fill(dt, s0, m, 7, kind)
                if dt == np.float64:
                    return oracle.fast_max_f64(xs, 8, 1), float(np.add.reduce(xs))
                return oracle.fast_max_i64(xs, 8, 1), oracle.fast_sum_i64(xs, 8, 1)
            parts = oracle_stream(n3_total, c3_chunk, host_threads())
            want["max"] = max(p[0] for p in parts)
            want["sum"] = (float(np.add.reduce(np.array([p[1] for p in parts]))) if dt == np.float64
                           else int(np.add.reduce(np.array([p[1] for p in parts], dtype=np.int64))))
        solo_ms = {}
        if G > 1 and rank == 0:
            xs_ = rt.empty(dt, n3_total)
            rt.fill(xs_, 0, 7, kind)
            rs_ = rt.empty(dt, 1)
            for op, opname in ((abi.OP_MAX, "max"), (abi.OP_SUM, "sum")):
                e = Expr(dt)
                e.materialize(e.unop(op, e.array()))
                pl = abi.compile_plan(e)
                solo_ms[opname], _ = timed(lambda _: rt.launch(pl, [rs_.ptr], [xs_.ptr], 0, n3_total, sizes.ptr), k, w, solo=True)
            xs_.free(); rs_.free()
            lib.nl_pool_trim(0)
        ranks.barrier()
        x = rt.empty(dt, n3)
        rt.fill(x, rank * n3, 7, kind)
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

            def c3_local(_):
                rt.launch(plan, [res.ptr], [x.ptr], 0, n3, sizes.ptr)

            def par(v):
                if not check:
                    return {}
                ok, fsum = True, opname == "sum" and dt == np.float64
                if rank == 0:                    # rank 0 holds the oracle's answer; every rank holds the same device result
                    if fsum:
                        ok = abs(float(v) - want["sum"]) <= 1e-12 * abs(want["sum"])
                    else:
                        ok = (float(v) == want[opname]) if dt == np.float64 else (int(v) == want[opname])
                ok = all_ok(ok)
                if fsum:
                    return {"parity": ("within rtol 1e-12 of the oracle: " if ok else "MISMATCH vs oracle: ") +
                            f"sum over all {n3_total} elements = {float(v)!r}" + (f" (oracle {want['sum']!r})" if rank == 0 else "")}
                return {"parity": verdict(ok, f"{opname} over all {n3_total} elements = {v!r}")}
            sm = ranks.max(solo_ms.get(opname, 0.0)) if G > 1 else None
            if G > 1 and fused:
                # the finish runs inside the reduction kernel over NVLink mailboxes; the NCCL finish is timed beside it
                ms, _ = timed(c3_fused, k, w)
                v_fused = rt.download(res)[0]
                ms_nccl, _ = timed(c3, k, w)
                v_nccl = rt.download(res)[0]
                ms_local, _ = timed(c3_local, k, w)
                same = bool(np.array_equal(v_fused, v_nccl)) if opname == "max" or nm == "i64" else bool(abs(v_fused - v_nccl) <= 1e-12 * abs(v_nccl))
                extra = {"ms_per_step_nccl_finish": round(ms_nccl / k, 5), "ms_per_step_without_finish": round(ms_local / k, 5),
                         "same_as_nccl": same}
                extra.update(par(v_fused))
                report(f"C3 {nm} {opname}", f"C3 Tests/max.py a.{opname}() + NVLink mailbox finish (in-kernel)", nm, n3, 8 * n3, ms,
                       extra, solo_ms=sm, solo_bytes=8 * n3_total)
            else:
                ms, _ = timed(c3, k, w)
                extra = par(rt.download(res)[0])
                report(f"C3 {nm} {opname}", f"C3 Tests/max.py a.{opname}()" + (" + NCCL allreduce" if G > 1 else ""), nm, n3, 8 * n3, ms,
                       extra, solo_ms=sm if G > 1 else None, solo_bytes=8 * n3_total)
        x.free(); res.free()
        lib.nl_pool_trim(0)

    # C4: a[a > t] on 2^30 f64 at 1/10/50/90 % selectivity, counts scanned across shards
    n4_total = 1 << args.log2n
    n4 = n4_total // G
    sels = (0.01, 0.10, 0.50, 0.90)
    want4 = None
    if check:
        # every rank: the oracle's compaction of ITS shard, all four thresholds in one pass over the data
        def c4_chunk(s0, m):
            xs = oracle.fill(np.float64, rank * n4 + s0, m, 9, 3)
            buf = np.empty(m)
            return [words_checksum(oracle.fast_filter_gt_f64(xs, 1.0 - sel, buf, 8, 1)) for sel in sels]
        parts = oracle_stream(n4, c4_chunk, threads)
        want4 = [combine_checksums([p[j] for p in parts]) for j in range(len(sels))]
    solo4 = {}
    if G > 1 and rank == 0:
        xs_, ys_ = rt.empty(np.float64, n4_total), rt.empty(np.float64, n4_total)
        rt.fill(xs_, 0, 9, 3)
        for sel in sels:
            e = Expr(np.float64)
            xa = e.array()
            e.materialize(e.filter(xa, e.gt(e.ref(xa), e.scalar(1.0 - sel))))
            pl = abi.compile_plan(e)
            ms1, _ = timed(lambda _: rt.launch(pl, [ys_.ptr], [xs_.ptr, None], 0, n4_total, sizes.ptr), k, w, solo=True)
            solo4[sel] = (ms1, 8 * n4_total + 8 * int(rt.download(sizes, 1)[0]))
        xs_.free(); ys_.free()
        lib.nl_pool_trim(0)
    ranks.barrier()
    x = rt.empty(np.float64, n4)
    rt.fill(x, rank * n4, 9, 3)
    y = rt.empty(np.float64, n4)
    offs = rt.empty(np.int64, G + 1)
    for j, sel in enumerate(sels):
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

        def c4_local(_):
            rt.launch(plan, [y.ptr], [x.ptr, None], 0, n4, sizes.ptr)

        def par4(cnt, offsets):
            if not check:
                return {}
            got = device_checksum(rt, lib, y.ptr, 8 * cnt, csum)
            ok = cnt == want4[j][0] and got == want4[j][1:]
            what = f"count, values and ORDER of the {cnt} survivors of this rank's shard (64-bit checksums incl. position-weighted)"
            if offsets is not None:
                counts = ranks.all_gather_int(want4[j][0])
                ex = [0]
                for cval in counts:
                    ex.append(ex[-1] + cval)
                ok &= [int(v) for v in offsets] == ex
                what += f"; offsets vector {ex} on every rank"
            return {"parity": verdict(all_ok(ok), what)}
        kname4 = None
        if G > 1 and fused:
            # the shard counts are exchanged and scanned inside the compaction kernel (NVLink mailboxes)
            ms, _ = timed(c4_fused, k, w)
            o_fused = rt.download(offs)
            cnt = int(rt.download(sizes, 1)[0])
            extra = par4(cnt, o_fused)
            ms_nccl, _ = timed(c4, k, w)
            o_nccl = rt.download(offs)
            ms_local, _ = timed(c4_local, k, w)
            total_kept = int(o_fused[G])
            s1 = solo4.get(sel)
            sm, sb = (ranks.max(s1[0] if s1 else 0.0), int(ranks.max(float(s1[1]) if s1 else 0.0)))
            extra.update({"kept": cnt, "kept_total": total_kept, "selectivity": round(total_kept / n4_total, 5),
                          "ms_per_step_nccl_finish": round(ms_nccl / k, 5), "ms_per_step_without_finish": round(ms_local / k, 5),
                          "same_as_nccl": bool(np.array_equal(o_fused, o_nccl))})
            # bytes: every rank reads its shard and writes its survivors; aggregate = 8 n_total + 8 kept_total
            gb_per_gpu = (8 * n4_total + 8 * total_kept) / G
            report(f"C4 {int(sel * 100)}%", f"C4 Tests/filter.py a[a > t] sel={int(sel * 100)}% + NVLink mailbox count scan (in-kernel)", "f64", n4,
                   gb_per_gpu, ms, extra, solo_ms=sm, solo_bytes=sb)
        else:
            ms, _ = timed(c4, k, w)
            kname4 = rt.last_launch()[0]
            cnt = int(rt.download(sizes, 1)[0])
            extra = {"kept": cnt, "selectivity": round(cnt / n4, 5), "kernel": kname4}
            extra.update(par4(cnt, rt.download(offs) if G > 1 else None))
            report(f"C4 {int(sel * 100)}%", f"C4 Tests/filter.py a[a > t] sel={int(sel * 100)}%" + (" + count scan" if G > 1 else ""), "f64", n4,
                   8 * n4 + 8 * cnt, ms, extra)
    # a filter that has no build-time shape: (a*2+1)[a > t] runs the super-tile kernel under the step interpreter
    if G == 1:
        for sel in (0.10, 0.50):
            e = Expr(np.float64)
            xa = e.array()
            val = e.add(e.mul(xa, e.scalar(2.0)), e.scalar(1.0))
            e.materialize(e.filter(val, e.gt(e.ref(xa), e.scalar(1.0 - sel))))
            plan = abi.compile_plan(e)
            ms, _ = timed(lambda _: rt.launch(plan, [y.ptr], [x.ptr, None, None, None], 0, n4, sizes.ptr), k, w)
            cnt = int(rt.download(sizes, 1)[0])
            extra = {"kept": cnt, "kernel": rt.last_launch()[0]}
            if check:
                def c4d_chunk(s0, m, thr=1.0 - sel):
                    xs = oracle.fill(np.float64, s0, m, 9, 3)
                    return words_checksum(xs[xs > thr] * 2.0 + 1.0)
                wantd = combine_checksums(oracle_stream(n4, c4d_chunk, threads))
                got = device_checksum(rt, lib, y.ptr, 8 * cnt, csum)
                extra["parity"] = verdict(cnt == wantd[0] and got == wantd[1:], f"count, values and order of the {cnt} survivors")
            report(f"C4x {int(sel * 100)}%", f"filter off the static table: (a*2+1)[a > t] sel={int(sel * 100)}% (interpreter-driven compaction)", "f64",
                   n4, 8 * n4 + 8 * cnt, ms, extra)
    x.free(); y.free(); offs.free()
    lib.nl_pool_trim(0)

    # C5: (a[a >= t] * k).max() over 4e9 f32 on 8 GPUs = 5e8 per GPU, single pass, no compaction
    n5 = args.n5
    x = rt.empty(np.float32, n5)
    rt.fill(x, rank * n5, 11, 3)
    res = rt.empty(np.float32, 1)
    e = Expr(np.float32)
    xa = e.array()
    e.materialize(e.max(e.mul(e.filter(xa, e.ge(e.ref(xa), e.scalar(0.5))), e.scalar(3.0))))
    plan = abi.compile_plan(e)
    ptrs = abi.void_array([res.ptr])
    want5 = None
    if check:
        def c5_chunk(s0, m):
            return oracle.fast_filter_mul_max_f32(oracle.fill(np.float32, rank * n5 + s0, m, 11, 3), 0.5, 3.0, 8, 1)
        mine = max(oracle_stream(n5, c5_chunk, threads))
        want5 = max(float(v[0]) for v in ranks.all_gather_array(np.array([mine], dtype=np.float32)))

    def c5(_):
        rt.launch(plan, [res.ptr], [x.ptr, None, None], 0, n5, sizes.ptr)
        abi.check(lib.nl_finish_reduce(abi.OP_MAX, abi.NL_F32, ptrs, 0))

    def c5_fused(_):
        rt.launch_allreduce(plan, [res.ptr], [x.ptr, None, None], 0, n5, sizes.ptr)

    def c5_local(_):
        rt.launch(plan, [res.ptr], [x.ptr, None, None], 0, n5, sizes.ptr)

    def par5(v):
        return {"parity": verdict(all_ok(float(v) == want5), f"max over all {G * n5} elements = {float(v)!r}")} if check else {}
    if G > 1 and fused:
        sm = solo(lambda: timed(c5_local, k, w, solo=True)[0])       # weak scaling: one GPU's share IS the N = 1 workload
        ms, _ = timed(c5_fused, k, w)
        v_fused = float(rt.download(res)[0])
        ms_nccl, _ = timed(c5, k, w)
        v_nccl = float(rt.download(res)[0])
        ms_local, _ = timed(c5_local, k, w)
        extra = {"result": v_fused, "ms_per_step_nccl_finish": round(ms_nccl / k, 5), "ms_per_step_without_finish": round(ms_local / k, 5),
                 "same_as_nccl": v_fused == v_nccl, "scaling": "weak (5e8 per GPU at every N)"}
        extra.update(par5(v_fused))
        report("C5", "C5 (a[a >= t] * k).max() fused + NVLink mailbox finish (in-kernel)", "f32", n5, 4 * n5, ms, extra,
               solo_ms=sm, solo_bytes=4 * n5)
    else:
        ms, _ = timed(c5, k, w)
        v = float(rt.download(res)[0])
        extra = {"result": v}
        extra.update(par5(v))
        report("C5", "C5 (a[a >= t] * k).max() fused" + (" + NCCL allreduce" if G > 1 else ""), "f32", n5, 4 * n5, ms, extra)
    x.free(); res.free()

    # next row: the sort() full breaker as a device radix sort, one GPU's shard
    if G == 1 and not args.no_sort:
        n6 = 1 << args.log2n_sort
        for dt, kind, nm, sid in ((np.int64, 1, "i64 full range", "sort i64"), (np.float64, 3, "f64 in [0,1)", "sort f64"),
                                  (np.int32, 1, "i32 full range", "sort i32"), (np.int64, 0, "i64 in [1,100) like Tests/", "sort i64 1..99")):
            x = rt.empty(dt, n6)
            tmp = rt.empty(dt, n6)
            best, passes, sum_before, before = None, 0, None, None
            for rep in range(3):
                rt.fill(x, 0, 5 + rep, kind)
                if np.dtype(dt).kind == "i":
                    sum_before = _reduce_device(rt, abi, Expr, abi.OP_SUM, dt, x.ptr, n6, sizes)
                # 8-byte keys: sum / xor of the 64-bit words; 4-byte keys pair up differently after the sort, so their
                # second permutation-invariant figure is the wrapping sum of squares
                before = device_checksum(rt, lib, x.ptr, 8 * n6, csum) if np.dtype(dt).itemsize == 8 else _sum_of_squares(rt, abi, Expr, dt, x.ptr, n6, sizes)
                rt.sync()
                l0 = lib.nl_launch_count()
                e0, e1 = rt.event(), rt.event()
                rt.record(e0)
                abi.check(lib.nl_sort(0, 0, abi.nl_dtype_of(dt), C.c_void_p(x.ptr), C.c_void_p(tmp.ptr), n6))
                rt.record(e1)
                rt.sync()
                ms1 = rt.elapsed_ms(e0, e1)
                best = ms1 if best is None else min(best, ms1)
                passes = lib.nl_launch_count() - l0
            after = device_checksum(rt, lib, x.ptr, 8 * n6, csum) if np.dtype(dt).itemsize == 8 else _sum_of_squares(rt, abi, Expr, dt, x.ptr, n6, sizes)
            props = sort_properties(rt, abi, Expr, lib, x, n6, dt, sum_before, sizes)
            props["same_multiset_checksums"] = (before[:2] == after[:2]) if np.dtype(dt).itemsize == 8 else bool(before == after)
            good = props["inversions"] == 0 and props["same_multiset_checksums"] and props.get("sum_unchanged", True)
            out.append({"id": sid, "config": f"sort() device radix sort, {nm}", "dtype": np.dtype(dt).name, "elements_per_gpu": n6,
                        "ms_per_step": round(best, 4), "Gkeys/s": round(n6 / best / 1e6, 2), "kernel_launches": int(passes),
                        "properties_at_full_size": props,
                        "parity": ("bit-exact by properties: " if good else "MISMATCH: ") + "zero inversions over all keys and unchanged sum/xor checksums (sorted permutation of the input)"})
            x.free(); tmp.free()
    sizes.free(); csum.free()
    # sort() ACROSS the GPUs is the one-process mode of the jit drop-in (sample sort over NVLink, csrc/jit/device_sort.cpp):
    # rank 0 runs it in a process of its own over all G devices while the ranks — one GPU each — wait with empty pools
    if G > 1 and not args.no_sort:
        abi.check(lib.nl_pool_trim(0))
        ranks.barrier()
        if ranks.rank == 0:
            try:
                env = dict(os.environ, NL_DEVICES=str(G))
                for k_ in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "CUDA_VISIBLE_DEVICES"):
                    env.pop(k_, None)
                tool = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools", "jit_sort_bench.py")
                res = subprocess.run([sys.executable, tool, str(min(args.log2n_sort, 27)), "int64", "--json"], env=env, stdout=subprocess.PIPE,
                                     stderr=subprocess.PIPE, text=True, timeout=600)
                line = [l for l in res.stdout.splitlines() if l.startswith("{")]
                if res.returncode == 0 and line:
                    row = json.loads(line[-1])
                    out.append(row)
                else:
                    out.append({"id": f"sort x{G}", "parity": "unchecked", "error": (res.stderr or res.stdout)[-400:]})
            except Exception as exc:
                out.append({"id": f"sort x{G}", "parity": "unchecked", "error": f"{type(exc).__name__}: {exc}"})
        ranks.barrier()
    return out


# ------------------------------------------------------------------------------ CPU arms
def cpu_columns(n, threads):
    import oracle
    a, b = np.empty(n, np.float32), np.empty(n, np.float32)

    def gen(s, m):
        a[s:s + m] = oracle.fill(np.float32, s, m, 42, 2)
        b[s:s + m] = oracle.fill(np.float32, s, m, 43, 2)
        return 0
    oracle_stream(n, gen, threads)
    return a, b, np.empty(n, np.float32)


def cpu_mul2(a, b, c, reps, chunks, workers):
    """The oracle port of the headline pipeline (c = a*b, f32): the reference's chunking (scheduler.cpp:133-147)
    with `chunks` ranges on `workers` pthreads; one step = the forced pipeline + the in-place update."""
    import oracle
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        oracle.fast_mul2_f32(a, b, c, chunks, workers)
        a[0] = 1000.0                                    # the in-place update of the step
        times.append(time.perf_counter() - t0)
    return times


def cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_baseline(args):
    """the reference's CPU path in its three forms, at the config's own size."""
    threads = host_threads()
    n = 1 << args.log2n
    a, b, c = cpu_columns(n, threads)
    chunks = max(8, threads)
    cpu_mul2(a, b, c, 1, chunks, threads)                                   # page-faults the output
    intended = statistics.median(cpu_mul2(a, b, c, 5, chunks, threads))     # what the 8-chunk design is for
    faithful = statistics.median(cpu_mul2(a, b, c, 2, 8, 1))                # what the reference does: ONE worker (scheduler.cpp:227-231)
    t_np = []
    for _ in range(2):
        t0 = time.perf_counter()
        np.multiply(a, b, out=c)
        a[0] = 1000.0
        t_np.append(time.perf_counter() - t0)
    gbs = lambda t: round(12 * n / t / 1e9, 2)
    return {"value": gbs(intended), "unit": "GB/s", "cores": threads, "kind": "port",
            "sample": f"c = a*b on 2^{n.bit_length() - 1} f32 elements (the config's size) x 5 repetitions (median), gcc -O3 compiled loop, "
                      f"{chunks} chunks on {threads} pthreads ('intended' form of the 8-chunk scheduler)",
            "elements_per_s": round(n / intended, 1),
            "faithful_one_worker": {"value": gbs(faithful), "unit": "GB/s", "cores": 1,
                                    "sample": "8 chunks on the reference's single worker thread (scheduler.cpp:227-231), 2 repetitions"},
            "numpy": {"value": gbs(statistics.median(t_np)), "unit": "GB/s", "cores": 1, "sample": "numpy.multiply(a, b, out=c), 2 repetitions"},
            "cpu_model": cpu_model(), "nproc": os.cpu_count()}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference
    cannot be built here (CPython 2 + LLVM 3.9), so this is the oracle port with all host
    threads on the config's own size; rank 0 alone runs it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    n = 1 << args.log2n
    K, W = min(args.steps, 50), max(args.warmup, 1)
    a, b, c = cpu_columns(n, threads)
    chunks = max(8, threads)
    times = cpu_mul2(a, b, c, K + W, chunks, threads)[W:]
    total = sum(times)
    value = 12 * n * len(times) / total / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": "GB/s", "n_gpus": args.gpus,
            "steps": len(times), "warmup": W, "ms_per_step": round(1e3 * total / len(times), 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": headline_config(n, args.gpus),
            "cpu_baseline": {"value": round(value, 2), "unit": "GB/s", "cores": threads, "kind": "port",
                             "sample": f"one 2^{n.bit_length() - 1}-element shard (one GPU's share of the config) per step, {chunks} chunks on {threads} pthreads; "
                                       "rank 0's host only", "cpu_model": cpu_model(), "nproc": os.cpu_count()},
            "e2e": {"value": round(value, 2), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "elements_per_s": round(n * len(times) / total, 1)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=30, help="log2 of elements per GPU for C2 / of total elements for C4")
    ap.add_argument("--log2n-reduce", dest="log2n_reduce", type=int, default=31, help="log2 of total elements for C3")
    ap.add_argument("--log2n-sort", dest="log2n_sort", type=int, default=28)
    ap.add_argument("--n1", type=int, default=10_000_000, help="elements of C1")
    ap.add_argument("--n5", type=int, default=500_000_000, help="elements per GPU of C5")
    ap.add_argument("--e2e-steps", dest="e2e_steps", type=int, default=5)
    ap.add_argument("--cabi-e2e", dest="cabi_e2e", action="store_true", help="also time the chunk-streamed step through the bare C-ABI")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-sort", action="store_true")
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
