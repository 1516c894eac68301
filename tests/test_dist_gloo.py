"""N > 1 host logic on CPU: two gloo ranks shard a column, run the ORACLE on their shard, exchange
counts / partials the way the GPU path does (scan of shard counts, ordered combine of partials)
and must reproduce the single-process result.  Also checks that the Python shard_range()
mirrors the library's nl_partition()."""
import ctypes as C
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from numpyllvm_b200 import abi
from numpyllvm_b200.dist import combine_partials, exclusive_scan, shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("size", [0, 5, 1023, 1024, 100000, (1 << 30) + 12345])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_range_mirrors_nl_partition(size, world):
    lib = abi.load_library()
    prev = 0
    for r in range(world):
        s, e = C.c_int64(), C.c_int64()
        assert lib.nl_partition(size, world, 1024, r, C.byref(s), C.byref(e)) == 0
        assert shard_range(size, world, r) == (s.value, e.value)
        assert s.value == prev and s.value % 1024 == 0
        prev = e.value
    assert prev == size


def test_exclusive_scan_and_combine():
    assert exclusive_scan([9, 7, 10, 9, 7, 9, 9, 13]) == [0, 9, 16, 26, 35, 42, 51, 60, 73]   # Tests/filter.py chunk counts
    assert exclusive_scan([]) == [0]
    assert combine_partials("sum", [686, 265, 605, 480, 468, 340, 490, 783], np.int64) == 4117   # Tests/max.py partials
    assert combine_partials("max", [3.0, 9.5, -1.0], np.float64) == 9.5
    assert np.isnan(combine_partials("max", [3.0, np.nan], np.float32))
    with np.errstate(over="ignore"):
        assert combine_partials("sum", [2**62, 2**62, 2**62], np.int64) == np.int64(2**62) * 3   # wraps like the kernels


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np
    import oracle
    from numpyllvm_b200.abi import Expr
    from numpyllvm_b200.dist import Ranks, shard_range, exclusive_scan, combine_partials

    ranks = Ranks(2)
    assert ranks.world == 2
    n = 300007
    s, e = shard_range(n, ranks.world, ranks.rank)
    col = oracle.fill(np.float64, s, e - s, 9, 3)          # this rank's shard of the synthetic column
    whole = oracle.fill(np.float64, 0, n, 9, 3)

    # filter: per-shard compaction, counts scanned across ranks, runs concatenated in rank order
    ex = Expr(np.float64); xa = ex.array()
    ex.materialize(ex.filter(xa, ex.gt(ex.ref(xa), ex.scalar(0.5))))
    (mine,) = oracle.run(ex, [col, None], e - s, thread_count=1)
    counts = ranks.all_gather_int(len(mine))
    offs = exclusive_scan(counts)
    parts = ranks.all_gather_array(mine)
    merged = np.empty(offs[-1])
    for r, p in enumerate(parts):
        merged[offs[r]:offs[r + 1]] = p
    (ref,) = oracle.run(ex, [whole, None], n)
    assert np.array_equal(merged, ref), "sharded filter differs"

    # reductions: one partial per shard, folded in rank order
    for op in ("max", "min", "sum"):
        ex = Expr(np.int64)
        ex.materialize(getattr(ex, op)(ex.array()))
        icol = oracle.fill(np.int64, s, e - s, 3, 1)
        iwhole = oracle.fill(np.int64, 0, n, 3, 1)
        (p,) = oracle.run(ex, [icol], e - s, thread_count=1)
        allp = [int(v) for v in np.concatenate(ranks.all_gather_array(p))]
        got = combine_partials(op, allp, np.int64)
        (want,) = oracle.run(ex, [iwhole], n)
        assert got == want[0], (op, got, want)
    # timing plumbing
    assert ranks.max(float(ranks.rank)) == 1.0 and ranks.sum(1.0) == 2.0
    assert ranks.bcast_bytes(b"id-from-rank-%%d" %% ranks.rank) == b"id-from-rank-0"
    # NVLink mailbox handshake (host side): every rank exports a 64-byte handle, all ranks import the
    # same rank-ordered table; a rank that cannot export keeps EVERYBODY on the NCCL finish
    from numpyllvm_b200 import abi

    class FakeLib:
        def __init__(self, rank, can_export=True):
            self.rank, self.can_export, self.blob, self.ready = rank, can_export, None, 0
        def nl_comm_mailbox_export(self, dev, buf):
            if not self.can_export:
                return abi.NL_ERR_CUDA
            buf.raw = bytes([self.rank + 1]) * abi.NL_MAILBOX_HANDLE_BYTES
            return 0
        def nl_comm_mailbox_import(self, buf):
            self.blob = bytes(buf.raw[:2 * abi.NL_MAILBOX_HANDLE_BYTES]); self.ready = 1
            return 0
        def nl_comm_peer_ready(self):
            return self.ready

    lib = FakeLib(ranks.rank)
    assert ranks.init_mailboxes(lib) is True
    assert lib.blob == bytes([1]) * 64 + bytes([2]) * 64, "handles must arrive in rank order"
    lib = FakeLib(ranks.rank, can_export=(ranks.rank == 0))
    assert ranks.init_mailboxes(lib) is False and lib.blob is None

    ranks.finish()
    print("rank", ranks.rank, "ok")
""") % ROOT


def test_two_gloo_ranks_reproduce_the_single_process_result(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = 29600 + (os.getpid() % 300)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o}"
        assert f"rank {r} ok" in o
