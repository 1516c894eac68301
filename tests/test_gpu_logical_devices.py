"""The multi-shard paths on a box with ONE GPU: NL_LOGICAL_DEVICES=k lays k devices of the runtime over the visible
GPUs (own streams, scratch arena and mailbox each), so the in-kernel exchange between devices (reductions, the
count scan of filters), ragged filter results, repartition, gather, assignment across shards and the sort across
devices run exactly as they do between GPUs — peer stores land in the same HBM instead of crossing NVLink, and NCCL,
which refuses two ranks on one GPU, is left out.  The multi-device tests are run again in a process of their own
with two and three logical devices; every case checks against the oracle / NumPy as before."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_suite(k, files, extra=()):
    env = dict(os.environ, NL_LOGICAL_DEVICES=str(k), NL_DEVICES=str(k), CUDA_MODULE_LOADING="EAGER")
    cmd = [sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider", *extra, *files]
    res = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    tail = "\n".join(res.stdout.splitlines()[-25:])
    assert res.returncode == 0, tail
    return tail


@pytest.mark.parametrize("k", [2, 3])
def test_fused_exchange_between_logical_devices(k):
    tail = run_suite(k, ["tests/test_gpu_peer_reduce.py"])
    assert " passed" in tail and "needs two GPUs" not in tail


@pytest.mark.parametrize("k", [2, 3])
def test_jit_over_logical_devices(k):
    # the whole jit suite: sharded thunks, ragged filters, repartition, gather, assignment across shards, sort()
    tail = run_suite(k, ["tests/test_jit_gpu.py"])
    assert " passed" in tail


def test_the_runtime_reports_the_logical_devices():
    code = ("import os; os.environ['NL_LOGICAL_DEVICES']='4'; os.environ['NL_DEVICES']='4'\n"
            "import numpy as np\n"
            "from numpyllvm_b200 import jit\n"
            "a = jit.thunk(np.arange(100003, dtype=np.int64))\n"
            "assert jit.device_count() == 4 and len(a.device_shards()) == 4\n"
            "assert int((a[a >= 50000] * 2).sum().asnumpyarray()[0]) == int((np.arange(50000, 100003) * 2).sum())\n"
            "assert np.array_equal(jit.thunk(np.arange(100003, dtype=np.int64)[::-1].copy()).sort().asnumpyarray(), np.arange(100003))\n"
            "print('ok')\n")
    res = subprocess.run([sys.executable, "-c", code], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert res.returncode == 0 and "ok" in res.stdout, res.stdout[-2000:]
