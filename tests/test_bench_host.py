"""CPU checks of bench.py's host-side logic: the chunk-wise checksum algebra of the full-size parity
checks (the host twin of nl_checksum, nl_abi.h), the shared `config` of both arms and the reference arm."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int64, np.uint8])
def test_chunkwise_checksums_equal_the_one_shot_checksum(dtype):
    rng = np.random.default_rng(1)
    n = 100_003 if dtype == np.uint8 else 50_000
    x = rng.integers(0, 255, n * np.dtype(dtype).itemsize, dtype=np.uint8).view(dtype)
    whole = bench.words_checksum(x)
    # chunks that hold whole 64-bit words
    step = 4096
    parts = [bench.words_checksum(x[s:s + step]) for s in range(0, x.size, step)]
    assert bench.combine_checksums(parts) == whole
    # the weighted sum sees the ORDER, the plain sums do not
    y = x.copy()
    y[:8], y[8:16] = x[8:16].copy(), x[:8].copy()
    if not np.array_equal(x[:8], x[8:16]):
        other = bench.words_checksum(y)
        assert other[1:3] == whole[1:3] and other[3] != whole[3]


def test_words_checksum_definition():
    w = np.array([3, 5, 7], dtype=np.uint64)
    assert bench.words_checksum(w) == (3, 15, 3 ^ 5 ^ 7, 3 * 1 + 5 * 3 + 7 * 5)
    assert bench.words_checksum(np.array([1], dtype=np.uint8)) == (1, 1, 1, 1)          # zero-padded tail word
    big = np.full(4, 2 ** 63, dtype=np.uint64)
    assert bench.words_checksum(big)[1] == 0                                             # sums wrap mod 2^64
    assert bench.words_checksum(np.empty(0)) == (0, 0, 0, 0)


def test_both_arms_print_the_same_config():
    assert bench.headline_config(1 << 30, 4) == bench.headline_config(1 << 30, 4)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--log2n", "20", "--steps", "3",
                          "--warmup", "1", "--gpus", "2"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["config"] == bench.headline_config(1 << 20, 2)
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port"
    # the other ranks of a torchrun launch print nothing and exit 0
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--log2n", "20", "--gpus", "2"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
