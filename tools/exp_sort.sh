NL_DEBUG_ALLOC=1 NL_DEBUG_SORT=1 python tools/jit_sort_bench.py 27 int64 2>&1 | tail -14
python tools/jit_sort_bench.py 27 float32 2>&1 | tail -1
python tools/jit_sort_bench.py 28 int32 2>&1 | tail -1
python -m pytest tests/test_jit_gpu.py -x -q 2>&1 | tail -3
python -m pytest tests/test_gpu_peer_reduce.py tests/test_gpu_sort.py -x -q 2>&1 | tail -3
