# sort: device tests, nl_sort on 2^28 keys, sort() through the jit API over every visible GPU
python -m pytest tests/test_gpu_sort.py tests/test_jit_gpu.py -x -q 2>&1 | tail -3
python tools/sort_bench.py 28 0 2>&1 | tail -6
NL_DEBUG_SORT=1 python tools/jit_sort_bench.py 27 int64 2>&1 | tail -9
python tools/jit_sort_bench.py 27 float64 2>&1 | tail -1
python tools/jit_sort_bench.py 28 int32 2>&1 | tail -1
