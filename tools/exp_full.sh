timeout 1200 python -m pytest tests/test_gpu_logical_devices.py -x -q 2>&1 | tail -15
NL_DEVICES=1 python tools/jit_sort_bench.py 27 int64 | tail -1
