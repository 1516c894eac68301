python -m pytest tests/test_gpu_sort.py -x -q 2>&1 | tail -5
python tools/sort_bench.py 28 0 2>&1 | tail -7
python tools/sort_bench.py 30 0 2>&1 | tail -7
