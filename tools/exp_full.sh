python -m pytest tests/test_gpu_parity.py -x -q -k "arena or graph or recycled or filter" 2>&1 | tail -3
