timeout 600 python -m pytest tests/test_gpu_compact_pipe.py tests/test_gpu_peer_reduce.py -m gpu -x -q 2>&1 | tail -2
python tools/sanitize_compact.py | tail -1
for c in c4_1 c4_10 c4_50 c4_90 c4d_10 c4d_50; do python tools/profile_one.py $c --log2n 30 --reps 100 --warmup 2 --trend; sleep 2; done
