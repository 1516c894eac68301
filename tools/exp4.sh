timeout 600 python -m pytest tests/test_gpu_compact_pipe.py tests/test_gpu_peer_reduce.py -m gpu -x -q 2>&1 | tail -3
python tools/sanitize_compact.py | tail -1
for c in c4_1 c4_10 c4_50 c4_90; do python tools/profile_one.py $c --log2n 30 --reps 200 --warmup 2 --trend; sleep 3; done
python tools/profile_one.py c4r_10 --log2n 30 --reps 45 --warmup 2
python tools/profile_one.py c4_10 --log2n 30 --reps 45 --warmup 2 --tier 2
