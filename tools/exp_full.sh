python -m pytest tests/test_gpu_parity.py tests/test_gpu_compact_pipe.py -x -q 2>&1 | tail -2
for c in c4_1 c4_10 c4_50 c4_90 c4d_10 c4d_50; do for v in 1 0; do echo -n "early=$v "; timeout 120 python tools/profile_one.py $c --log2n 30 --reps 60 --warmup 20 --tune super_early=$v | tail -1; done; done
for c in c4_50; do for v in 1 0; do echo -n "early=$v "; timeout 120 python tools/profile_one.py $c --log2n 27 --reps 60 --warmup 20 --tune super_early=$v | tail -1; done; done
