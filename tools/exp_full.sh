for rep in 1 2; do for sc in 1 0; do echo -n "self_clean=$sc "; timeout 120 python tools/profile_one.py c4_1 --log2n 30 --reps 200 --warmup 25 --tune compact_self_clean=$sc | tail -1; done; done
for sc in 1 0; do echo -n "self_clean=$sc "; timeout 120 python tools/profile_one.py c4_50 --log2n 30 --reps 100 --warmup 25 --tune compact_self_clean=$sc | tail -1; done
nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv,noheader
