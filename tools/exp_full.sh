for rep in 1 2; do
for c in c4_1 c4_10; do for sc in 1 0; do echo -n "self_clean=$sc "; timeout 120 python tools/profile_one.py $c --log2n 27 --reps 40 --warmup 10 --tune compact_self_clean=$sc | tail -1; done; done
done
for c in c4_1 c4_50; do for sc in 1 0; do echo -n "self_clean=$sc "; timeout 120 python tools/profile_one.py $c --log2n 30 --reps 20 --warmup 5 --tune compact_self_clean=$sc | tail -1; done; done
for c in c4_1; do for sc in 1 0; do echo -n "self_clean=$sc "; timeout 120 python tools/profile_one.py $c --log2n 22 --reps 100 --warmup 10 --tune compact_self_clean=$sc | tail -1; done; done
