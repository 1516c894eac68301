# sustained vs short runs of the compaction configs, clocks sampled beside them
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.active --format=csv,noheader -lms 100 > gpurun_out/exp1_smi.txt &
SMI=$!
for reps in 3 45 200; do python tools/profile_one.py c4_10 --log2n 30 --reps $reps --warmup 2; done
for reps in 3 45; do python tools/profile_one.py c4_1 --log2n 30 --reps $reps --warmup 2; python tools/profile_one.py c4_50 --log2n 30 --reps $reps --warmup 2; python tools/profile_one.py c4_90 --log2n 30 --reps $reps --warmup 2; done
for c in c4d_10 c4d_50 c4b_50; do python tools/profile_one.py $c --log2n 30 --reps 20 --warmup 2;  python tools/profile_one.py $c --log2n 30 --reps 20 --warmup 2 --tier 0; done
python tools/profile_one.py c2 --log2n 30 --reps 45 --warmup 2
python tools/interp_bench.py 2>&1 | tail -30
kill $SMI
sort gpurun_out/exp1_smi.txt | uniq -c | sort -rn | head -12
