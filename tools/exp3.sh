nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader -lms 100 > gpurun_out/exp3_smi.txt &
SMI=$!
for ns in 0 20 100 400; do echo "poll_sleep_ns=$ns"; python tools/profile_one.py c4_10 --log2n 30 --reps 300 --warmup 2 --trend --tune poll_sleep_ns=$ns; sleep 4; done
echo "max 3 ctas"; python tools/profile_one.py c4_10 --log2n 30 --reps 300 --warmup 2 --trend --tune max_ctas_per_sm=3; sleep 4
echo "sel 50"; python tools/profile_one.py c4_50 --log2n 30 --reps 300 --warmup 2 --trend; sleep 4
echo "c3 max"; python tools/profile_one.py c3_f64_max --log2n 31 --reps 200 --warmup 2 --trend
kill $SMI
sort -t, -k2 -n -r gpurun_out/exp3_smi.txt | head -8
