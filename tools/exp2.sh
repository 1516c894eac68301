nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,temperature.memory --format=csv,noheader -lms 200 > gpurun_out/exp2_smi.txt &
SMI=$!
python tools/profile_one.py c4_10 --log2n 30 --reps 400 --warmup 2 --trend
sleep 3
python tools/profile_one.py c2 --log2n 30 --reps 300 --warmup 2 --trend
sleep 3
python tools/profile_one.py c4_10 --log2n 30 --reps 45 --warmup 2 --tune max_ctas_per_sm=3
python tools/profile_one.py c4_10 --log2n 28 --reps 400 --warmup 2 --trend
kill $SMI
awk 'NR%3==1' gpurun_out/exp2_smi.txt | head -60
