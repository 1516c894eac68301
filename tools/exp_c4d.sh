# ncu --set full of the super-tile compaction kernel (static and family shape) — each command plain first
mkdir -p /tmp/ncu
for c in c4_10 c4_50 c4d_50; do
P="python tools/profile_one.py $c --log2n 28 --reps 2 --warmup 1"
$P > gpurun_out/plain_$c.log 2>&1; tail -1 gpurun_out/plain_$c.log
ncu --set full --clock-control none --import-source on -k regex:compact_super -s 1 -c 1 -f -o /tmp/ncu/$c $P > gpurun_out/ncu_$c.log 2>&1
ncu -i /tmp/ncu/$c.ncu-rep --page raw --csv > gpurun_out/r2c_ncu_${c}_raw.csv 2>/dev/null
done
