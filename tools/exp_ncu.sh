# ncu captures of the round: launch list of the bench command + `--set full` of the dominant kernels.
# The .ncu-rep files stay on the box (tens of MB each); their raw pages come back as CSV.
B="python bench.py --steps 5 --warmup 3 --no-extra --no-e2e --no-cpu --no-check"
$B > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_bench_c2.csv $B > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
mkdir -p /tmp/ncu
run() {   # name log2n pattern extra-args
  P="python tools/profile_one.py $1 --log2n $2 --reps 2 --warmup 1 $4"
  $P > gpurun_out/plain_$1$5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$3 -s 1 -c 1 -f -o /tmp/ncu/$1$5 $P > gpurun_out/ncu_$1$5.log 2>&1
  echo "$1$5 rc=$?"; tail -1 gpurun_out/plain_$1$5.log
  ncu -i /tmp/ncu/$1$5.ncu-rep --page raw --csv > gpurun_out/r2_ncu_$1$5_raw.csv 2>/dev/null
}
run c2 30 pipeline_kernel "" ""
run c4_10 28 compact_super "" ""
run c4_50 28 compact_super "" ""
run c4d_10 28 compact_super "" ""
run c4d_10 28 compact_super "--tier 0" _interp
ncu -i /tmp/ncu/c4_10.ncu-rep --page source --csv > gpurun_out/r2_ncu_c4_10_source.csv 2>/dev/null
ncu -i /tmp/ncu/c4d_10.ncu-rep --page source --csv > gpurun_out/r2_ncu_c4d_10_source.csv 2>/dev/null
du -sh gpurun_out
