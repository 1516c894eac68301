python tools/sort_bench.py 28 0 > gpurun_out/r2_sort_bench.txt 2>&1; cat gpurun_out/r2_sort_bench.txt
ncu --set full --clock-control none --import-source on --kernel-name-base function -k onesweep_kernel -s 9 -c 1 -f -o gpurun_out/r2_sort_onesweep python tools/sort_bench.py 27 0 > gpurun_out/r2_sort_ncu.log 2>&1; tail -2 gpurun_out/r2_sort_ncu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_sort_launches.csv python tools/sort_bench.py 28 0 > /dev/null 2>&1; tail -3 gpurun_out/r2_sort_launches.csv
