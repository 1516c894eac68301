V="${1:-v2 256x16 match0 ctas3 lb4}"
ncu --set full --clock-control none --import-source on --kernel-name-base function -k sweep2 -s 10 -c 1 -f -o gpurun_out/sort_lab_sweep2 build/sort_lab 27 "$V" > gpurun_out/sort_lab_ncu.log 2>&1
tail -3 gpurun_out/sort_lab_ncu.log
