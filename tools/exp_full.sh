python tools/sort_bench.py 28 0 > gpurun_out/r2_sort_bench.txt 2>&1; cat gpurun_out/r2_sort_bench.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest5.log 2>&1; tail -2 gpurun_out/r2_pytest5.log
python bench.py > gpurun_out/r2_bench_n1_v4.json 2> gpurun_out/r2_bench_n1_v4.err; tail -c 1200 gpurun_out/r2_bench_n1_v4.json
