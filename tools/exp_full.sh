python bench.py > gpurun_out/r2_bench_n1_v6.json 2> gpurun_out/r2_bench_n1_v6.err; tail -c 1100 gpurun_out/r2_bench_n1_v6.json
