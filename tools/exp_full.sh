python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest4.log 2>&1; tail -3 gpurun_out/r2_pytest4.log
python bench.py > gpurun_out/r2_bench_n1_v3.json 2> gpurun_out/r2_bench_n1_v3.err; tail -c 1500 gpurun_out/r2_bench_n1_v3.json
