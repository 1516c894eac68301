# what runs on a GPU box before a round is handed in: every GPU test, the smoke check, the bench line
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1200 gpurun_out/bench_n1.json
