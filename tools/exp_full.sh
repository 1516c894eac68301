python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_2gpu_v2.log 2>&1; tail -3 gpurun_out/r2_pytest_2gpu_v2.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
