python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_2gpu_v3.log 2>&1; tail -2 gpurun_out/r2_pytest_2gpu_v3.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_v3.json 2> gpurun_out/r2_bench_n2_v3.err
tail -c 900 gpurun_out/r2_bench_n2_v3.json; tail -2 gpurun_out/r2_bench_n2_v3.err
